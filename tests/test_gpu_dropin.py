"""End-to-end drop-in check of the extraction stage: frozen .pb + feats.scp in, embedding ark/scp out, through the
same command line as the reference's tf_extract.py (tensorflow/tf_extract.py:46-113), then scored by the snorm drop-in
and read back by the reference-format ark reader."""
import os

import numpy as np
import pytest

from oracle import cmn_oracle, net_oracle
from voxsrc2020_speaker_verification_b200 import arch, kaldi_ark, pb_loader

pytestmark = pytest.mark.gpu


def _write_feats(tmp_path, utts):
    ark = str(tmp_path / "raw_fbank.ark")
    scp = str(tmp_path / "feats.1.scp")
    with open(ark, "wb") as fa, open(scp, "w") as fs:
        for key, m in utts.items():
            off = kaldi_ark.write_mat(fa, m, key)
            fs.write("%s %s:%d\n" % (key, ark, off))
    return scp[:-4]


def test_tf_extract_cli_end_to_end(tmp_path):
    from voxsrc2020_speaker_verification_b200 import tf_extract
    cfg = arch.get_config("tdnn")
    params = net_oracle.init_params(cfg, 40, seed=4321)
    pb = str(tmp_path / "tdnn.pb")
    pb_loader.write_pb(pb, params, cfg, 40)
    rng = np.random.default_rng(3)
    utts = {"id1/a.wav": None, "id1/b.wav": None, "id2/c.wav": None, "id3/long.wav": None}
    for key, t in zip(utts, (57, 301, 25, 1130)):
        utts[key] = (rng.standard_normal((t, 40)) * 2 + rng.standard_normal(40) * 5).astype(np.float32)   # raw FBANK-like: non-zero mean
    rspec = _write_feats(tmp_path, utts)
    wspec = str(tmp_path / "xvector.1")
    assert tf_extract.main(["--pb-file", pb, "--expand-dim", "2", "--rspec", rspec, "--wspec", wspec]) == 0
    # output ark + scp, in input order, readable record by record
    got = dict(kaldi_ark.read_vec_flt_ark(wspec + ".ark"))
    assert list(got) == list(utts)
    scp = kaldi_ark.read_scp(wspec + ".scp")
    assert [k for k, _, _ in scp] == list(utts)
    for key, path, off in scp:
        with open(path, "rb") as f:
            f.seek(off)
            np.testing.assert_array_equal(kaldi_ark.read_vec_flt(f), got[key])
    # values: the oracle on host-normalised features with the chunk rule (1130 frames → 1000 + 130)
    for key, m in utts.items():
        want = net_oracle.extract_utterance(cfg, params, cmn_oracle.apply_cmvn_sliding(m))
        cos = float(np.dot(got[key], want) / np.linalg.norm(got[key]) / np.linalg.norm(want))
        assert cos >= 0.9999, (key, cos)


def test_tf_extract_cli_rejects_short_utterance(tmp_path):
    from voxsrc2020_speaker_verification_b200 import tf_extract
    cfg = arch.get_config("tdnn")
    params = net_oracle.init_params(cfg, 40, seed=4321)
    pb = str(tmp_path / "tdnn.pb")
    pb_loader.write_pb(pb, params, cfg, 40)
    rspec = _write_feats(tmp_path, {"short": np.zeros((24, 40), np.float32)})
    with pytest.raises(ZeroDivisionError):    # tf_extract.py:102,111 divides by zero for < 25 frames
        tf_extract.main(["--pb-file", pb, "--expand-dim", "2", "--rspec", rspec, "--wspec", str(tmp_path / "x")])


def test_eval_inference_model_pipeline(tmp_path):
    """eval_inference_model.sh stand-in on a miniature data directory: per-GPU scp shards in, concatenated arks and
    cosine_/snorm_ score files out, scores equal to the score oracle fed with the oracle's embeddings."""
    from oracle import score_oracle
    from voxsrc2020_speaker_verification_b200 import eval_inference_model
    cfg = arch.get_config("tdnn")
    params = net_oracle.init_params(cfg, 40, seed=4321)
    pb = str(tmp_path / "model.pb")
    pb_loader.write_pb(pb, params, cfg, 40)
    rng = np.random.default_rng(11)
    data = tmp_path / "data"
    feats = {}
    for ds, spk, per in (("voxceleb2_dev", 40, 2), ("voxceleb1", 6, 2)):
        d = data / ds / "1-split"
        d.mkdir(parents=True)
        ark = str(data / ds / "raw.ark")
        with open(ark, "wb") as fa, open(str(d / "feats.1.scp"), "w") as fs, open(str(data / ds / "spk2utt"), "w") as fu:
            for s in range(spk):
                utts = []
                for u in range(per):
                    key = "%s_id%02d/u%d.wav" % (ds[-4:], s, u)
                    m = (rng.standard_normal((int(rng.integers(40, 90)), 40)) + s).astype(np.float32)
                    off = kaldi_ark.write_mat(fa, m, key)
                    fs.write("%s %s:%d\n" % (key, ark, off))
                    feats[key] = m
                    utts.append(key)
                fu.write("%s_id%02d %s\n" % (ds[-4:], s, " ".join(utts)))
    test_keys = [k for k in feats if k.startswith("leb1")]
    (data / "voxceleb1_trials").mkdir()
    with open(str(data / "voxceleb1_trials" / "list_test_T.txt"), "w") as f:
        for i in range(12):
            a, b = rng.choice(len(test_keys), 2, replace=False)
            f.write("%d %s %s\n" % (i % 2, test_keys[a], test_keys[b]))
    assert eval_inference_model.main([pb, "2", "--data-root", str(data), "--num-gpus", "1", "--topk", "20"]) == 0
    out = str(tmp_path / "model_embeddings" / "voxceleb1")
    got = [ln.split() for ln in open(os.path.join(out, "snorm_T.txt"))]
    cos = [ln.split() for ln in open(os.path.join(out, "cosine_T.txt"))]
    assert len(got) == 12 and len(cos) == 12
    # (1) the scoring stage on the embeddings the extraction stage WROTE (cohort of 40 speakers, top-20): the reference's scoring
    #     restatement on the same arks must agree within the north star's 1e-3 absolute
    emb_dir = str(tmp_path / "model_embeddings")
    test_w = score_oracle.normalise_xvectors(dict(kaldi_ark.read_vec_flt_ark(os.path.join(emb_dir, "voxceleb1", "xvector.ark"))))
    spk2utt = score_oracle.read_spk2utt(str(data / "voxceleb2_dev" / "spk2utt"))
    cohort_w = score_oracle.read_speaker_xvector(
        score_oracle.normalise_xvectors(dict(kaldi_ark.read_vec_flt_ark(os.path.join(emb_dir, "voxceleb2_dev", "xvector.ark")))), spk2utt)
    mean, std = score_oracle.get_cohort_mean_std(test_w, cohort_w, 20)
    for (a, b, s), (a2, b2, c) in zip(got, cos):
        assert (a, b) == (a2, b2)
        want_c = float(np.dot(test_w[a], test_w[b]))
        want_s = 0.5 * ((want_c - mean[a]) / std[a] + (want_c - mean[b]) / std[b])
        assert abs(float(c) - want_c) < 1e-5
        assert abs(float(s) - want_s) < 1e-3
    # (2) end to end against the oracle's own embeddings of the CMN-normalised features: the embedding tolerance (cosine >= 0.9999)
    #     is divided by the cohort std (~0.05) in the normalised score
    emb = {k: net_oracle.extract_utterance(cfg, params, cmn_oracle.apply_cmvn_sliding(m)) for k, m in feats.items()}
    test = score_oracle.normalise_xvectors({k: emb[k] for k in test_keys})
    cohort = score_oracle.read_speaker_xvector(score_oracle.normalise_xvectors({k: v for k, v in emb.items() if k.startswith("_dev")}), spk2utt)
    mean, std = score_oracle.get_cohort_mean_std(test, cohort, 20)
    for (a, b, s), (_, _, c) in zip(got, cos):
        want_c = float(np.dot(test[a], test[b]))
        want_s = 0.5 * ((want_c - mean[a]) / std[a] + (want_c - mean[b]) / std[b])
        assert abs(float(c) - want_c) < 1e-3
        assert abs(float(s) - want_s) < 2e-2 * max(1.0, abs(want_s))


def test_compressed_matrix_decode_on_device_is_bit_exact(golden_dir):
    """svx_decode_compressed on the reference-written 'CM ' ark vs the matrices the reference's own kaldi_io decoded from it
    (tests/golden/io/ref_feats.npz, generated by oracle/gen_golden.py from /root/reference/tensorflow/kaldi_io.py)."""
    from voxsrc2020_speaker_verification_b200.extractor import Extractor
    g = os.path.join(golden_dir, "io")
    ref = np.load(os.path.join(g, "ref_feats.npz"))
    recs = list(kaldi_ark.read_mat_ark_raw(os.path.join(g, "feats_cm.ark")))
    assert recs and all(kind == "CM " for _, kind, _, _, _ in recs)
    cfg = arch.get_config("tdnn")
    for feat_dim in sorted({c for _, _, _, _, c in recs}):          # the fixture mixes 40- and 80-bin matrices
        group = [r for r in recs if r[4] == feat_dim]
        ex = Extractor("tdnn", feat_dim).load_params(net_oracle.init_params(cfg, feat_dim, seed=1, calib_frames=32, calib_batch=2))
        dev = ex.decode_compressed([p for _, _, p, _, _ in group], [r for _, _, _, r, _ in group]).cpu().numpy()
        row = 0
        for key, _, _, rows, cols in group:
            want = ref["cm:" + key]
            assert want.shape == (rows, cols)
            np.testing.assert_array_equal(dev[row:row + rows], want)
            row += rows


def test_tf_extract_cli_on_compressed_ark(tmp_path):
    """Compressed ('CM ') FBANK in: records are decoded, mean-normalised and embedded on the device; the result must equal the
    oracle run on the host-decoded, host-normalised matrices."""
    import struct
    from voxsrc2020_speaker_verification_b200 import tf_extract
    cfg = arch.get_config("tdnn")
    params = net_oracle.init_params(cfg, 40, seed=4321)
    pb = str(tmp_path / "tdnn.pb")
    pb_loader.write_pb(pb, params, cfg, 40)
    rng = np.random.default_rng(8)
    ark, scp = str(tmp_path / "cm.ark"), str(tmp_path / "feats.1.scp")
    keys = []
    with open(ark, "wb") as fa, open(scp, "w") as fs:
        for i, rows in enumerate((64, 33, 410)):
            key = "id%d/x.wav" % i
            keys.append(key)
            fa.write((key + " ").encode())
            off = fa.tell()
            heads = np.sort(rng.integers(0, 65536, (40, 4)), axis=1).astype("<u2")       # any sorted percentiles are a valid header
            fa.write(b"\0BCM " + struct.pack("<ffii", -12.5, 30.0, rows, 40) + heads.tobytes() +
                     rng.integers(0, 256, (40, rows)).astype(np.uint8).tobytes())
            fs.write("%s %s:%d\n" % (key, ark, off))
    wspec = str(tmp_path / "xv")
    assert tf_extract.main(["--pb-file", pb, "--expand-dim", "2", "--rspec", scp[:-4], "--wspec", wspec]) == 0
    got = dict(kaldi_ark.read_vec_flt_ark(wspec + ".ark"))
    assert list(got) == keys
    for key, m in kaldi_ark.read_mat_scp(scp):
        want = net_oracle.extract_utterance(cfg, params, cmn_oracle.apply_cmvn_sliding(np.asarray(m, np.float32)))
        cos = float(np.dot(got[key], want) / np.linalg.norm(got[key]) / np.linalg.norm(want))
        assert cos >= 0.9999, (key, cos)


def test_feature_dimension_mismatch_is_rejected(tmp_path, golden_dir):
    """A 40-dim ark fed to an 80-dim model (or the reverse) must fail loudly on both paths — the dense one and the compressed one,
    whose device decoder would otherwise index the column headers and the byte plane with the wrong width."""
    from voxsrc2020_speaker_verification_b200 import lib, tf_extract
    from voxsrc2020_speaker_verification_b200.extractor import Extractor
    cfg = arch.get_config("tdnn")
    params = net_oracle.init_params(cfg, 80, seed=4321, calib_frames=32, calib_batch=2)
    pb = str(tmp_path / "tdnn80.pb")
    pb_loader.write_pb(pb, params, cfg, 80)
    g = os.path.join(golden_dir, "io")
    recs = [r for r in kaldi_ark.read_mat_ark_raw(os.path.join(g, "feats_cm.ark")) if r[4] == 40]
    assert recs
    ex = Extractor("tdnn", 80).load_params(params)
    with pytest.raises(ValueError):
        ex.decode_compressed([p for _, _, p, _, _ in recs], [r for _, _, _, r, _ in recs])
    # straight through the C ABI with a lying caller: the library checks every record's own header on the device
    import ctypes
    import torch
    payloads, rows = [p for _, _, p, _, _ in recs], [r for _, _, _, r, _ in recs]
    rec_off = np.zeros(len(payloads), np.int64)
    np.cumsum([len(p) for p in payloads[:-1]], out=rec_off[1:])
    offs = np.zeros(len(payloads) + 1, np.int32)
    np.cumsum([r // 3 for r in rows], out=offs[1:])            # a third of the rows, so that rows x 80 fits the record: only the header is wrong
    blob = torch.frombuffer(bytearray(b"".join(payloads)), dtype=torch.uint8).cuda()
    out = torch.zeros((int(offs[-1]), 80), dtype=torch.float32, device="cuda")
    L = lib.load()
    st = L.svx_decode_compressed(ctypes.c_void_p(blob.data_ptr()), int(blob.numel()), rec_off.ctypes.data_as(ctypes.c_void_p),
                                 offs.ctypes.data_as(ctypes.c_void_p), len(payloads), 80, ctypes.c_void_p(out.data_ptr()),
                                 ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert st != 0 and b"header" in L.svx_last_error()
    assert float(out.abs().sum()) == 0.0                        # nothing was decoded
    # the CLI on an uncompressed 40-dim ark
    rspec = _write_feats(tmp_path, {"a": np.zeros((30, 40), np.float32)})
    with pytest.raises(ValueError):
        tf_extract.main(["--pb-file", pb, "--expand-dim", "2", "--rspec", rspec, "--wspec", str(tmp_path / "xv")])
