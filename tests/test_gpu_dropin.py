"""End-to-end drop-in check of the extraction stage: frozen .pb + feats.scp in, embedding ark/scp out, through the
same command line as the reference's tf_extract.py (tensorflow/tf_extract.py:46-113), then scored by the snorm drop-in
and read back by the reference-format ark reader."""
import os

import numpy as np
import pytest

from oracle import net_oracle
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
        want = net_oracle.extract_utterance(cfg, params, kaldi_ark.apply_cmvn_sliding(m))
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
