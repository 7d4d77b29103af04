"""Generate tests/golden/* by running the REFERENCE's own Python (snorm.py, kaldi_io.py) in the build
container.  Run once:  python oracle/gen_golden.py   (needs /root/reference; the fixtures travel, it does not).

Fixtures:
  score/test.ark, cohort.ark, cohort_spk2utt, test_spk2utt, trials.txt     inputs, written with the
                                                                            reference kaldi_io.write_vec_flt
  score/ref_topk{50,300,400}.npz   mean/std from reference snorm.get_cohort_mean_std (+ keys)
  score/ref_cosine.txt, ref_snorm_topk400.txt   score files exactly as reference snorm.py __main__ writes them
  score/ref_spk_*.{txt,npz}        same with --test_spk2utt speaker-level enrolment
  io/feats_cm.ark, feats_fm.ark, ref_feats.npz    matrices as decoded by reference kaldi_io.read_mat_ark
  eer/trials.txt, scores.txt, ref_eer.npy, ref_stdout.txt   what the reference's eer_minDCF.py computes / prints for them
  net/<model>.npz                  oracle (NOT reference: TensorFlow is absent) embeddings on seeded inputs —
                                   a regression pin for the restatement only
"""
import os
import struct
import subprocess
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/tensorflow"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")
import kaldi_io  # noqa: E402  (the reference's)
import snorm     # noqa: E402  (the reference's)


def l2(x):
    return x / np.linalg.norm(x, axis=-1, keepdims=True)


def gen_score(out):
    os.makedirs(out, exist_ok=True)
    rng = np.random.default_rng(20201)
    d, n_test, n_spk = 64, 1100, 300
    # test vectors: clustered so that cosine scores span a realistic range; un-normalised on disk
    centres = rng.standard_normal((40, d)).astype(np.float32)
    test = (centres[rng.integers(0, 40, n_test)] + 0.8 * rng.standard_normal((n_test, d))).astype(np.float32)
    test *= rng.uniform(0.5, 3.0, (n_test, 1)).astype(np.float32)
    test_keys = ["id%05d/vid%03d/%05d.wav" % (10000 + i // 7, i % 7, i) for i in range(n_test)]
    test[17] = test[3]                      # identical vectors → tied cohort rows
    with open(os.path.join(out, "test.ark"), "wb") as f:
        for k, v in zip(test_keys, test):
            kaldi_io.write_vec_flt(f, v, key=k)
    # cohort utterances: 3 per speaker, speakers 290..299 duplicate speakers 0..9 exactly (ties in top-k)
    spk_c = (centres[rng.integers(0, 40, n_spk)] + 0.9 * rng.standard_normal((n_spk, d))).astype(np.float32)
    utts, keys, spk2utt = [], [], {}
    for s in range(n_spk):
        src = s - 290 if s >= 290 else s
        r = np.random.default_rng(7000 + src)
        for j in range(3):
            keys.append("c%04d-%d" % (s, j))
            utts.append((spk_c[src] + 0.5 * r.standard_normal(d)).astype(np.float32))
            spk2utt.setdefault("c%04d" % s, []).append(keys[-1])
    order = rng.permutation(len(keys))     # ark order != spk2utt order
    with open(os.path.join(out, "cohort.ark"), "wb") as f:
        for i in order:
            kaldi_io.write_vec_flt(f, utts[i], key=keys[i])
    with open(os.path.join(out, "cohort_spk2utt"), "w") as f:
        for s, us in spk2utt.items():
            f.write(s + " " + " ".join(us) + "\n")
    # trial list: 'label utt1 utt2' (Vox1-O style), includes repeated pairs and self pairs
    idx1 = rng.integers(0, n_test, 2000)
    idx2 = rng.integers(0, n_test, 2000)
    idx2[:5] = idx1[:5]
    with open(os.path.join(out, "trials.txt"), "w") as f:
        for a, b in zip(idx1, idx2):
            f.write("%d %s %s\n" % (int(test_keys[a].split("/")[0] == test_keys[b].split("/")[0]), test_keys[a], test_keys[b]))
    # ---- run the reference functions
    xv = snorm.read_xvector(os.path.join(out, "test.ark"))
    cohort = snorm.get_cohort_xvector(os.path.join(out, "cohort.ark"), os.path.join(out, "cohort_spk2utt"))
    np.savez(os.path.join(out, "ref_cohort.npz"), keys=np.array(list(cohort.keys())),
             matrix=np.array(list(cohort.values())))
    for topk in (50, 300, 400):
        m, s = snorm.get_cohort_mean_std(xv, cohort, topk=topk)
        np.savez(os.path.join(out, "ref_topk%d.npz" % topk), keys=np.array(list(m.keys())),
                 mean=np.array(list(m.values()), np.float32), std=np.array(list(s.values()), np.float32))
    # ---- run the reference CLI end to end (default topk = 400 > cohort size)
    subprocess.check_call([sys.executable, os.path.join(REF, "snorm.py"),
                           "--trial", os.path.join(out, "trials.txt"), "--test_ark", os.path.join(out, "test.ark"),
                           "--cosine_score", os.path.join(out, "ref_cosine.txt"),
                           "--cohort_ark", os.path.join(out, "cohort.ark"),
                           "--cohort_spk2utt", os.path.join(out, "cohort_spk2utt"),
                           "--snorm_score", os.path.join(out, "ref_snorm_topk400.txt")],
                          cwd=REF, stderr=subprocess.DEVNULL)
    # ---- speaker-level enrolment (--test_spk2utt): models 'spkNN' = mean of 4 test utterances
    with open(os.path.join(out, "test_spk2utt"), "w") as f:
        for s in range(20):
            f.write("spk%02d " % s + " ".join(test_keys[s * 4 + j] for j in range(4)) + "\n")
    with open(os.path.join(out, "trials_spk.txt"), "w") as f:
        for i in range(300):
            f.write("0 spk%02d %s\n" % (i % 20, test_keys[100 + i]))
    subprocess.check_call([sys.executable, os.path.join(REF, "snorm.py"),
                           "--trial", os.path.join(out, "trials_spk.txt"), "--test_ark", os.path.join(out, "test.ark"),
                           "--test_spk2utt", os.path.join(out, "test_spk2utt"),
                           "--cosine_score", os.path.join(out, "ref_spk_cosine.txt"),
                           "--cohort_ark", os.path.join(out, "cohort.ark"),
                           "--cohort_spk2utt", os.path.join(out, "cohort_spk2utt"),
                           "--snorm_score", os.path.join(out, "ref_spk_snorm.txt")],
                          cwd=REF, stderr=subprocess.DEVNULL)


def gen_eer(out):
    """Trial / score files with realistic class overlap, ties and repeated pairs, and what the REFERENCE's eer_minDCF.py prints for
    them (three operating points).  eer_minDCF.py:68-94."""
    os.makedirs(out, exist_ok=True)
    import eer_minDCF  # noqa: E402  (the reference's)
    rng = np.random.default_rng(31337)
    n = 2500
    label = (rng.random(n) < 0.12).astype(int)
    score = np.where(label == 1, rng.normal(0.55, 0.18, n), rng.normal(0.12, 0.16, n)).astype(np.float32)
    score[rng.integers(0, n, 200)] = np.float32(0.25)            # a block of tied scores across both classes
    score[:50] = np.round(score[:50], 1)                         # more ties
    keys = ["u%03d" % i for i in range(400)]
    pairs = [(keys[rng.integers(0, 400)], keys[rng.integers(0, 400)]) for _ in range(n)]
    with open(os.path.join(out, "trials.txt"), "w") as ft, open(os.path.join(out, "scores.txt"), "w") as fs:
        for (a, b), l, s_ in zip(pairs, label, score):
            ft.write("%d %s %s\n" % (l, a, b))
            fs.write("%s %s %s\n" % (a, b, str(np.float32(s_))))
    # the reference keys both files by (utt1, utt2): repeated pairs keep the LAST line (eer_minDCF.py:25-40)
    pair_label = eer_minDCF.read_trial_file(os.path.join(out, "trials.txt"))
    pair_score = eer_minDCF.read_score_file(os.path.join(out, "scores.txt"))
    y = [pair_label[p] for p in pair_label]
    y_pred = [pair_score[p] for p in pair_label]
    rows = []
    for c_miss, c_fa, p_target in ((1, 1, 0.01), (1, 1, 0.05), (10, 1, 0.001)):
        rows.append([c_miss, c_fa, p_target] + [float(v) for v in eer_minDCF.compute_eer_and_min_dcf(y, y_pred, c_miss, c_fa, p_target)])
    np.save(os.path.join(out, "ref_eer.npy"), np.array(rows, np.float64))
    with open(os.path.join(out, "ref_stdout.txt"), "w") as f:
        f.write(subprocess.check_output([sys.executable, os.path.join(REF, "eer_minDCF.py"), "--trial", os.path.join(out, "trials.txt"),
                                         "--score", os.path.join(out, "scores.txt")], cwd=REF, stderr=subprocess.DEVNULL).decode())


def compress_cm(m):
    """Minimal Kaldi 'CM ' (kSpeechFeature) writer used only to make a fixture the reference decodes."""
    rows, cols = m.shape
    gmin, gmax = float(m.min()), float(m.max())
    grange = max(gmax - gmin, 1e-6)
    out = struct.pack("<ffii", gmin, grange, rows, cols)

    def to16(v):
        return int(min(65535, max(0, round((v - gmin) / grange * 65535.0))))
    heads, body = b"", b""
    for c in range(cols):
        col = np.sort(m[:, c])
        p = [col[0], col[rows // 4], col[(3 * rows) // 4], col[-1]]
        q = [to16(x) for x in p]
        q[1] = max(q[1], q[0] + 1); q[2] = max(q[2], q[1] + 1); q[3] = max(q[3], q[2] + 1)
        heads += struct.pack("<4H", *q)
        pf = [gmin + grange * x / 65535.0 for x in q]
        v = m[:, c].astype(np.float64)
        b = np.where(v < pf[1], (v - pf[0]) / (pf[1] - pf[0]) * 64.0,
                     np.where(v < pf[2], 64 + (v - pf[1]) / (pf[2] - pf[1]) * 128.0,
                              192 + (v - pf[2]) / (pf[3] - pf[2]) * 63.0))
        body += np.clip(np.round(b), 0, 255).astype(np.uint8).tobytes()
    return out + heads + body


def gen_io(out):
    os.makedirs(out, exist_ok=True)
    rng = np.random.default_rng(77)
    mats = {"uttA": rng.standard_normal((37, 40)).astype(np.float32) * 3 + 1,
            "spk/uttB.wav": rng.standard_normal((301, 80)).astype(np.float32) * 5 - 2,
            "uttC": rng.standard_normal((25, 40)).astype(np.float32)}
    with open(os.path.join(out, "feats_fm.ark"), "wb") as f:
        for k, m in mats.items():
            kaldi_io.write_mat(f, m, key=k)
    with open(os.path.join(out, "feats_cm.ark"), "wb") as f:
        for k, m in mats.items():
            f.write((k + " ").encode() + b"\0BCM " + compress_cm(m))
    dec = {}
    for tag in ("fm", "cm"):
        for k, m in kaldi_io.read_mat_ark(os.path.join(out, "feats_%s.ark" % tag)):
            dec["%s:%s" % (tag, k)] = np.array(m, np.float32)
    np.savez(os.path.join(out, "ref_feats.npz"), **dec)


def gen_net(out):
    os.makedirs(out, exist_ok=True)
    from oracle import net_oracle
    from voxsrc2020_speaker_verification_b200 import arch
    for mid, fd, frames in (("tdnn", 40, 61), ("res2net50_w8_s6_c16", 40, 53), ("dpn68", 40, 50)):
        cfg = arch.get_config(mid)
        params = net_oracle.init_params(cfg, fd, seed=4321)
        x = net_oracle.synth_feats(np.random.default_rng(1234), 2, frames, fd)
        y = net_oracle.forward(cfg, params, x)
        np.savez(os.path.join(out, mid + ".npz"), feat_dim=fd, frames=frames, emb=y)


if __name__ == "__main__":
    g = os.path.join(ROOT, "tests", "golden")
    gen_score(os.path.join(g, "score"))
    gen_io(os.path.join(g, "io"))
    gen_net(os.path.join(g, "net"))
    gen_eer(os.path.join(g, "eer"))
    print("golden fixtures written under", g)
