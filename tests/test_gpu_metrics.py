"""EER / minDCF: the device implementation (svx_eer_min_dcf) against the numbers and the printed lines of the
REFERENCE's eer_minDCF.py (tests/golden/eer), and the north star's end-to-end criterion — EER and minDCF of the score files the
GPU scoring stage writes agree with those of the score files the reference's snorm.py wrote to within 0.01 absolute."""
import os

import numpy as np
import pytest

from oracle import eer_oracle

pytestmark = pytest.mark.gpu


def test_device_eer_equals_reference(golden_dir):
    from voxsrc2020_speaker_verification_b200 import eer_minDCF
    g = os.path.join(golden_dir, "eer")
    pair_label = eer_minDCF.read_trial_file(os.path.join(g, "trials.txt"))
    pair_score = eer_minDCF.read_score_file(os.path.join(g, "scores.txt"))
    y = [pair_label[p] for p in pair_label]
    y_pred = [pair_score[p] for p in pair_label]
    for c_miss, c_fa, p_target, eer, eer_thr, dcf, dcf_thr in np.load(os.path.join(g, "ref_eer.npy")):
        got = eer_minDCF.compute_eer_and_min_dcf(y, y_pred, c_miss, c_fa, p_target)
        np.testing.assert_allclose([got[0], got[2]], [eer, dcf], rtol=0, atol=1e-12)
        # thresholds: the reference holds float(text) of a float32, the device the float32 itself — the same score to 2^-24
        np.testing.assert_allclose([got[1], got[3]], [eer_thr, dcf_thr], rtol=6e-8, atol=0)


def test_cli_prints_what_the_reference_prints(golden_dir, capsys):
    from voxsrc2020_speaker_verification_b200 import eer_minDCF
    g = os.path.join(golden_dir, "eer")
    assert eer_minDCF.main(["--trial", os.path.join(g, "trials.txt"), "--score", os.path.join(g, "scores.txt")]) == 0
    got = capsys.readouterr().out.strip().splitlines()
    want = open(os.path.join(g, "ref_stdout.txt")).read().strip().splitlines()
    assert got[0] == want[0]
    assert got[1].split(" (")[0] == want[1].split(" (")[0]          # the tail echoes argparse floats (1.0 vs 1): same numbers


@pytest.mark.parametrize("n", [2, 3, 1000, 300000])
def test_device_eer_vs_oracle_random(n):
    from voxsrc2020_speaker_verification_b200 import eer_minDCF
    rng = np.random.default_rng(n)
    y = (rng.random(n) < 0.2).astype(np.int32)
    y[0], y[-1] = 0, 1
    s = rng.normal(y * 0.4, 0.3).astype(np.float32)
    if n >= 1000:
        s[: n // 3] = np.round(s[: n // 3], 2)                      # ties
    got = eer_minDCF.compute_eer_and_min_dcf(y, s, 1.0, 1.0, 0.01)
    want = eer_oracle.compute_eer_and_min_dcf(y, s, 1.0, 1.0, 0.01)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)           # same float32 scores on both sides: everything exact


def _fields(path):
    return [ln.rstrip("\n").split(" ") for ln in open(path)]


def test_score_files_give_the_reference_eer_and_have_its_text_form(golden_dir, tmp_path):
    """snorm drop-in on the golden arks -> score files; eer_minDCF on them vs on the files the reference's snorm.py wrote:
    |dEER| and |dminDCF| < 0.01 absolute (north star), same trial keys line by line, and every score printed the way
    print(np.float32) prints it (shortest repr, snorm.py:166,182)."""
    from voxsrc2020_speaker_verification_b200 import snorm
    g = os.path.join(golden_dir, "score")
    cos_out, sn_out = str(tmp_path / "cosine.txt"), str(tmp_path / "snorm.txt")
    snorm.main(["--trial", os.path.join(g, "trials.txt"), "--test_ark", os.path.join(g, "test.ark"), "--cosine_score", cos_out,
                "--cohort_ark", os.path.join(g, "cohort.ark"), "--cohort_spk2utt", os.path.join(g, "cohort_spk2utt"),
                "--snorm_score", sn_out])
    for got_p, ref_p in ((cos_out, "ref_cosine.txt"), (sn_out, "ref_snorm_topk400.txt")):
        a = eer_oracle.score_file_metrics(os.path.join(g, "trials.txt"), got_p)
        b = eer_oracle.score_file_metrics(os.path.join(g, "trials.txt"), os.path.join(g, ref_p))
        assert abs(a[0] - b[0]) * 100 < 0.01 and abs(a[2] - b[2]) < 0.01, (a, b)
        got, ref = _fields(got_p), _fields(os.path.join(g, ref_p))
        assert len(got) == len(ref)
        for (a1, a2, s1), (b1, b2, s2) in zip(got, ref):
            assert (a1, a2) == (b1, b2)
            assert s1 == str(np.float32(float(s1)))                  # the reference's text form of a float32
            # the text itself differs in the last digits for most lines: np.dot's BLAS summation order is not the kernel's
            assert abs(float(s1) - float(s2)) <= 2e-4 * max(1.0, abs(float(s2)))
