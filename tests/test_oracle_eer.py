"""The EER / minDCF oracle against what the REFERENCE's eer_minDCF.py computed and printed for tests/golden/eer
(generated in the build container by oracle/gen_golden.py:gen_eer, which imports /root/reference/tensorflow/eer_minDCF.py)."""
import os

import numpy as np
import pytest

from oracle import eer_oracle


def test_matches_reference_numbers(golden_dir):
    g = os.path.join(golden_dir, "eer")
    ref = np.load(os.path.join(g, "ref_eer.npy"))
    for c_miss, c_fa, p_target, eer, eer_thr, dcf, dcf_thr in ref:
        got = eer_oracle.score_file_metrics(os.path.join(g, "trials.txt"), os.path.join(g, "scores.txt"), c_miss, c_fa, p_target)
        np.testing.assert_allclose(got, [eer, eer_thr, dcf, dcf_thr], rtol=0, atol=1e-12)


def test_matches_reference_stdout(golden_dir):
    g = os.path.join(golden_dir, "eer")
    eer, eer_thr, dcf, dcf_thr = eer_oracle.score_file_metrics(os.path.join(g, "trials.txt"), os.path.join(g, "scores.txt"))
    lines = ["EER is {:.4f}%, at threshold: {:.4f}".format(eer * 100, eer_thr),
             "minDCF is {:.4f}, at threshold: {:.4f} (p-target={}, c-miss={}, c-fa={})".format(dcf, dcf_thr, 0.01, 1.0, 1.0)]
    want = open(os.path.join(g, "ref_stdout.txt")).read().strip().splitlines()
    assert lines[0] == want[0]
    assert lines[1].split(" (")[0] == want[1].split(" (")[0]


def test_roc_equals_sklearn():
    sk = pytest.importorskip("sklearn.metrics")
    rng = np.random.default_rng(3)
    for n in (2, 3, 50, 4000):
        y = (rng.random(n) < 0.3).astype(int)
        y[0], y[-1] = 0, 1
        s = np.round(rng.normal(y * 0.5, 0.4), 2).astype(np.float32)       # rounded: many ties
        f1, t1, h1 = eer_oracle.roc_curve(y, s)
        f0, t0, h0 = sk.roc_curve(y, s.astype(np.float64), pos_label=1)
        np.testing.assert_array_equal(f1, f0)
        np.testing.assert_array_equal(t1, t0)
        np.testing.assert_array_equal(h1[1:], h0[1:])
