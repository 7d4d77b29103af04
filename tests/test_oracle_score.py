"""The scoring oracle (oracle/score_oracle.py) against the fixtures produced by the reference's own snorm.py."""
import os

import numpy as np

from oracle import score_oracle
from voxsrc2020_speaker_verification_b200 import kaldi_ark


def _load(g):
    keys, mat = kaldi_ark.read_vec_ark_matrix(os.path.join(g, "test.ark"))
    xv = score_oracle.normalise_xvectors(dict(zip(keys, mat)))
    ckeys, cmat = kaldi_ark.read_vec_ark_matrix(os.path.join(g, "cohort.ark"))
    cx = score_oracle.normalise_xvectors(dict(zip(ckeys, cmat)))
    cohort = score_oracle.read_speaker_xvector(cx, score_oracle.read_spk2utt(os.path.join(g, "cohort_spk2utt")))
    return xv, cohort


def test_cohort_matches_reference(golden_dir):
    g = os.path.join(golden_dir, "score")
    _, cohort = _load(g)
    ref = np.load(os.path.join(g, "ref_cohort.npz"))
    assert list(cohort.keys()) == list(ref["keys"])
    np.testing.assert_array_equal(np.array(list(cohort.values())), ref["matrix"])


def test_mean_std_matches_reference(golden_dir):
    g = os.path.join(golden_dir, "score")
    xv, cohort = _load(g)
    for topk in (50, 300, 400):
        ref = np.load(os.path.join(g, "ref_topk%d.npz" % topk))
        m, s = score_oracle.get_cohort_mean_std(xv, cohort, topk)
        assert list(m.keys()) == list(ref["keys"])
        np.testing.assert_array_equal(np.array(list(m.values()), np.float32), ref["mean"])
        np.testing.assert_array_equal(np.array(list(s.values()), np.float32), ref["std"])
        am, as_ = score_oracle.cohort_mean_std_arrays(np.array(list(xv.values())), np.array(list(cohort.values())), topk)
        np.testing.assert_array_equal(am, ref["mean"])
        np.testing.assert_array_equal(as_, ref["std"])


def test_score_files_match_reference_bytes(golden_dir, tmp_path):
    g = os.path.join(golden_dir, "score")
    xv, cohort = _load(g)
    cos = score_oracle.get_cosine_score(xv, os.path.join(g, "trials.txt"))
    m, s = score_oracle.get_cohort_mean_std(xv, cohort)        # default topk = 400, as snorm.py:176
    sn = score_oracle.get_asnorm1_score(m, s, cos)
    score_oracle.write_scores(str(tmp_path / "c.txt"), cos)
    score_oracle.write_scores(str(tmp_path / "s.txt"), sn)
    assert open(tmp_path / "c.txt").read() == open(os.path.join(g, "ref_cosine.txt")).read()
    assert open(tmp_path / "s.txt").read() == open(os.path.join(g, "ref_snorm_topk400.txt")).read()


def test_array_forms_agree_with_dict_forms(golden_dir):
    g = os.path.join(golden_dir, "score")
    xv, cohort = _load(g)
    keys = list(xv.keys())
    index = {k: i for i, k in enumerate(keys)}
    pairs = score_oracle.parse_trials(os.path.join(g, "trials.txt"))
    x = np.array(list(xv.values()))
    m, s = score_oracle.cohort_mean_std_arrays(x, np.array(list(cohort.values())), 300)
    i1 = np.array([index[a] for a, _ in pairs]); i2 = np.array([index[b] for _, b in pairs])
    cos, sn = score_oracle.trial_scores_arrays(x, i1, i2, m, s)
    ref = score_oracle.get_cosine_score(xv, os.path.join(g, "trials.txt"))
    np.testing.assert_allclose(cos, [r[2] for r in ref], atol=2e-7)
