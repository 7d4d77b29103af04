"""Drop-in for the reference's ``tensorflow/snorm.py`` stage: same flags, same score files.

    python -m voxsrc2020_speaker_verification_b200.snorm --trial L --test_ark A --cosine_score OUT1 \
        [--test_spk2utt S] [--cohort_ark B --cohort_spk2utt U | --weight_matrix W.pkl] [--snorm_score OUT2]

The importable functions of the reference module are re-exported with the same names.
"""
from __future__ import annotations

import argparse
import sys

from .scoring import (get_asnorm1_score, get_cohort_mean_std, get_cohort_xvector, get_cosine_score,  # noqa: F401
                      get_projection_weight, l2norm, read_speaker_xvector, read_spk2utt, read_xvector, score_files)


def main(argv=None) -> int:
    p = argparse.ArgumentParser()
    p.add_argument("--test_ark", type=str, help="the ark file of test xvectors")
    p.add_argument("--test_spk2utt", type=str, default=None, help="the spk2utt file of test speakers")
    p.add_argument("--trial", type=str, help="the trial file")
    p.add_argument("--cosine_score", type=str, help="cosine score file")
    p.add_argument("--cohort_ark", type=str, default=None, required=False, help="the ark file of cohort xvectors")
    p.add_argument("--cohort_spk2utt", type=str, default=None, required=False, help="the spk2utt file of cohort xvectors")
    p.add_argument("--weight_matrix", type=str, default=None, required=False, help="the projection weight matrix file")
    p.add_argument("--snorm_score", type=str, default=None, required=False, help="snorm score file")
    p.add_argument("--topk", type=int, default=400, help="cohort size of the adaptive norm (reference default: 400, snorm.py:83)")
    a = p.parse_args(argv)
    score_files(a.trial, a.test_ark, a.cosine_score, a.test_spk2utt, a.cohort_ark, a.cohort_spk2utt, a.weight_matrix,
                a.snorm_score, a.topk)
    return 0


if __name__ == "__main__":
    sys.exit(main())
