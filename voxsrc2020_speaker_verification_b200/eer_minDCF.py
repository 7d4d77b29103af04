"""Drop-in for the reference's ``tensorflow/eer_minDCF.py`` with the metric computed on the GPU.

    python -m voxsrc2020_speaker_verification_b200.eer_minDCF --trial L --score OUT [--c-miss 1 --c-fa 1 --p-target 0.01]

Same flags, same two printed lines (eer_minDCF.py:68-94).  ``compute_eer_and_min_dcf`` keeps the reference's name and
signature (eer_minDCF.py:41).  The reference script itself keeps working on our score files unchanged; this module only
saves the sklearn ROC + Python loop over ~580 k thresholds when sweeping many systems.
"""
from __future__ import annotations

import argparse
import ctypes
import sys

import numpy as np
import torch

from . import lib


def read_score_file(score_file):
    """eer_minDCF.py:23-29 — a repeated pair keeps its last score."""
    pair_score = {}
    with open(score_file, "r") as f:
        for line in f:
            utt1, utt2, score = line.strip().split()
            pair_score[(utt1, utt2)] = float(score)
    return pair_score


def read_trial_file(trial_file):
    """eer_minDCF.py:32-38."""
    pair_label = {}
    with open(trial_file, "r") as f:
        for line in f:
            label, utt1, utt2 = line.strip().split()
            pair_label[(utt1, utt2)] = int(label)
    return pair_label


def compute_eer_and_min_dcf(y, y_pred, c_miss, c_fa, p_target, device: int = 0):
    """eer_minDCF.py:41-64 → (eer, eer_threshold, min_dcf, min_c_det_threshold).  Scores are compared as float32 (what the
    score files hold)."""
    dv = torch.device("cuda", device)
    s = torch.as_tensor(np.ascontiguousarray(np.asarray(y_pred, dtype=np.float32))).to(dv)
    l = torch.as_tensor(np.ascontiguousarray(np.asarray(y, dtype=np.int32))).to(dv)
    out = (ctypes.c_double * 4)()
    lib.check(lib.load().svx_eer_min_dcf(ctypes.c_void_p(s.data_ptr()), ctypes.c_void_p(l.data_ptr()), s.shape[0], float(c_miss), float(c_fa),
                                         float(p_target), ctypes.cast(out, ctypes.c_void_p),
                                         ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)))
    return out[0], out[1], out[2], out[3]


def main(argv=None) -> int:
    parser = argparse.ArgumentParser()
    parser.add_argument("--c-miss", type=float, dest="c_miss", default=1, help="Cost of a missed detection.  This is usually not changed.")
    parser.add_argument("--c-fa", type=float, dest="c_fa", default=1, help="Cost of a spurious detection.  This is usually not changed.")
    parser.add_argument("--p-target", type=float, dest="p_target", default=0.01, help="The prior probability of the target speaker in a trial.")
    parser.add_argument("--trial", type=str, help="the trial file")
    parser.add_argument("--score", type=str, help="the score file")
    args = parser.parse_args(argv)
    pair_label = read_trial_file(args.trial)
    pair_score = read_score_file(args.score)
    y, y_pred = [], []
    for pair in pair_label.keys():
        y.append(pair_label[pair])
        y_pred.append(pair_score[pair])          # KeyError if a trial has no score, as in the reference
    eer, eer_threshold, min_dcf, min_c_det_threshold = compute_eer_and_min_dcf(y, y_pred, args.c_miss, args.c_fa, args.p_target)
    print("EER is {:.4f}%, at threshold: {:.4f}".format(eer * 100, eer_threshold))
    print("minDCF is {:.4f}, at threshold: {:.4f} (p-target={}, c-miss={}, c-fa={})".format(min_dcf, min_c_det_threshold, args.p_target,
                                                                                              args.c_miss, args.c_fa))
    return 0


if __name__ == "__main__":
    sys.exit(main())
