"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's EER / minDCF computation
(reference tensorflow/eer_minDCF.py:23-64).

The reference builds the ROC with a third-party call, ``sklearn.metrics.roc_curve(y, y_pred, pos_label=1)`` (scikit-learn is
unpinned in the reference's Dockerfile; 1.9.0 in this image).  Its published algorithm [ext], restated here in NumPy:

  * sort the scores descending; a threshold exists where the score changes (and at the last element);
  * tps = cumulative positives at those indices, fps = 1 + index - tps;
  * drop_intermediate=True: keep a point only if it is the first, the last, or a corner (second difference of fps or tps non-zero);
  * prepend the origin (tps = fps = 0) with threshold +inf;  fpr = fps / fps[-1], tpr = tps / tps[-1].

Pinned: tests/test_oracle_eer.py checks this file against the numbers and the printed lines the reference's own eer_minDCF.py
produced for tests/golden/eer (oracle/gen_golden.py gen_eer), and the ROC against scikit-learn's where it is installed.
Only tests/, __graft_entry__.smoke() and bench.py may import this.
"""
from __future__ import annotations

import numpy as np


def roc_curve(y, y_pred):
    y = np.asarray(y).astype(np.float64)
    s = np.asarray(y_pred, dtype=np.float64)
    order = np.argsort(s, kind="mergesort")[::-1]
    s, y = s[order], y[order]
    distinct = np.where(np.diff(s))[0]
    idx = np.r_[distinct, y.size - 1]
    tps = np.cumsum(y)[idx]
    fps = 1 + idx - tps
    thr = s[idx]
    if len(fps) > 2:
        keep = np.where(np.r_[True, np.logical_or(np.diff(fps, 2), np.diff(tps, 2)), True])[0]
        fps, tps, thr = fps[keep], tps[keep], thr[keep]
    tps, fps, thr = np.r_[0, tps], np.r_[0, fps], np.r_[np.inf, thr]
    return fps / fps[-1], tps / tps[-1], thr


def read_score_file(score_file):
    """eer_minDCF.py:23-29 — keyed by (utt1, utt2): a repeated pair keeps its last score."""
    out = {}
    for line in open(score_file):
        a, b, s = line.strip().split()
        out[(a, b)] = float(s)
    return out


def read_trial_file(trial_file):
    """eer_minDCF.py:32-38."""
    out = {}
    for line in open(trial_file):
        l, a, b = line.strip().split()
        out[(a, b)] = int(l)
    return out


def compute_eer_and_min_dcf(y, y_pred, c_miss=1.0, c_fa=1.0, p_target=0.01):
    """eer_minDCF.py:41-64 → (eer, eer_threshold, min_dcf, min_dcf_threshold)."""
    fprs, tprs, thr = roc_curve(y, y_pred)
    fnrs = 1.0 - tprs
    i = int(np.nanargmin(np.absolute(fnrs - fprs)))
    eer, eer_thr = fprs[i], thr[i]
    c_det = c_miss * fnrs * p_target + c_fa * fprs * (1 - p_target)
    j = int(np.argmin(c_det))                    # first minimum, like the reference's strict `<` scan
    c_def = min(c_miss * p_target, c_fa * (1 - p_target))
    return float(eer), float(eer_thr), float(c_det[j] / c_def), float(thr[j])


def score_file_metrics(trial_file, score_file, c_miss=1.0, c_fa=1.0, p_target=0.01):
    """eer_minDCF.py:68-94 without the printing."""
    pair_label, pair_score = read_trial_file(trial_file), read_score_file(score_file)
    y = [pair_label[p] for p in pair_label]
    y_pred = [pair_score[p] for p in pair_label]
    return compute_eer_and_min_dcf(y, y_pred, c_miss, c_fa, p_target)
