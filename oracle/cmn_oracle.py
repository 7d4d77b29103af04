"""ORACLE (test infrastructure, not product code): CPU restatement of Kaldi's sliding-window cepstral mean normalisation,
the pipe `apply-cmvn-sliding --norm-vars=false --center=true --cmn-window=300` that the reference reads its features through
(reference tensorflow/tf_extract.py:63).

The algorithm lives in a third-party dependency that is absent from /root/reference: Kaldi (`src/feat/feature-functions.cc`,
`SlidingWindowCmn`), cloned at an unpinned git HEAD by the reference's Dockerfile (`Dockerfile:42`).  Its published window rule
[ext], restated here frame by frame:

    center:      window_start = t - cmn_window/2 ; window_end = window_start + cmn_window
    otherwise:   window_start = t - cmn_window   ; window_end = t + 1
    if window_start < 0:   window_end -= window_start ; window_start = 0
    if not center and window_end > t:   window_end = max(t + 1, min_window)          (--min-cmn-window, default 100)
    if window_end > T:     window_start -= window_end - T ; window_end = T ; window_start = max(window_start, 0)
    out[t] = x[t] - mean(x[window_start:window_end])                                   (sums in double precision)

Parity status: the rule is pinned by hand-computed vectors (tests/test_oracle_cmn.py), not by a Kaldi binary (none exists
in this image) — "parity unpinned" against Kaldi itself.  Only tests/, __graft_entry__.smoke() and bench.py may import this.
"""
from __future__ import annotations

import numpy as np


def window(t: int, T: int, cmn_window: int = 300, center: bool = True, min_window: int = 100):
    """[start, end) of the frames whose mean is subtracted from frame t of a T-frame utterance."""
    if center:
        ws = t - cmn_window // 2
        we = ws + cmn_window
    else:
        ws = t - cmn_window
        we = t + 1
    if ws < 0:
        we -= ws
        ws = 0
    if not center and we > t:
        we = max(t + 1, min_window)
    if we > T:
        ws -= we - T
        we = T
        ws = max(ws, 0)
    return ws, we


def apply_cmvn_sliding_naive(feats: np.ndarray, cmn_window: int = 300, center: bool = True, min_window: int = 100) -> np.ndarray:
    """Per-frame loop, the definition."""
    T = feats.shape[0]
    out = np.empty_like(feats, dtype=np.float32)
    x = feats.astype(np.float64)
    for t in range(T):
        ws, we = window(t, T, cmn_window, center, min_window)
        out[t] = (x[t] - x[ws:we].sum(0) / (we - ws)).astype(np.float32)
    return out


def apply_cmvn_sliding(feats: np.ndarray, cmn_window: int = 300, center: bool = True, min_window: int = 100) -> np.ndarray:
    """The same through running sums (what Kaldi and the CUDA kernels do), for sizes the naive loop is slow on."""
    T = feats.shape[0]
    csum = np.concatenate([np.zeros((1, feats.shape[1]), np.float64), np.cumsum(feats.astype(np.float64), axis=0)])
    se = np.array([window(t, T, cmn_window, center, min_window) for t in range(T)], dtype=np.int64).reshape(T, 2)
    ws, we = se[:, 0], se[:, 1]
    mean = (csum[we] - csum[ws]) / (we - ws)[:, None]
    return (feats.astype(np.float64) - mean).astype(np.float32)
