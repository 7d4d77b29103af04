"""AS-norm statistics: fused vs unfused kernels on BASELINE config 5 (145 160 x 256 vs 5994, top-300) — correctness against the
oracle on a row sample, fall-back counts, device time per job.  Usage: python tools/score_bench.py [--n N --c C --d D --topk K]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import score_oracle
from voxsrc2020_speaker_verification_b200.scoring import Scorer

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=145160)
ap.add_argument("--c", type=int, default=5994)
ap.add_argument("--d", type=int, default=256)
ap.add_argument("--topk", type=int, default=300)
ap.add_argument("--trials", type=int, default=579818)
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
rng = np.random.default_rng(99)


def unit(n, d):
    x = rng.standard_normal((n, d)).astype(np.float32)
    return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)


x = unit(a.n, a.d)
cohort = ((unit(a.c, a.d) + unit(a.c, a.d) + unit(a.c, a.d)) / np.float32(3)).astype(np.float32)
xt, ct = torch.from_numpy(x).cuda(), torch.from_numpy(cohort).cuda()
i1 = torch.from_numpy(rng.integers(0, a.n, a.trials).astype(np.int32)).cuda()
i2 = torch.from_numpy(rng.integers(0, a.n, a.trials).astype(np.int32)).cuda()
sc = Scorer(0)
rows = rng.choice(a.n, min(a.n, 1024), replace=False)
wm, ws = score_oracle.cohort_mean_std_arrays(x[rows], cohort, a.topk)
for fused in (1, 0):
    sc.set_option("fused", fused)
    mean, std = sc.cohort_mean_std(xt, ct, a.topk)
    torch.cuda.synchronize()
    path = sc.last_path()
    em = float(np.abs(mean.cpu().numpy()[rows] - wm).max())
    es = float(np.abs(std.cpu().numpy()[rows] - ws).max())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(10):          # the oracle above left the GPU idle for seconds: let the clocks come back before timing
        sc.cohort_mean_std(xt, ct, a.topk)
    e0.record()
    for _ in range(a.reps):
        mean, std = sc.cohort_mean_std(xt, ct, a.topk)
    e1.record()
    torch.cuda.synchronize()
    ms_stats = e0.elapsed_time(e1) / a.reps
    e0.record()
    for _ in range(a.reps):
        sc.trial_scores(xt, i1, i2, mean, std)
    e1.record()
    torch.cuda.synchronize()
    ms_tr = e0.elapsed_time(e1) / a.reps
    print("fused=%d: stats %.3f ms  trials %.3f ms  job %.3f ms = %.1f M trials/s | max |d mean| %.2e |d std| %.2e | fused rows %d, handed back %d"
          % (fused, ms_stats, ms_tr, ms_stats + ms_tr, a.trials / (ms_stats + ms_tr) / 1e3, em, es, path[0], path[1]), flush=True)
