"""One AS-norm job of BASELINE config 5 for ncu captures / timing."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voxsrc2020_speaker_verification_b200.scoring import Scorer
N, C, D, K, T = 145160, 5994, 256, 300, 579818
rng = np.random.default_rng(99)
sc = Scorer(0)
x = sc.l2norm(torch.from_numpy(rng.standard_normal((N, D), dtype=np.float32)).cuda())
c3 = torch.from_numpy(rng.standard_normal((3, C, D), dtype=np.float32)).cuda()
cohort = (sc.l2norm(c3[0]) + sc.l2norm(c3[1]) + sc.l2norm(c3[2])) / 3.0
i1 = torch.from_numpy(rng.integers(0, N, T).astype(np.int32)).cuda()
i2 = torch.from_numpy(rng.integers(0, N, T).astype(np.int32)).cuda()
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    mean, std = sc.cohort_mean_std(x, cohort, K)
    sc.trial_scores(x, i1, i2, mean, std)
    torch.cuda.synchronize(); print("job ms", (time.perf_counter() - t0) * 1e3)
