"""torchrun check: cohort-sharded AS-norm statistics (all-to-all exchange) == single-GPU statistics, plus timing."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, ".")
from voxsrc2020_speaker_verification_b200 import dist as svdist
from voxsrc2020_speaker_verification_b200.scoring import Scorer
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(5)
N, C, D, K = 145161, 5994, 256, 300
sc = Scorer(local)
x = sc.l2norm(torch.from_numpy(rng.standard_normal((N, D), dtype=np.float32)).cuda())
cohort = sc.l2norm(torch.from_numpy(rng.standard_normal((C, D), dtype=np.float32)).cuda()) * 0.6
m0, s0 = sc.cohort_mean_std(x, cohort, K)
m1, s1 = svdist.sharded_cohort_mean_std(sc, x, cohort, K)
err = max(float((m0 - m1).abs().max()), float((s0 - s1).abs().max()))
for _ in range(2):
    svdist.sharded_cohort_mean_std(sc, x, cohort, K)
torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
for _ in range(5):
    svdist.sharded_cohort_mean_std(sc, x, cohort, K)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    sc.cohort_mean_std(x, cohort, K)
torch.cuda.synchronize(); dt1 = (time.perf_counter() - t0) / 5
if rank == 0:
    print("world %d: max |sharded - single| = %.2e; sharded %.2f ms, single GPU %.2f ms" % (world, err, dt * 1e3, dt1 * 1e3))
assert err < 2e-6, err
m2, s2 = svdist.rows_sharded_cohort_mean_std(sc, x, cohort, K)
err2 = max(float((m0 - m2).abs().max()), float((s0 - s2).abs().max()))
assert err2 < 2e-6, err2
if rank == 0:
    print("OK rows-sharded max |diff| = %.2e" % err2)
dist.destroy_process_group()
