"""Multi-GPU host logic on CPU: world_size 2 over gloo.  The CUDA scorer is replaced by a NumPy stand-in with the
same two methods, so what is tested is the sharding, the all-gather and the merge."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from voxsrc2020_speaker_verification_b200 import dist as svdist


class NumpyScorer:
    def cohort_topk_values(self, test, shard, k):
        s = test.numpy() @ shard.numpy().T
        top = -np.sort(-s, axis=1)[:, :k]
        if top.shape[1] < k:
            top = np.concatenate([top, np.full((top.shape[0], k - top.shape[1]), -1e30, np.float32)], 1)
        return torch.from_numpy(np.ascontiguousarray(top.astype(np.float32)))

    def cohort_mean_std(self, test, cohort, k):
        top = -np.sort(-(test.numpy() @ cohort.numpy().T), axis=1)[:, :k]
        return torch.from_numpy(top.mean(1)), torch.from_numpy(top.std(1))

    def topk_stats(self, vals, k):
        top = -np.sort(-vals.numpy(), axis=1)[:, :k]
        return torch.from_numpy(top.mean(1)), torch.from_numpy(top.std(1))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.standard_normal((51, 16)).astype(np.float32))
    cohort = torch.from_numpy(rng.standard_normal((37, 16)).astype(np.float32))
    mean, std = svdist.sharded_cohort_mean_std(NumpyScorer(), x, cohort, 10)
    mean2, std2 = svdist.rows_sharded_cohort_mean_std(NumpyScorer(), x, cohort, 10)
    assert torch.allclose(mean, mean2, atol=1e-6) and torch.allclose(std, std2, atol=1e-6)
    # extraction side: every rank "extracts" its balanced share, then the embeddings are gathered
    lengths = [300, 25, 2999, 1000, 57, 640, 41]
    mine = svdist.balance_by_frames(lengths, world)[rank]
    local = torch.tensor([[float(i), float(lengths[i])] for i in mine])
    allv = svdist.gather_embeddings(local, torch.tensor(mine), len(lengths))
    # the same through extract_sharded: one all-gather, the partition derived on every rank
    allv2 = svdist.extract_sharded(lambda idx: torch.tensor([[float(i), float(lengths[i])] for i in idx]).reshape(len(idx), 2), lengths)
    assert torch.equal(allv, allv2)
    if rank == 0:
        q.put((mean.numpy(), std.numpy(), allv.numpy()))
    dist.destroy_process_group()


def test_world2_sharded_cohort_and_gather():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    mean, std, allv = q.get(timeout=120)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((51, 16)).astype(np.float32)
    cohort = rng.standard_normal((37, 16)).astype(np.float32)
    top = -np.sort(-(x @ cohort.T), axis=1)[:, :10]
    np.testing.assert_allclose(mean, top.mean(1), atol=1e-6)
    np.testing.assert_allclose(std, top.std(1), atol=1e-6)
    lengths = [300, 25, 2999, 1000, 57, 640, 41]
    np.testing.assert_array_equal(allv, np.array([[i, l] for i, l in enumerate(lengths)], np.float32))


def test_shard_bounds_and_balance():
    assert svdist.shard_bounds(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert svdist.shard_bounds(5994, 8)[0] == (0, 750) and svdist.shard_bounds(5994, 8)[-1] == (5245, 5994)
    rng = np.random.default_rng(1)
    lengths = rng.integers(300, 3001, 1000).tolist()
    parts = svdist.balance_by_frames(lengths, 8)
    assert sorted(i for p in parts for i in p) == list(range(1000))
    loads = [sum(lengths[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= 3000
    g = torch.arange(24, dtype=torch.float32).reshape(2, 3, 4)
    m = svdist.merge_candidates(g)
    assert m.shape == (3, 8) and m[1].tolist() == [4, 5, 6, 7, 16, 17, 18, 19]
