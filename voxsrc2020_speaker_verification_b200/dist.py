"""Multi-GPU layout of the hot path: one process per GPU, torch.distributed (NCCL over NVLink) for the plumbing.

* Extraction shards by utterance — the reference's own layout (eval_inference_model.sh:29-39: static contiguous
  scp shards, one process per GPU, no communication).  ``balance_by_frames`` deals length-sorted utterances so
  every rank gets the same number of frames; ``gather_embeddings`` is the only collective (all-gather of [n,E]).
* AS-norm statistics shard the COHORT by rows (BASELINE.json north_star): every rank scores all test rows against
  its cohort slice and keeps the per-row top-k candidates, one all-gather exchanges the candidate lists, and the
  top-k of their union is reduced to mean/std.  The exchange is the path's only real collective.
The helpers take the process group explicitly so the same code runs under gloo on CPU in the tests.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous near-equal split of n items (first n % world shards get one more), like utils/split_scp.pl:212-243."""
    base, rem = divmod(n, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def balance_by_frames(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Longest-processing-time greedy: utterance indices per rank with near-equal total frames; within a rank
    indices stay sorted by length (descending) so that batches are length-homogeneous."""
    order = np.argsort(-np.asarray(lengths, dtype=np.int64), kind="stable")
    loads = [0] * world
    parts: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(loads))
        parts[r].append(int(i))
        loads[r] += int(lengths[i])
    return parts


def gather_sharded(local: torch.Tensor, parts: Sequence[Sequence[int]], group=None) -> torch.Tensor:
    """Embeddings of a ``balance_by_frames`` partition back in the caller's order on every rank: ONE all-gather of the ranks'
    [n_max, E] blocks (every rank derives the same partition, so neither counts nor indices travel)."""
    world = dist.get_world_size(group)
    n_total = sum(len(p) for p in parts)
    n_max = max(1, max(len(p) for p in parts))
    e = local.shape[1]
    pad = torch.zeros((n_max, e), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    allb = torch.empty((world * n_max, e), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(allb, pad, group=group)
    out = torch.empty((n_total, e), dtype=local.dtype, device=local.device)
    for r, p in enumerate(parts):
        if len(p):
            out[torch.as_tensor(list(p), dtype=torch.int64, device=local.device)] = allb[r * n_max: r * n_max + len(p)]
    return out


def extract_sharded(extract_local, lengths: Sequence[int], group=None) -> torch.Tensor:
    """The extraction stage over the ranks of ``group`` (reference layout: eval_inference_model.sh:29-39, one process per GPU,
    static scp shards, no communication — here the shards are balanced by frame count and the result is gathered).

    ``extract_local(indices)`` → CUDA tensor [len(indices), E]: the embeddings of this rank's utterances, in that order.
    Returns [len(lengths), E] in the caller's order on every rank."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    parts = balance_by_frames(lengths, world)
    local = extract_local(parts[rank])
    return gather_sharded(local, parts, group)


def gather_embeddings(local: torch.Tensor, local_index: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather per-rank embeddings [n_r, E] with their global utterance indices → [n_total, E] on every rank."""
    world = dist.get_world_size(group)
    counts = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device), group=group)
    n_max = int(max(int(c.item()) for c in counts))
    e = local.shape[1]
    pad = torch.zeros((n_max, e), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    idx = torch.full((n_max,), -1, dtype=torch.int64, device=local.device)
    idx[: local.shape[0]] = local_index.to(torch.int64)
    all_e = [torch.empty_like(pad) for _ in range(world)]
    all_i = [torch.empty_like(idx) for _ in range(world)]
    dist.all_gather(all_e, pad, group=group)
    dist.all_gather(all_i, idx, group=group)
    out = torch.zeros((n_total, e), dtype=local.dtype, device=local.device)
    for ee, ii in zip(all_e, all_i):
        keep = ii >= 0
        out[ii[keep]] = ee[keep]
    return out


def merge_candidates(gathered: torch.Tensor) -> torch.Tensor:
    """[world, n, k] per-rank candidate lists → [n, world*k] row-wise union."""
    w, n, k = gathered.shape
    return gathered.permute(1, 0, 2).reshape(n, w * k).contiguous()


def sharded_cohort_mean_std(scorer, test: torch.Tensor, cohort: torch.Tensor, topk: int, group=None):
    """get_cohort_mean_std (snorm.py:83-110) with the cohort row-sharded over the ranks of ``group``.

    ``scorer`` provides cohort_topk_values(test, shard, k) → [n,k] and topk_stats(vals, k) → (mean, std);
    every rank holds the full ``test`` and ``cohort`` (or at least its own slice of the cohort rows).

    Exchange: each rank keeps the per-row top-k candidates of ITS cohort slice for all test rows, then the candidate
    lists are exchanged so that rank r receives, from every rank, the candidates of ITS slice of the test rows
    (all-to-all over NVLink; on backends without all-to-all — gloo in the CPU tests — an all-gather followed by the
    same row slice).  Each rank merges only its rows (top-k of the union of ``world`` lists) and a small all-gather
    returns the [n] statistics to everyone, so neither the merge nor the received bytes grow with the world size."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    c = cohort.shape[0]
    k_eff = min(int(topk), c)
    lo, hi = shard_bounds(c, world)[rank]
    k_shard = max(1, min(k_eff, max(b - a for a, b in shard_bounds(c, world))))
    vals = scorer.cohort_topk_values(test, cohort[lo:hi], k_shard).contiguous()    # [n, k_shard], padded with -1e30
    n, k = vals.shape
    rows = shard_bounds(n, world)
    r_lo, r_hi = rows[rank]
    n_mine = r_hi - r_lo
    if dist.get_backend(group) == "nccl":
        recv = torch.empty((world * n_mine, k), dtype=vals.dtype, device=vals.device)
        dist.all_to_all_single(recv, vals, output_split_sizes=[n_mine] * world, input_split_sizes=[b - a for a, b in rows], group=group)
        mine = recv.view(world, n_mine, k)
    else:
        gathered = torch.empty((world * n, k), dtype=vals.dtype, device=vals.device)
        dist.all_gather_into_tensor(gathered, vals, group=group)
        mine = gathered.view(world, n, k)[:, r_lo:r_hi]
    mean_r, std_r = scorer.topk_stats(merge_candidates(mine.contiguous()), k_eff)
    # statistics of every row back to every rank (row slices differ by at most one row: padded to the longest)
    return _allgather_rows(mean_r, std_r, rows, group)


def _allgather_rows(mean_r: torch.Tensor, std_r: torch.Tensor, rows, group=None):
    world = dist.get_world_size(group)
    n_max = max(b - a for a, b in rows)
    pad = torch.zeros((2, n_max), dtype=mean_r.dtype, device=mean_r.device)
    pad[0, : mean_r.shape[0]] = mean_r
    pad[1, : std_r.shape[0]] = std_r
    out = torch.empty((world, 2, n_max), dtype=pad.dtype, device=pad.device)
    dist.all_gather_into_tensor(out.view(world * 2, n_max), pad, group=group)
    mean = torch.cat([out[r, 0, : b - a] for r, (a, b) in enumerate(rows)])
    std = torch.cat([out[r, 1, : b - a] for r, (a, b) in enumerate(rows)])
    return mean, std


def rows_sharded_cohort_mean_std(scorer, test: torch.Tensor, cohort: torch.Tensor, topk: int, group=None):
    """The natural layout: the TEST rows are sharded, the 6 MB cohort is replicated; every rank computes
    the statistics of its own rows (``scorer.cohort_mean_std``) and one all-gather of 2 floats per row returns them to
    everyone.  No candidate exchange at all — kept beside the cohort-sharded layout the north star names."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    rows = shard_bounds(test.shape[0], world)
    lo, hi = rows[rank]
    mean_r, std_r = scorer.cohort_mean_std(test[lo:hi], cohort, topk)
    return _allgather_rows(mean_r, std_r, rows, group)
