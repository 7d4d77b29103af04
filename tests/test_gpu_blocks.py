"""Block-by-block parity of the CUDA networks with the oracle.

Whole-network cosine on a damped synthetic net hides two things: a gain error (cosine is scale-blind) and an error inside one
residual branch (damping divides its visibility).  Here every block is checked on its own: the library dumps every op's output
(svx_extractor_set_dump_dir), the test cuts the block INPUT the GPU actually produced out of the tall image, replays exactly that
block in the fp32 oracle (oracle.net_oracle.trace_blocks) and compares with the block OUTPUT the GPU produced — relative L2 and
max-abs, so magnitude counts.  The weights are the UNDAMPED ones (damp=1: every residual branch closes with a unit-variance batch
norm, as a trained reference network does — BN there has no gamma, models.py:64-67).

Bounds: one block is 3-6 convs with 16-bit intermediate activations: ~sqrt(5) * 2^-12 ~ 5e-4 relative L2 in fp16, 8x that in bf16.
"""
import glob
import os
import re

import numpy as np
import pytest
import torch

from oracle import net_oracle
from voxsrc2020_speaker_verification_b200 import arch

pytestmark = pytest.mark.gpu

REL_L2 = {"fp16": 2.5e-3, "bf16": 2.0e-2}
MAX_ABS = {"fp16": 1.0e-2, "bf16": 8.0e-2}      # relative to max |oracle output|


def _ceil_half(n):
    return (n + 1) // 2


def _block_out_ops(cfg):
    """Index (in the library's op list) of the op that writes every block's output, in oracle trace order."""
    if cfg.family == arch.FAMILY_TDNN:
        return list(range(1, 1 + len(cfg.tdnn_filters)))        # op 0 packs the input
    out, idx = [0], 1                                           # op 0: stem
    if cfg.family == arch.FAMILY_RES2NET:
        for li, nblocks in enumerate(cfg.block_sizes):
            for b in range(nblocks):
                stride = cfg.block_strides[li] if b == 0 else 1
                # projection shortcut of a first block: its own op, except in the first stage at stride 1, where it is folded into
                # conv3 as extra K (model.cu build_res2net)
                own_shortcut = b == 0 and not (li == 0 and stride == 1)
                idx += (1 if own_shortcut else 0) + 1 + (cfg.split - 1) + (1 if stride == 2 else 0)
                out.append(idx)                                 # conv3
                idx += 1
    else:
        for si in range(4):
            for b in range(cfg.k_sec[si]):
                idx += 5 if b == 0 else 3
                out.append(idx)                                 # conv_c
                idx += 1
    return out


def _load_dump(path, dtype):
    m = re.search(r"op(\d+)_k(\d+)_t(\d+)_r(\d+)_w(\d+)_c(\d+)_off(\d+)_n(\d+)\.bin$", path)
    idx, kind, tid, rows, wp, C, off, n = (int(g) for g in m.groups())
    raw = np.fromfile(path, dtype=np.uint16).reshape(rows, wp, C)[:, :, off:]      # the op's slice of its destination tensor
    t = torch.from_numpy(np.ascontiguousarray(raw).astype(np.int32)).to(torch.int16).view(torch.float16 if dtype == "fp16" else torch.bfloat16)
    return idx, t.float().numpy()


def _segment(cfg, arr, T, F, channels):
    """[rows, Wp, C] tall image of ONE segment -> NCHW [1, channels, h, W] (the library's layout rule, model.cu layout_segments)."""
    wp = arr.shape[1]
    if cfg.family == arch.FAMILY_TDNN:
        return np.ascontiguousarray(arr[1:1 + T, :1, :channels].transpose(2, 0, 1))[None]
    n_stages = 1 + sum(1 for s in cfg.block_strides if s == 2) if cfg.family == arch.FAMILY_RES2NET else 4
    hs, ws = [T], [F]
    for _ in range(n_stages - 1):
        hs.append(_ceil_half(hs[-1])); ws.append(_ceil_half(ws[-1]))
    offs = [0] * n_stages
    offs[-1] = 1
    for s in range(n_stages - 2, -1, -1):
        offs[s] = 2 * offs[s + 1] - 1 + (hs[s] & 1) if cfg.family == arch.FAMILY_DPN else 2 * offs[s + 1]
    s = ws.index(wp - 1)
    return np.ascontiguousarray(arr[offs[s]:offs[s] + hs[s], :ws[s], :channels].transpose(2, 0, 1))[None]


def _check_blocks(model_id, feat_dim, T, precision, tmp_path, report):
    from voxsrc2020_speaker_verification_b200.extractor import Extractor
    cfg = arch.get_config(model_id)
    params = net_oracle.init_params(cfg, feat_dim, seed=4321, calib_frames=64, calib_batch=4, damp=1.0)
    ex = Extractor(model_id, feat_dim, precision=precision).load_params(params)
    x = net_oracle.synth_feats(np.random.default_rng(5), 1, T, feat_dim)[0]
    d = str(tmp_path / ("dump_%s_%s" % (model_id, precision)))
    os.makedirs(d)
    ex.set_dump_dir(d)
    got_emb = ex.run_segments(torch.from_numpy(x).cuda(), np.array([0, T], np.int32)).cpu().numpy()[0]
    ex.set_dump_dir(None)
    dumps = dict(_load_dump(p, precision) for p in glob.glob(os.path.join(d, "op*.bin")))
    want_emb, trace = net_oracle.trace_blocks(cfg, params, x)
    ops = _block_out_ops(cfg)
    assert len(ops) == len(trace) - 1, (len(ops), len(trace))
    worst = (0.0, "")
    prev = None
    for rec, op in zip(trace[:-1], ops):
        want_shape = rec["out"].shape
        gpu_out = _segment(cfg, dumps[op], T, feat_dim, want_shape[1])
        assert gpu_out.shape == tuple(want_shape), (rec["name"], gpu_out.shape, tuple(want_shape))
        # replay this block on the input the GPU produced (the network input itself for the first block)
        blk_in = rec["in"] if prev is None else torch.from_numpy(prev)
        want = np.asarray(rec["replay"](blk_in))
        rel = float(np.linalg.norm(gpu_out - want) / max(np.linalg.norm(want), 1e-12))
        mab = float(np.abs(gpu_out - want).max() / max(np.abs(want).max(), 1e-12))
        report.append("%-28s %-5s %-16s rel-L2 %.2e  max-abs/max %.2e" % (model_id, precision, rec["name"], rel, mab))
        assert rel <= REL_L2[precision], (model_id, precision, rec["name"], rel)
        assert mab <= MAX_ABS[precision], (model_id, precision, rec["name"], mab)
        worst = max(worst, (rel, rec["name"]))
        prev = gpu_out
    # the tail (pooling + folded BN / dense / BN, fp32 on the GPU) on the GPU's last activation
    want_tail = np.asarray(trace[-1]["replay"](torch.from_numpy(prev)))[0]
    rel_tail = float(np.linalg.norm(got_emb - want_tail) / np.linalg.norm(want_tail))
    assert rel_tail <= 2e-4, (model_id, precision, "tail", rel_tail)
    # whole network, undamped: cosine AND relative L2 (magnitude counts: the chunk rule averages un-normalised embeddings)
    cos = float(np.dot(got_emb, want_emb) / np.linalg.norm(got_emb) / np.linalg.norm(want_emb))
    rel_all = float(np.linalg.norm(got_emb - want_emb) / np.linalg.norm(want_emb))
    report.append("%-28s %-5s %-16s cosine %.6f  rel-L2 %.2e  (undamped weights; worst block %s %.2e)"
                  % (model_id, precision, "whole network", cos, rel_all, worst[1], worst[0]))
    return cos, rel_all


@pytest.mark.parametrize("model_id,feat_dim,T", [("res2net50_w24_s4_c32", 80, 72), ("res2net50_w8_s6_c16", 40, 57),
                                                  ("dpn68", 80, 57), ("dpn68", 40, 64), ("tdnn", 40, 120)])
@pytest.mark.parametrize("precision", ["fp16", "bf16"])
def test_every_block_matches_the_oracle(model_id, feat_dim, T, precision, tmp_path):
    report = []
    try:
        cos, rel = _check_blocks(model_id, feat_dim, T, precision, tmp_path, report)
    finally:
        print("\n".join(report))
        out = os.environ.get("SVX_PARITY_REPORT")
        if out:
            with open(out, "a") as f:
                f.write("\n".join(report) + "\n")
    # Undamped random networks amplify rounding chaotically (measured, profiles/r02_parity_blocks.txt: fp16 0.9964-1.0000, bf16
    # 0.889-0.99997 depending on the network), and no trained checkpoint exists to measure the figure that matters; the
    # whole-network number is REPORTED here and bounded on the damped weights in test_gpu_extract.py.  The tight bounds are the
    # per-block ones above.
    assert np.isfinite(cos) and cos > 0.5, (cos, rel)
