"""ORACLE (test infrastructure, not product code): CPU fp32 restatement of the reference forward graphs.

PARITY UNPINNED for the network: the reference ships no golden vectors, no weights and no tests for
this path, and TensorFlow 1.x cannot be imported here.  This file restates, function
by function, what the reference's graph builders ask TensorFlow to compute, using PyTorch CPU fp32
ops whose semantics are themselves unit-tested against naive NumPy loops in tests/test_oracle_net.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.

Tensors are kept NCHW internally (H = time, W = feature axis) but every comment speaks the
reference's NHWC.  ``params`` maps TF variable names (see voxsrc2020_speaker_verification_b200/arch.py)
to float32 numpy arrays in TF layouts (kernels HWIO).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

from voxsrc2020_speaker_verification_b200 import arch
from voxsrc2020_speaker_verification_b200.arch import (BN_EPS_2D, BN_EPS_4D, FAMILY_DPN, FAMILY_RES2NET,
                                                         FAMILY_TDNN, POOL_EPS, ModelConfig)


class _Ctx:
    """Carries the parameter dict, the TF scope counters and the optional calibration mode."""

    def __init__(self, params: Dict[str, np.ndarray], calibrate: bool = False, dtype=torch.float32,
                 quant=None, damp: float = 1.0):
        self.params = params
        self.calibrate = calibrate
        self.damp = damp    # calibration only: residual-branch damping of the synthetic weights
        self.dtype = dtype
        self.root = arch._Namer()
        self.quant = quant  # optional callable emulating bf16 storage (numerics study only)
        self.trace = None   # optional list: one record per block (input, output, a replay closure), see trace_blocks()

    def record(self, name, x_in, x_out, fn):
        """Block-level tap for the per-block parity tests: ``fn(x)`` re-runs exactly this block (same variables) on another
        input.  The namer state at block entry is captured so that the replay resolves the same TF variable names."""
        if self.trace is not None:
            self.trace.append({"name": name, "in": x_in, "out": x_out, "replay": fn})

    def tensor(self, name: str) -> torch.Tensor:
        return torch.from_numpy(np.asarray(self.params[name])).to(self.dtype)


def tf_same_pad(n: int, k: int, s: int, d: int = 1):
    """TF 'SAME' padding for one axis → (out, pad_before, pad_after) [ext: tf.nn.conv2d]."""
    out = -(-n // s)
    eff = (k - 1) * d + 1
    total = max((out - 1) * s + eff - n, 0)
    return out, total // 2, total - total // 2


def conv2d(ctx: _Ctx, namer, x, filters, kernel_size, strides=1, padding="valid", dilation=(1, 1), cardinality=1,
           branch_end: bool = False):
    """reference models.py:173-203 — bias-free NHWC conv, HWIO kernel [kh,kw,Cin/g,Cout]."""
    scope = namer.next("conv2d")
    if ctx.calibrate and branch_end and ctx.damp != 1.0:
        ctx.params[scope + "/kernel"] = (ctx.params[scope + "/kernel"] / np.float32(ctx.damp)).astype(np.float32)
    k = ctx.tensor(scope + "/kernel")
    return _conv_raw(x, k, strides, padding, dilation, cardinality)


def _conv_raw(x, k_hwio, strides, padding, dilation=(1, 1), groups=1):
    kh, kw = k_hwio.shape[0], k_hwio.shape[1]
    w = k_hwio.permute(3, 2, 0, 1).contiguous()      # HWIO → OIHW
    s = (strides, strides) if isinstance(strides, int) else tuple(strides)
    if padding.upper() == "SAME":
        _, pt, pb = tf_same_pad(x.shape[2], kh, s[0], dilation[0])
        _, pl, pr = tf_same_pad(x.shape[3], kw, s[1], dilation[1])
        x = F.pad(x, (pl, pr, pt, pb))
    return F.conv2d(x, w, None, stride=s, padding=0, dilation=dilation, groups=groups)


def fixed_padding(x, kernel_size):
    """reference models.py:107-152 — symmetric explicit zero pad that depends only on the kernel."""
    pad_total = kernel_size - 1
    beg = pad_total // 2
    end = pad_total - beg
    return F.pad(x, (beg, end, beg, end))


def conv2d_fixed_padding(ctx, namer, x, filters, kernel_size, strides):
    """reference models.py:155-168 — stride>1: explicit pad + VALID, else SAME."""
    padding = "SAME"
    if strides > 1:
        x = fixed_padding(x, kernel_size)
        padding = "VALID"
    return conv2d(ctx, namer, x, filters, kernel_size, strides, padding)


def batch_norm(ctx: _Ctx, namer, x, branch_end: bool = False):
    """reference models.py:62-67 — inference BN without gamma/beta: (x-mean)*rsqrt(var+eps)."""
    scope = namer.next("batch_normalization")
    four_d = x.dim() == 4
    eps = BN_EPS_4D if four_d else BN_EPS_2D
    if ctx.calibrate:
        dims = (0, 2, 3) if four_d else (0,)
        xm = x.double()
        mean = xm.mean(dim=dims)
        var = xm.var(dim=dims, unbiased=False)
        ctx.params[scope + "/moving_mean"] = mean.float().numpy().copy()
        var = torch.clamp(var, min=1e-4) * (ctx.damp ** 2 if branch_end else 1.0)
        ctx.params[scope + "/moving_variance"] = var.float().numpy().copy()
    mean = ctx.tensor(scope + "/moving_mean")
    var = ctx.tensor(scope + "/moving_variance")
    inv = torch.rsqrt(var + eps)
    if four_d:
        return (x - mean.view(1, -1, 1, 1)) * inv.view(1, -1, 1, 1)
    return (x - mean.view(1, -1)) * inv.view(1, -1)


def stats_pool(x):
    """reference models.py:262-269 — mean and sqrt(population var + 1e-5) over H, concat on C."""
    mean = x.mean(dim=2, keepdim=True)
    var = x.var(dim=2, unbiased=False, keepdim=True)
    return torch.cat([mean, torch.sqrt(var + POOL_EPS)], dim=1)


def att_stats_pool(ctx, x, att_dim):
    """reference models.py:273-303 — attentive statistics pooling (att_with_mean_std=True): per (w, c) a softmax over time
    of conv1x1(tanh(conv1x1(concat(x, tiled mean, tiled std)))) weights the mean and the second moment."""
    scope = ctx.root.next("att_stats_pool")
    inner = arch._Namer(scope + "/")
    c = x.shape[1]
    mean = x.mean(dim=2, keepdim=True)
    var = x.var(dim=2, unbiased=False, keepdim=True)
    mean_std = torch.cat([mean, torch.sqrt(var + POOL_EPS)], dim=1).expand(-1, -1, x.shape[2], -1)   # tf.tile over time
    att_in = torch.cat([x, mean_std], dim=1)
    h = torch.tanh(conv2d(ctx, inner, att_in, att_dim, 1))
    logits = conv2d(ctx, inner, h, c, 1)
    w = torch.softmax(logits, dim=2)
    wmean = (x * w).sum(dim=2, keepdim=True)
    wss = (x * x * w).sum(dim=2, keepdim=True)
    wstd = torch.sqrt(wss - wmean * wmean + POOL_EPS)
    return torch.cat([wmean, wstd], dim=1)


def flatten_nhwc(x):
    """tf.layers.flatten of [N,1,W,2C] (NHWC): index = w*2C + k."""
    return x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)


def dense(ctx: _Ctx, x):
    """reference models.py:306-309 — bias-free matmul with 'dense/kernel' [Dflat,E]."""
    return x @ ctx.tensor("dense/kernel")


def _q(ctx, x):
    return ctx.quant(x) if ctx.quant is not None else x


def _tail(ctx: _Ctx, x, cfg=None):
    """stats_pool (or att_stats_pool) → flatten → BN → dense → BN (tdnn_model.py:142-153, res2net_model.py:229-243,
    dpn_model.py:153-167)."""
    pooled = att_stats_pool(ctx, x, cfg.att_dim) if (cfg is not None and cfg.att_pool) else stats_pool(x)
    x = flatten_nhwc(pooled)
    x = batch_norm(ctx, ctx.root, x)
    x = dense(ctx, x)
    x = batch_norm(ctx, ctx.root, x)
    return x


# ---------------------------------------------------------------- TDNN
def tdnn_forward(ctx: _Ctx, cfg: ModelConfig, x):
    """reference tdnn_model.py:24-30,128-155: five conv(SAME,dilated) → ReLU → BN blocks."""
    for li, (f, k, d) in enumerate(zip(cfg.tdnn_filters, cfg.tdnn_kernels, cfg.tdnn_dilations)):
        def layer(ctx_, t, f=f, k=k, d=d):
            t = conv2d(ctx_, ctx_.root, t, f, (k, 1), 1, "same", (d, 1))
            t = F.relu(t)
            return _q(ctx_, batch_norm(ctx_, ctx_.root, t))
        x = _traced(ctx, "tdnn%d" % (li + 1), layer, x)
    return _traced(ctx, "tail", lambda c_, t: _tail(c_, t), x)


# ---------------------------------------------------------------- Res2Net
def res2net_pad_conv_bn_relu(ctx: _Ctx, x, strides, split, width):
    """reference res2net_model.py:26-78 — hierarchical split 3x3."""
    if strides > 1:
        x = fixed_padding(x, 3)                                   # :27-28 (whole tensor, before the split)
    scope = ctx.root.next("conv2d")                               # :30
    kernel = ctx.tensor(scope + "/kernel")                        # [3,3,w,w*(split-1)]  :44-50
    inner = arch._Namer(scope + "/")
    padding = "SAME" if strides == 1 else "VALID"                 # :37
    xs = torch.split(x, width, dim=1)                             # :53
    ks = torch.split(kernel, width, dim=3)                        # :54
    outs = [_q(ctx, F.relu(batch_norm(ctx, inner, _conv_raw(xs[0], ks[0], strides, padding))))]   # :56-60
    for idx in range(1, split - 1):                               # :62-72
        inp = xs[idx]
        if strides == 1:
            inp = _q(ctx, inp + outs[idx - 1])
        outs.append(_q(ctx, F.relu(batch_norm(ctx, inner, _conv_raw(inp, ks[idx], strides, padding)))))
    if strides == 1:
        outs.append(xs[split - 1])                                # :74-75
    else:
        outs.append(_q(ctx, F.avg_pool2d(xs[split - 1], 3, strides, 0)))   # :76-77 VALID on the padded tensor
    return torch.cat(outs, dim=1)                                 # :78


def bottleneck_block_v1(ctx: _Ctx, x, filters, project, strides, split, width):
    """reference res2net_model.py:81-103."""
    shortcut = x
    if project:
        shortcut = conv2d_fixed_padding(ctx, ctx.root, x, filters * 4, 1, strides)     # :119-127
        shortcut = batch_norm(ctx, ctx.root, shortcut)                                 # :87
    y = conv2d_fixed_padding(ctx, ctx.root, x, split * width, 1, 1)                    # :89
    y = _q(ctx, F.relu(batch_norm(ctx, ctx.root, y)))                                  # :90-91
    y = res2net_pad_conv_bn_relu(ctx, y, strides, split, width)                        # :93-94
    y = conv2d_fixed_padding(ctx, ctx.root, y, filters * 4, 1, 1)                      # :98
    y = batch_norm(ctx, ctx.root, y, branch_end=True)                                  # :99
    return _q(ctx, F.relu(y + shortcut))                                               # :100-101


def res2net_forward(ctx: _Ctx, cfg: ModelConfig, x):
    """reference res2net_model.py:185-243."""
    def stem(ctx_, t):
        t = conv2d_fixed_padding(ctx_, ctx_.root, t, cfg.num_filters[0], 3, 1)         # :192-194
        return _q(ctx_, F.relu(batch_norm(ctx_, ctx_.root, t)))                        # :201-203
    x = _traced(ctx, "stem", stem, x)
    for li, nblocks in enumerate(cfg.block_sizes):                                     # :212-221
        for b in range(nblocks):
            x = _traced(ctx, "layer%d/block%d" % (li + 1, b),
                        lambda c_, t, li=li, b=b: bottleneck_block_v1(c_, t, cfg.num_filters[li], b == 0,
                                                                      cfg.block_strides[li] if b == 0 else 1, cfg.split, cfg.width[li]), x)
    return _traced(ctx, "tail", lambda c_, t: _tail(c_, t, cfg), x)


# ---------------------------------------------------------------- DPN
def bn_relu_conv(ctx: _Ctx, x, filters, kernel_size, strides, cardinality, branch_end=False):
    """reference dpn_model.py:40-45 — BN → ReLU → conv(SAME)."""
    x = _q(ctx, F.relu(batch_norm(ctx, ctx.root, x)))
    return conv2d(ctx, ctx.root, x, filters, kernel_size, strides, "same", (1, 1), cardinality, branch_end)


def dual_path_block(ctx: _Ctx, inputs, r, bw, inc, projection_type, cardinality):
    """reference dpn_model.py:57-87."""
    strides = 2 if projection_type == "downsampled" else 1
    if projection_type == "normal":
        res0, dense0 = inputs
        x = torch.cat(inputs, dim=1)                                          # :68-71
    else:
        x = torch.cat(inputs, dim=1) if isinstance(inputs, (list, tuple)) else inputs   # :73-74
        p = bn_relu_conv(ctx, x, bw + 2 * inc, 1, strides, 1)                 # :75
        res0, dense0 = p[:, :bw], p[:, bw:]                                   # :76-79
    y = bn_relu_conv(ctx, x, r, 1, 1, 1)                                      # :49
    y = bn_relu_conv(ctx, y, r, 3, strides, cardinality)                      # :50
    y = bn_relu_conv(ctx, y, bw + inc, 1, 1, 1, branch_end=True)              # :53
    return [_q(ctx, res0 + y[:, :bw]), _q(ctx, torch.cat([dense0, y[:, bw:]], dim=1))]   # :83-87


def dpn_forward(ctx: _Ctx, cfg: ModelConfig, x):
    """reference dpn_model.py:111-168."""
    def stem(ctx_, t):
        t = conv2d(ctx_, ctx_.root, t, cfg.init_features, 3, 1, "same")       # :33-34
        return _q(ctx_, F.relu(batch_norm(ctx_, ctx_.root, t)))               # :35-36
    x = _traced(ctx, "stem", stem, x)
    types = ["projected", "downsampled", "downsampled", "downsampled"]        # :92
    # between blocks the two paths travel as one tensor [res | dense] (dpn_model.py:77,84: the first bw channels are the residual path)
    for si, (_, _, r, bw, inc) in enumerate(arch.dpn_stage_channels(cfg)):
        for b in range(cfg.k_sec[si]):
            def block(ctx_, t, r=r, bw=bw, inc=inc, kind=(types[si] if b == 0 else "normal")):
                inp = t if kind != "normal" else [t[:, :bw], t[:, bw:]]
                return torch.cat(dual_path_block(ctx_, inp, r, bw, inc, kind, cfg.cardinality), dim=1)
            x = _traced(ctx, "stage%d/block%d" % (si + 1, b), block, x)

    def tail(ctx_, t):
        t = _q(ctx_, F.relu(batch_norm(ctx_, ctx_.root, t)))                  # :24-29
        return _tail(ctx_, t)
    return _traced(ctx, "tail", tail, x)


def _traced(ctx: _Ctx, name: str, fn, x):
    """Run one block; when tracing, keep its input / output and a closure that replays the block on another input."""
    counts = dict(ctx.root.counts)
    y = fn(ctx, x)
    if ctx.trace is not None:
        def replay(x_new, counts=counts):
            c2 = _Ctx(ctx.params, False, ctx.dtype, ctx.quant)
            c2.root.counts = dict(counts)
            with torch.no_grad():
                return fn(c2, x_new)
        ctx.record(name, x, y, replay)
    return y


_FORWARD = {FAMILY_TDNN: tdnn_forward, FAMILY_RES2NET: res2net_forward, FAMILY_DPN: dpn_forward}


def forward(cfg: ModelConfig, params: Dict[str, np.ndarray], feats: np.ndarray, calibrate: bool = False,
            dtype=torch.float32, quant=None, damp: float = 1.0) -> np.ndarray:
    """The frozen graph: ``inputs`` [N,T,F] (the singleton axis of export_inference_graph.py:40-43 is
    implied by ``cfg.expand_dim``) → ``outputs`` [N,E]."""
    x = torch.from_numpy(np.ascontiguousarray(feats)).to(dtype)
    assert x.dim() == 3
    if cfg.expand_dim == 2:      # [N,T,1,F] NHWC → NCHW [N,F,T,1]
        x = x.permute(0, 2, 1).unsqueeze(3)
    else:                        # [N,T,F,1] NHWC → NCHW [N,1,T,F]
        x = x.unsqueeze(1)
    ctx = _Ctx(params, calibrate, dtype, quant, damp)
    with torch.no_grad():
        y = _FORWARD[cfg.family](ctx, cfg, x)
    return y.float().numpy()


def trace_blocks(cfg: ModelConfig, params: Dict[str, np.ndarray], feats: np.ndarray):
    """Forward pass of ONE segment [T,F] that also returns the block-level trace: a list of records
    {"name", "in", "out" (NCHW float32 tensors; [N,E] for the tail), "replay": f(x) -> y} in execution order."""
    x = torch.from_numpy(np.ascontiguousarray(feats[None])).float()
    x = x.permute(0, 2, 1).unsqueeze(3) if cfg.expand_dim == 2 else x.unsqueeze(1)
    ctx = _Ctx(params)
    ctx.trace = []
    with torch.no_grad():
        y = _FORWARD[cfg.family](ctx, cfg, x)
    return y.float().numpy()[0], ctx.trace


def extract_utterance(cfg: ModelConfig, params, feats_tf: np.ndarray, dtype=torch.float32) -> np.ndarray:
    """reference tf_extract.py:96-111 — ≤1000-frame chunks, length-weighted mean of chunk embeddings."""
    plan = arch.chunk_plan(feats_tf.shape[0])
    if not plan:
        raise ZeroDivisionError("utterance shorter than %d frames (reference tf_extract.py:102,111)" % arch.MIN_FRAMES)
    acc = None
    total = 0
    for start, length in plan:
        y = forward(cfg, params, feats_tf[None, start:start + length], dtype=dtype)[0]
        acc = y * np.float32(length) if acc is None else acc + y * np.float32(length)
        total += length
    return (acc / np.float32(total)).astype(np.float32)


def synth_feats(rng: np.random.Generator, n: int, frames: int, feat_dim: int) -> np.ndarray:
    """Synthetic post-CMN FBANK-like features [n,frames,feat_dim]: zero-mean in time, per-utterance
    per-bin gains and AR(1) temporal correlation, so that utterances differ like speakers do."""
    z = rng.standard_normal((n, frames, feat_dim)).astype(np.float32)
    gain = np.exp(0.5 * rng.standard_normal((n, 1, feat_dim))).astype(np.float32)
    x = z * gain
    x[:, 1:] = np.float32(0.6) * x[:, :-1] + np.float32(0.8) * x[:, 1:]
    return np.ascontiguousarray(x, dtype=np.float32)


def init_params(cfg: ModelConfig, feat_dim: int, seed: int = 4321, calib_frames: int = 96,
                calib_batch: int = 16, damp: float = 4.0) -> Dict[str, np.ndarray]:
    """Seeded synthetic weights that behave like a trained network.

    Kernels: normal with variance 1/fan_in (the reference uses variance_scaling_initializer,
    models.py:193,308).  BN moving mean/variance := batch statistics of a seeded synthetic batch, so
    activations stay O(1) through the whole depth.  A random BN network is chaotic (perturbations grow
    ~1.4x per residual block, so bf16 rounding is amplified ~50x over 16 blocks), which no trained
    checkpoint is; ``damp`` scales every residual branch by 1/damp (Res2Net: branch-closing BN variance
    x damp^2; DPN: last 1x1 kernel / damp) to restore trained-like conditioning.  damp=1 gives the
    raw chaotic network.
    """
    rng = np.random.default_rng(seed)
    params: Dict[str, np.ndarray] = {}
    for spec in arch.enumerate_variables(cfg, feat_dim).specs:
        if spec.name.endswith("/kernel"):
            fan_in = int(np.prod(spec.shape[:-1]))
            params[spec.name] = (rng.standard_normal(spec.shape) / math.sqrt(fan_in)).astype(np.float32)
        elif spec.name.endswith("moving_mean"):
            params[spec.name] = np.zeros(spec.shape, np.float32)
        else:
            params[spec.name] = np.ones(spec.shape, np.float32)
    x = synth_feats(rng, calib_batch, calib_frames, feat_dim)
    forward(cfg, params, x, calibrate=True, damp=damp)
    return params
