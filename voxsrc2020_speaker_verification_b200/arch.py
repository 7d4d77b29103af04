"""Architecture table for the embedding extractors and the frozen-graph weight-name contract.

The reference selects a model by ``(--module_source, --model_id)`` and builds it at the root
variable scope (reference tensorflow/export_inference_graph.py:38-48).  This module restates the
module-level instances as plain data:

  * ``tdnn``                       reference tensorflow/models/tdnn_model.py:158-161
  * ``res2net50_w24_s4_c64``       reference tensorflow/models/res2net_model.py:246-250
  * ``res2net50_w24_s4_c32``       reference tensorflow/models/res2net_model.py:252-256
  * ``res2net50_w8_s6_c16``        reference tensorflow/models/res2net_model.py:258-262
  * ``res2net200_w8_s6_c16``       composed: depth-200 block sizes (res2net_model.py:278) with the
                                   w8_s6_c16 widths (res2net_model.py:258-262); named in README.md:47
  * ``res2net{101,152,200}_w24_s4_c32_att``  reference tensorflow/models/res2net_model.py:264-280 (attentive statistics pooling)
  * ``dpn68``                      reference tensorflow/models/dpn_model.py:171

and enumerates, in TF1 creation order, the variables a frozen ``.pb`` of each model holds
(``conv2d{,_k}/kernel``, ``batch_normalization{,_k}/moving_{mean,variance}``, ``dense/kernel``).
TF1 uniquifies ``default_name`` scopes per enclosing scope as ``name, name_1, name_2 …``
(reference models.py:174, res2net_model.py:30).  No checkpoint ships with the reference, so the
table is derived, and the loader validates every tensor by shape.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

FAMILY_TDNN, FAMILY_RES2NET, FAMILY_DPN = 0, 1, 2

BN_EPS_4D = 1.001e-5   # TF fused batch norm clamps eps to >= 1.001e-5 for 4-D inputs [ext]
BN_EPS_2D = 1e-5       # reference models.py:20; 2-D inputs take the non-fused path
POOL_EPS = 1e-5        # reference models.py:262
MAX_CHUNK_FRAMES = 1000  # reference tf_extract.py:96
MIN_FRAMES = 25          # reference tf_extract.py:101-102


@dataclass(frozen=True)
class ModelConfig:
    model_id: str
    family: int
    expand_dim: int                 # 2: [N,T,1,F] (TDNN); 3: [N,T,F,1] (2-D conv models)
    embed_dim: int
    # TDNN (tdnn_model.py:158-161)
    tdnn_filters: Tuple[int, ...] = ()
    tdnn_kernels: Tuple[int, ...] = ()
    tdnn_dilations: Tuple[int, ...] = ()
    # Res2Net (res2net_model.py:139-183)
    num_filters: Tuple[int, ...] = ()
    width: Tuple[int, ...] = ()
    split: int = 0
    block_sizes: Tuple[int, ...] = ()
    block_strides: Tuple[int, ...] = ()
    # DPN (dpn_model.py:90-109)
    init_features: int = 0
    bw: int = 0
    k_r: int = 0
    cardinality: int = 0
    k_sec: Tuple[int, ...] = ()
    inc_sec: Tuple[int, ...] = ()
    # attentive statistics pooling instead of plain statistics pooling (models.py:273-303, res2net_model.py:264-280)
    att_pool: bool = False
    att_dim: int = 128


MODELS: Dict[str, ModelConfig] = {
    "tdnn": ModelConfig("tdnn", FAMILY_TDNN, 2, 256,
                        tdnn_filters=(512, 512, 512, 512, 1536),
                        tdnn_kernels=(5, 3, 3, 1, 1),
                        tdnn_dilations=(1, 2, 3, 1, 1)),
    "res2net50_w24_s4_c64": ModelConfig("res2net50_w24_s4_c64", FAMILY_RES2NET, 3, 256,
                                        num_filters=(64, 128, 256, 512), width=(24, 48, 96, 192), split=4,
                                        block_sizes=(3, 4, 6, 3), block_strides=(1, 2, 2, 2)),
    "res2net50_w24_s4_c32": ModelConfig("res2net50_w24_s4_c32", FAMILY_RES2NET, 3, 256,
                                        num_filters=(32, 64, 128, 256), width=(24, 48, 96, 192), split=4,
                                        block_sizes=(3, 4, 6, 3), block_strides=(1, 2, 2, 2)),
    "res2net50_w8_s6_c16": ModelConfig("res2net50_w8_s6_c16", FAMILY_RES2NET, 3, 192,
                                       num_filters=(16, 32, 64, 128), width=(8, 16, 32, 64), split=6,
                                       block_sizes=(3, 4, 6, 3), block_strides=(1, 2, 2, 2)),
    "res2net200_w8_s6_c16": ModelConfig("res2net200_w8_s6_c16", FAMILY_RES2NET, 3, 192,
                                        num_filters=(16, 32, 64, 128), width=(8, 16, 32, 64), split=6,
                                        block_sizes=(3, 24, 36, 3), block_strides=(1, 2, 2, 2)),
    "res2net101_w24_s4_c32_att": ModelConfig("res2net101_w24_s4_c32_att", FAMILY_RES2NET, 3, 256,
                                             num_filters=(32, 64, 128, 256), width=(24, 48, 96, 192), split=4,
                                             block_sizes=(3, 4, 23, 3), block_strides=(1, 2, 2, 2), att_pool=True),
    "res2net152_w24_s4_c32_att": ModelConfig("res2net152_w24_s4_c32_att", FAMILY_RES2NET, 3, 256,
                                             num_filters=(32, 64, 128, 256), width=(24, 48, 96, 192), split=4,
                                             block_sizes=(3, 8, 36, 3), block_strides=(1, 2, 2, 2), att_pool=True),
    "res2net200_w24_s4_c32_att": ModelConfig("res2net200_w24_s4_c32_att", FAMILY_RES2NET, 3, 256,
                                             num_filters=(32, 64, 128, 256), width=(24, 48, 96, 192), split=4,
                                             block_sizes=(3, 24, 36, 3), block_strides=(1, 2, 2, 2), att_pool=True),
    # small attentive model for tests (not in the reference): res2net50_w8_s6_c16 with att_stats_pool
    "res2net50_w8_s6_c16_att": ModelConfig("res2net50_w8_s6_c16_att", FAMILY_RES2NET, 3, 192,
                                           num_filters=(16, 32, 64, 128), width=(8, 16, 32, 64), split=6,
                                           block_sizes=(3, 4, 6, 3), block_strides=(1, 2, 2, 2), att_pool=True),
    "dpn68": ModelConfig("dpn68", FAMILY_DPN, 3, 256,
                         init_features=10, bw=64, k_r=128, cardinality=32,
                         k_sec=(3, 4, 12, 3), inc_sec=(16, 32, 32, 64)),
}
# aliases used by BASELINE.json / the training scripts (same graphs, different training heads)
MODELS["tdnn_voxsrc2020"] = MODELS["tdnn"]
MODELS["dpn68_voxsrc2020"] = MODELS["dpn68"]


def get_config(model_id: str) -> ModelConfig:
    try:
        return MODELS[model_id]
    except KeyError:
        raise KeyError("unknown model_id %r; known: %s" % (model_id, sorted(MODELS))) from None


def ceil_half(n: int) -> int:
    return (n + 1) // 2


class _Namer:
    """TF1 ``variable_scope(None, default_name=…)`` uniquifier for one enclosing scope."""

    def __init__(self, prefix: str = ""):
        self.prefix = prefix
        self.counts: Dict[str, int] = {}

    def next(self, base: str) -> str:
        k = self.counts.get(base, 0)
        self.counts[base] = k + 1
        name = base if k == 0 else "%s_%d" % (base, k)
        return self.prefix + name


@dataclass
class VarSpec:
    name: str
    shape: Tuple[int, ...]


@dataclass
class GraphVars:
    """Variables of one model in creation order plus bookkeeping the builders share."""
    specs: List[VarSpec] = field(default_factory=list)

    def add(self, name: str, shape) -> str:
        self.specs.append(VarSpec(name, tuple(int(s) for s in shape)))
        return name

    def names(self) -> List[str]:
        return [s.name for s in self.specs]

    def shapes(self) -> Dict[str, Tuple[int, ...]]:
        return {s.name: s.shape for s in self.specs}


def _bn(gv: GraphVars, namer: _Namer, channels: int) -> str:
    scope = namer.next("batch_normalization")
    gv.add(scope + "/moving_mean", (channels,))
    gv.add(scope + "/moving_variance", (channels,))
    return scope


def _conv(gv: GraphVars, namer: _Namer, kh: int, kw: int, cin_per_group: int, cout: int) -> str:
    scope = namer.next("conv2d")
    gv.add(scope + "/kernel", (kh, kw, cin_per_group, cout))
    return scope


def flat_dim(cfg: ModelConfig, feat_dim: int) -> int:
    """Length of the flattened statistics vector fed to ``dense`` (models.py:262-269 + flatten)."""
    if cfg.family == FAMILY_TDNN:
        return 2 * cfg.tdnn_filters[-1]
    w = feat_dim
    if cfg.family == FAMILY_RES2NET:
        for s in cfg.block_strides:
            if s == 2:
                w = ceil_half(w)
        return w * 2 * cfg.num_filters[-1] * 4
    if cfg.family == FAMILY_DPN:
        for _ in range(3):
            w = ceil_half(w)
        return w * 2 * dpn_stage_channels(cfg)[-1][1]
    raise ValueError(cfg.family)


def dpn_stage_channels(cfg: ModelConfig) -> List[Tuple[int, int, int, int, int]]:
    """Per stage: (C_in, C_out, r, bw, inc) following dpn_model.py:115-149."""
    out = []
    c = cfg.init_features
    for i in range(4):
        bw = cfg.bw * (2 ** i)
        inc = cfg.inc_sec[i]
        r = cfg.k_r * bw // cfg.bw
        c_out = bw + 2 * inc + (cfg.k_sec[i] - 1) * inc + inc   # proj gives bw+2inc, every block adds inc
        out.append((c, c_out, r, bw, inc))
        c = c_out
    return out


def enumerate_variables(cfg: ModelConfig, feat_dim: int) -> GraphVars:
    """All frozen-graph variables of ``cfg`` at ``feat_dim`` in TF1 creation order."""
    gv = GraphVars()
    root = _Namer()
    if cfg.family == FAMILY_TDNN:
        cin = feat_dim
        for f, k in zip(cfg.tdnn_filters, cfg.tdnn_kernels):
            _conv(gv, root, k, 1, cin, f)        # tdnn_model.py:25-27
            _bn(gv, root, f)                     # tdnn_model.py:29
            cin = f
    elif cfg.family == FAMILY_RES2NET:
        _conv(gv, root, 3, 3, 1, cfg.num_filters[0])          # res2net_model.py:192-194
        _bn(gv, root, cfg.num_filters[0])                     # res2net_model.py:202
        cin = cfg.num_filters[0]
        for li, nblocks in enumerate(cfg.block_sizes):
            filt, w, s = cfg.num_filters[li], cfg.width[li], cfg.split
            cout = filt * 4
            for b in range(nblocks):
                if b == 0:
                    _conv(gv, root, 1, 1, cin, cout)          # projection, res2net_model.py:125-127
                    _bn(gv, root, cout)                       # res2net_model.py:87
                _conv(gv, root, 1, 1, cin, s * w)             # res2net_model.py:89
                _bn(gv, root, s * w)                          # res2net_model.py:90
                hscope = root.next("conv2d")                  # res2net_model.py:30
                gv.add(hscope + "/kernel", (3, 3, w, w * (s - 1)))   # res2net_model.py:44-50
                inner = _Namer(hscope + "/")
                for _ in range(s - 1):
                    _bn(gv, inner, w)                         # res2net_model.py:56-72
                _conv(gv, root, 1, 1, s * w, cout)            # res2net_model.py:98
                _bn(gv, root, cout)                           # res2net_model.py:99
                cin = cout
        if cfg.att_pool:                                      # models.py:273-303: scope att_stats_pool, two bias-free 1x1 convs
            ascope = root.next("att_stats_pool")
            ainner = _Namer(ascope + "/")
            _conv(gv, ainner, 1, 1, 3 * cin, cfg.att_dim)     # on concat(inputs, tiled mean, tiled std)
            _conv(gv, ainner, 1, 1, cfg.att_dim, cin)
    elif cfg.family == FAMILY_DPN:
        _conv(gv, root, 3, 3, 1, cfg.init_features)           # dpn_model.py:33-34
        _bn(gv, root, cfg.init_features)                      # dpn_model.py:35
        for si, (c_in, c_out, r, bw, inc) in enumerate(dpn_stage_channels(cfg)):
            c = c_in
            cpg = r // cfg.cardinality
            for b in range(cfg.k_sec[si]):
                if b == 0:
                    _bn(gv, root, c)                          # dpn_model.py:75 → :41
                    _conv(gv, root, 1, 1, c, bw + 2 * inc)
                    c_block_in = c
                    c = bw + 2 * inc
                else:
                    c_block_in = c
                _bn(gv, root, c_block_in)                     # dpn_model.py:49
                _conv(gv, root, 1, 1, c_block_in, r)
                _bn(gv, root, r)                              # dpn_model.py:50
                _conv(gv, root, 3, 3, cpg, r)
                _bn(gv, root, r)                              # dpn_model.py:53
                _conv(gv, root, 1, 1, r, bw + inc)
                c = c + inc
            assert c == c_out, (c, c_out)
        _bn(gv, root, c)                                      # dpn_model.py:27
    else:
        raise ValueError(cfg.family)
    d = flat_dim(cfg, feat_dim)
    _bn(gv, root, d)                                          # 2-D BN before dense
    gv.add("dense/kernel", (d, cfg.embed_dim))                # models.py:306-309
    _bn(gv, root, cfg.embed_dim)                              # 2-D BN after dense
    return gv


def param_count(cfg: ModelConfig, feat_dim: int) -> int:
    """Trainable parameters (kernels only; BN has no gamma/beta, models.py:66-67)."""
    total = 0
    for spec in enumerate_variables(cfg, feat_dim).specs:
        if spec.name.endswith("/kernel"):
            n = 1
            for s in spec.shape:
                n *= s
            total += n
    return total


def chunk_plan(num_frames: int) -> List[Tuple[int, int]]:
    """(start, length) of every chunk the reference feeds for one utterance (tf_extract.py:101-110).

    ``num_chunks = 1 + (T - 25) // 1000``; a tail shorter than 25 frames is dropped; an utterance
    shorter than 25 frames yields zero chunks (the reference then divides by zero).
    """
    n = 1 + (num_frames - MIN_FRAMES) // MAX_CHUNK_FRAMES
    plan = []
    for i in range(max(n, 0)):
        if (i + 1) * MAX_CHUNK_FRAMES <= num_frames:
            length = MAX_CHUNK_FRAMES
        else:
            length = num_frames - i * MAX_CHUNK_FRAMES
        plan.append((i * MAX_CHUNK_FRAMES, length))
    return plan
