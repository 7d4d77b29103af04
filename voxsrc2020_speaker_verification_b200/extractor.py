"""Embedding extractor: the host-side mirror of the reference's frozen-graph session.

``Extractor`` plays the role of ``tf.Session`` + ``model/inputs:0 → model/outputs:0``
(reference tensorflow/tf_extract.py:75-82,94-111): it owns one libsvx extractor handle on one GPU, is fed
whole utterances [T, F] and returns one embedding per utterance, applying the reference's chunk rule inside
the library.  Unlike the reference (batch 1, one sess.run per chunk) it packs many utterances of any lengths
into one launch sequence.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch

from . import arch, lib, pb_loader


def _cfg_struct(cfg: arch.ModelConfig, feat_dim: int) -> lib.ModelConfigStruct:
    s = lib.ModelConfigStruct()
    s.family, s.feat_dim, s.embed_dim = cfg.family, feat_dim, cfg.embed_dim
    s.tdnn_layers = len(cfg.tdnn_filters)
    for i, v in enumerate(cfg.tdnn_filters):
        s.tdnn_filters[i] = v
    for i, v in enumerate(cfg.tdnn_kernels):
        s.tdnn_kernels[i] = v
    for i, v in enumerate(cfg.tdnn_dilations):
        s.tdnn_dilations[i] = v
    for i, v in enumerate(cfg.num_filters):
        s.num_filters[i] = v
    for i, v in enumerate(cfg.width):
        s.width[i] = v
    s.split = cfg.split
    for i, v in enumerate(cfg.block_sizes):
        s.block_sizes[i] = v
    for i, v in enumerate(cfg.block_strides):
        s.block_strides[i] = v
    s.init_features, s.bw, s.k_r, s.cardinality = cfg.init_features, cfg.bw, cfg.k_r, cfg.cardinality
    for i, v in enumerate(cfg.k_sec):
        s.k_sec[i] = v
    for i, v in enumerate(cfg.inc_sec):
        s.inc_sec[i] = v
    s.att_pool, s.att_dim = (1 if cfg.att_pool else 0), cfg.att_dim
    return s


class Extractor:
    def __init__(self, model_id: str, feat_dim: int, device: int = 0, precision: str = "fp16"):
        self.cfg = arch.get_config(model_id)
        self.feat_dim = int(feat_dim)
        self.device = int(device)
        self.precision = precision
        self._lib = lib.load()
        self._h = ctypes.c_void_p()
        prec = {"fp16": lib.PRECISION_FP16, "bf16": lib.PRECISION_BF16}[precision]
        cs = _cfg_struct(self.cfg, self.feat_dim)
        lib.check(self._lib.svx_extractor_create(ctypes.byref(cs), self.device, prec, ctypes.byref(self._h)))
        self.embed_dim = self._lib.svx_extractor_embed_dim(self._h)
        self._finalized = False
        self._pinned_in: Optional[torch.Tensor] = None
        self._pinned_out: Optional[torch.Tensor] = None

    # ------------------------------------------------------------------ weights
    def expected_tensors(self) -> Dict[str, tuple]:
        out = {}
        name, ndim, shape = ctypes.c_char_p(), ctypes.c_int(), (ctypes.c_int64 * 4)()
        for i in range(self._lib.svx_extractor_num_tensors(self._h)):
            lib.check(self._lib.svx_extractor_tensor_info(self._h, i, ctypes.byref(name), ctypes.byref(ndim), shape))
            out[name.value.decode()] = tuple(shape[j] for j in range(ndim.value))
        return out

    def load_params(self, params: Dict[str, np.ndarray]) -> "Extractor":
        """``params``: TF variable name → float32 array in TF layout (kernels HWIO)."""
        expected = self.expected_tensors()
        missing = [n for n in expected if n not in params]
        if missing:
            raise KeyError("weights missing for %d tensors, e.g. %s" % (len(missing), missing[:3]))
        for name, shape in expected.items():
            a = np.ascontiguousarray(params[name], dtype=np.float32)
            if tuple(a.shape) != shape:
                raise ValueError("%s: expected shape %s, got %s" % (name, shape, a.shape))
            shp = (ctypes.c_int64 * a.ndim)(*a.shape)
            lib.check(self._lib.svx_extractor_set_tensor(self._h, name.encode(), a.ctypes.data_as(ctypes.c_void_p), a.ndim, shp))
        lib.check(self._lib.svx_extractor_finalize(self._h))
        self._finalized = True
        return self

    @classmethod
    def from_pb(cls, pb_file: str, expand_dim: Optional[int] = None, device: int = 0, precision: str = "fp16",
                model_id: Optional[str] = None, feat_dim: Optional[int] = None) -> "Extractor":
        """Build from a frozen graph as tf_extract.py does from ``--pb-file`` (tf_extract.py:75-82)."""
        consts, in_shape = pb_loader.read_pb(pb_file)
        if model_id is None or feat_dim is None:
            cfg, fd = pb_loader.infer_model(consts, in_shape, expand_dim)
            model_id, feat_dim = cfg.model_id, fd
        ex = cls(model_id, feat_dim, device, precision)
        if expand_dim is not None and expand_dim != ex.cfg.expand_dim:
            raise ValueError("--expand-dim %d does not fit model %s (needs %d)" % (expand_dim, model_id, ex.cfg.expand_dim))
        return ex.load_params(consts)

    @classmethod
    def from_checkpoint(cls, prefix: str, model_id: str, feat_dim: int, device: int = 0, precision: str = "fp16") -> "Extractor":
        """Build straight from a training checkpoint ``model.ckpt-N`` (tensor bundle: ``.index`` + ``.data-*``) instead of the
        frozen ``.pb`` the reference exports from it (export_inference_graph.py:61-66, export_inference_model.sh:36-44).  A
        checkpoint does not say which architecture it holds, so ``model_id`` and ``feat_dim`` are required; every variable is
        validated by shape."""
        from . import ckpt_loader
        ex = cls(model_id, feat_dim, device, precision)
        return ex.load_params(ckpt_loader.load_model_params(prefix, ex.cfg, ex.feat_dim))

    def set_option(self, key: str, value: int) -> None:
        lib.check(self._lib.svx_extractor_set_option(self._h, key.encode(), int(value)))

    def set_dump_dir(self, path: Optional[str]) -> None:
        """Parity tooling: every op of the following runs writes its destination tensor (raw 16-bit NHWC tall image) into
        ``path``; None switches it off."""
        lib.check(self._lib.svx_extractor_set_dump_dir(self._h, path.encode() if path else None))

    @property
    def last_launches(self) -> int:
        return int(self._lib.svx_extractor_last_launches(self._h))

    def conv_time(self):
        """(ms, algorithmic FLOPs) of the tensor-core conv launches of the last call (option time_convs=1)."""
        ms, fl = ctypes.c_double(), ctypes.c_double()
        lib.check(self._lib.svx_extractor_conv_time(self._h, ctypes.byref(ms), ctypes.byref(fl)))
        return ms.value, fl.value

    # ------------------------------------------------------------------ running
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def cmvn_sliding(self, feats_dev: torch.Tensor, frame_offsets: np.ndarray, cmn_window: int = 300, center: bool = True,
                     min_window: int = 100) -> torch.Tensor:
        """In place on a CUDA tensor [total, F]: what ``apply-cmvn-sliding --norm-vars=false --center=true --cmn-window=300``
        does in front of the network (reference tf_extract.py:63).  ``min_window`` only matters for center=False
        (Kaldi's --min-cmn-window)."""
        frame_offsets = np.ascontiguousarray(frame_offsets, dtype=np.int32)
        lib.check(self._lib.svx_cmvn_sliding(ctypes.c_void_p(feats_dev.data_ptr()), ctypes.c_void_p(feats_dev.data_ptr()),
                                             frame_offsets.ctypes.data_as(ctypes.c_void_p), frame_offsets.shape[0] - 1, self.feat_dim,
                                             int(cmn_window), int(center), int(min(min_window, cmn_window)), ctypes.c_void_p(self._stream())))
        return feats_dev

    def decode_compressed(self, payloads: Sequence[bytes], rows: Sequence[int]) -> torch.Tensor:
        """Kaldi 'CM ' records (payloads from ``kaldi_ark.read_mat_raw``) → CUDA fp32 [sum(rows), F], decoded on the device
        bit-identically to kaldi_io._read_compressed_mat (reference kaldi_io.py:471-504)."""
        n = len(payloads)
        for i, (p, r) in enumerate(zip(payloads, rows)):       # the record's own header: min f32, range f32, rows i32, cols i32
            hr, hc = np.frombuffer(p, dtype="<i4", count=2, offset=8) if len(p) >= 16 else (-1, -1)
            if hc != self.feat_dim or hr != r:
                raise ValueError("compressed record %d is %d x %d, expected %d x %d (the model takes %d-dim features)"
                                 % (i, hr, hc, r, self.feat_dim, self.feat_dim))
        rec_off = np.zeros(n, np.int64)
        np.cumsum([len(p) for p in payloads[:-1]], out=rec_off[1:])
        offs = np.zeros(n + 1, np.int32)
        np.cumsum(np.asarray(rows, np.int64), out=offs[1:])
        blob = torch.frombuffer(bytearray(b"".join(payloads)), dtype=torch.uint8).to(torch.device("cuda", self.device))
        out = torch.empty((int(offs[-1]), self.feat_dim), dtype=torch.float32, device=blob.device)
        lib.check(self._lib.svx_decode_compressed(ctypes.c_void_p(blob.data_ptr()), int(blob.numel()), rec_off.ctypes.data_as(ctypes.c_void_p),
                                                  offs.ctypes.data_as(ctypes.c_void_p), n, self.feat_dim, ctypes.c_void_p(out.data_ptr()),
                                                  ctypes.c_void_p(self._stream())))
        return out

    def extract(self, utterances: Sequence[np.ndarray], cmvn: bool = False) -> np.ndarray:
        """Host → host: list of [T_i, F] float32 matrices → [n, E] float32 (chunk rule applied).  ``cmvn``: apply the
        sliding-window mean normalisation on the device first (raw FBANK in, as read from the feature ark)."""
        n = len(utterances)
        if n == 0:
            return np.zeros((0, self.embed_dim), np.float32)
        lens = np.array([u.shape[0] for u in utterances], np.int64)
        for u in utterances:
            if u.ndim != 2 or u.shape[1] != self.feat_dim:
                raise ValueError("expected [T, %d] features, got %s" % (self.feat_dim, u.shape))
        offs = np.zeros(n + 1, np.int32)
        np.cumsum(lens, out=offs[1:])
        total = int(offs[-1])
        if self._pinned_in is None or self._pinned_in.numel() < total * self.feat_dim:
            self._pinned_in = torch.empty(max(total * self.feat_dim, 1), dtype=torch.float32).pin_memory()
        if self._pinned_out is None or self._pinned_out.numel() < n * self.embed_dim:
            self._pinned_out = torch.empty(n * self.embed_dim, dtype=torch.float32).pin_memory()
        stage = self._pinned_in.numpy()[: total * self.feat_dim].reshape(total, self.feat_dim)
        for i, u in enumerate(utterances):
            stage[offs[i]:offs[i + 1]] = u
        if cmvn:
            dev = self._pinned_in[: total * self.feat_dim].to(torch.device("cuda", self.device), non_blocking=True).view(total, self.feat_dim)
            self.cmvn_sliding(dev, offs)
            self.extract_packed(dev, offs, self._pinned_out)
        else:
            self.extract_packed(self._pinned_in, offs, self._pinned_out)
        return self._pinned_out.numpy()[: n * self.embed_dim].reshape(n, self.embed_dim).copy()

    def extract_packed(self, feats: torch.Tensor, frame_offsets: np.ndarray, out: torch.Tensor) -> None:
        """Packed frames [total, F] (host-pinned or CUDA tensor) + int32 offsets [n+1] → ``out`` [n, E]
        (host or CUDA tensor).  Host buffers: the call returns after the device→host copy completed."""
        frame_offsets = np.ascontiguousarray(frame_offsets, dtype=np.int32)
        n = frame_offsets.shape[0] - 1
        lib.check(self._lib.svx_extractor_extract(
            self._h, ctypes.c_void_p(feats.data_ptr()), int(feats.is_cuda),
            frame_offsets.ctypes.data_as(ctypes.c_void_p), n,
            ctypes.c_void_p(out.data_ptr()), int(out.is_cuda), ctypes.c_void_p(self._stream())))

    def run_segments(self, feats_dev: torch.Tensor, frame_offsets: np.ndarray) -> torch.Tensor:
        """Device → device, no chunk rule: one graph evaluation per segment (sess.run semantics)."""
        frame_offsets = np.ascontiguousarray(frame_offsets, dtype=np.int32)
        n = frame_offsets.shape[0] - 1
        out = torch.empty((n, self.embed_dim), dtype=torch.float32, device=feats_dev.device)
        lib.check(self._lib.svx_extractor_run_segments(
            self._h, ctypes.c_void_p(feats_dev.data_ptr()), frame_offsets.ctypes.data_as(ctypes.c_void_p), n,
            ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(self._stream())))
        return out

    def extract_bucketed(self, utterances: Sequence[np.ndarray], max_frames: int = 60000, cmvn: bool = False) -> np.ndarray:
        """Length-bucketed batches (sorted by frame count, ≤ max_frames per launch sequence), results
        returned in the caller's order."""
        order = np.argsort([u.shape[0] for u in utterances], kind="stable")
        out = np.zeros((len(utterances), self.embed_dim), np.float32)
        batch: List[int] = []
        frames = 0
        for idx in list(order) + [None]:
            if idx is not None and (not batch or frames + utterances[idx].shape[0] <= max_frames):
                batch.append(int(idx))
                frames += utterances[idx].shape[0]
                continue
            if batch:
                out[batch] = self.extract([utterances[i] for i in batch], cmvn=cmvn)
            batch, frames = ([int(idx)], utterances[idx].shape[0]) if idx is not None else ([], 0)
        return out

    def close(self) -> None:
        if self._h:
            self._lib.svx_extractor_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
