"""Frozen-graph (.pb) weight source without TensorFlow.

The reference exports its models with ``freeze_graph --output_node_names=outputs``
(tensorflow/export_inference_model.sh:40-44): every variable becomes a ``Const`` node named by its TF
variable scope, the input placeholder is ``inputs`` (export_inference_graph.py:43) and the output identity is
``outputs``.  tf_extract.py:75-82 imports that GraphDef under the prefix ``model/``.  Here the GraphDef is
parsed with the protobuf definitions that ship with tensorboard and only the Const tensors are used; the
architecture is recovered by matching names and shapes against arch.enumerate_variables.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

from . import arch


def _protos():
    from tensorboard.compat.proto import graph_pb2, types_pb2  # noqa: WPS433
    from tensorboard.util import tensor_util
    return graph_pb2, types_pb2, tensor_util


def read_pb(path: str) -> Tuple[Dict[str, np.ndarray], Optional[Tuple[int, ...]]]:
    """→ ({const name: float32 array}, static shape of the ``inputs`` placeholder or None)."""
    graph_pb2, _, tensor_util = _protos()
    gd = graph_pb2.GraphDef()
    with open(path, "rb") as f:
        gd.ParseFromString(f.read())
    consts: Dict[str, np.ndarray] = {}
    in_shape = None
    for node in gd.node:
        name = node.name[6:] if node.name.startswith("model/") else node.name
        if node.op == "Const" and "value" in node.attr:
            arr = tensor_util.make_ndarray(node.attr["value"].tensor)
            if arr.dtype in (np.float32, np.float64, np.float16):
                consts[name] = np.asarray(arr, dtype=np.float32)
        elif node.op == "Placeholder" and name == "inputs" and "shape" in node.attr:
            in_shape = tuple(int(d.size) for d in node.attr["shape"].shape.dim)
    return consts, in_shape


def write_pb(path: str, params: Dict[str, np.ndarray], cfg: arch.ModelConfig, feat_dim: int) -> None:
    """Write a minimal frozen GraphDef holding ``params`` as Const nodes plus the ``inputs`` placeholder and
    an ``outputs`` identity — enough for this loader and shaped like the reference's export (used by tests
    and by users who keep weights as arrays)."""
    graph_pb2, types_pb2, tensor_util = _protos()
    gd = graph_pb2.GraphDef()
    ph = gd.node.add()
    ph.name, ph.op = "inputs", "Placeholder"
    ph.attr["dtype"].type = types_pb2.DT_FLOAT
    dims = [-1, -1, feat_dim]
    dims.insert(cfg.expand_dim, 1)                     # export_inference_graph.py:40-41
    for d in dims:
        ph.attr["shape"].shape.dim.add().size = d
    for name, arr in params.items():
        n = gd.node.add()
        n.name, n.op = name, "Const"
        n.attr["dtype"].type = types_pb2.DT_FLOAT
        n.attr["value"].tensor.CopyFrom(tensor_util.make_tensor_proto(np.asarray(arr, np.float32)))
    out = gd.node.add()
    out.name, out.op = "outputs", "Identity"
    out.input.append("batch_normalization/FusedBatchNorm")
    out.attr["T"].type = types_pb2.DT_FLOAT
    with open(path, "wb") as f:
        f.write(gd.SerializeToString())


def infer_model(consts: Dict[str, np.ndarray], in_shape, expand_dim: Optional[int] = None) -> Tuple[arch.ModelConfig, int]:
    """Find the (model config, feat_dim) whose variable table matches the graph's constants exactly."""
    feat_dims = []
    if in_shape is not None and len(in_shape) == 4 and in_shape[-1] > 1:
        feat_dims.append(in_shape[-1])                  # [N,T,1,F]
    if in_shape is not None and len(in_shape) == 4 and in_shape[2] > 1:
        feat_dims.append(in_shape[2])                   # [N,T,F,1]
    if "conv2d/kernel" in consts and consts["conv2d/kernel"].shape[1] == 1 and consts["conv2d/kernel"].shape[0] > 1:
        feat_dims.append(int(consts["conv2d/kernel"].shape[2]))   # TDNN first kernel [k,1,F,512]
    feat_dims += [40, 80]
    seen = set()
    candidates = []
    for cfg in arch.MODELS.values():
        if cfg.model_id in seen or (expand_dim is not None and cfg.expand_dim != expand_dim):
            continue
        seen.add(cfg.model_id)
        for fd in dict.fromkeys(feat_dims):
            shapes = arch.enumerate_variables(cfg, fd).shapes()
            if all(n in consts and tuple(consts[n].shape) == s for n, s in shapes.items()):
                candidates.append((cfg, fd))
                break
    if not candidates:
        raise ValueError("the frozen graph matches none of the known architectures: %s" % sorted(arch.MODELS))
    # several depths can share a prefix of names; the right one uses every kernel constant of the graph
    n_kernels = sum(1 for n in consts if n.endswith("/kernel"))
    exact = [c for c in candidates if sum(1 for n in arch.enumerate_variables(*c).names() if n.endswith("/kernel")) == n_kernels]
    return (exact or candidates)[0]
