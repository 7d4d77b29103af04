"""TensorFlow checkpoint (tensor bundle, "V2" format) weight source without TensorFlow.

The reference trains with ``tf.train.Saver`` (tf_train_tdnn.py:229-237) and exports the inference graph from
``model.ckpt-*`` (export_inference_graph.py:61-66, export_inference_model.sh:36-44).  A checkpoint is

  * ``<prefix>.index``                 an SSTable in LevelDB's table format [ext: tensorflow/core/lib/io/table]: data blocks of
                                       prefix-compressed ``key → value`` entries, an index block, a 48-byte footer with the block
                                       handles and the magic 0xdb4775248b80fb57.  Key ``""`` → BundleHeaderProto, every other key is
                                       a variable name → BundleEntryProto {dtype, shape, shard_id, offset, size, crc32c}.
  * ``<prefix>.data-0000k-of-0000n``   the raw little-endian tensor bytes, addressed by (shard_id, offset, size).

Only what inference needs is read: float tensors named like the frozen graph's constants (the optimizer's ``*/Momentum`` slots,
``global_step`` and the classifier's projection kernel are skipped by the caller's name filter).  Snappy-compressed index blocks
(``BundleWriter`` does not produce them) are rejected loudly.

No checkpoint ships with the reference and TensorFlow cannot be installed here, so this reader is checked against the published
format only: a writer for the same format (``write_checkpoint``, used by tests and by users who keep weights as arrays) and
hand-assembled table bytes in tests/test_ckpt_loader.py.  Parity with a real TF-written file is therefore UNPINNED.
"""
from __future__ import annotations

import os
import struct
import zlib
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
DT_FLOAT, DT_DOUBLE, DT_INT32, DT_INT64, DT_HALF = 1, 2, 3, 9, 19
_DTYPES = {DT_FLOAT: "<f4", DT_DOUBLE: "<f8", DT_INT32: "<i4", DT_INT64: "<i8", DT_HALF: "<f2"}


# ------------------------------------------------------------------------------------------------ varints / protobuf wire format
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _put_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _fields(buf: bytes) -> Iterable[Tuple[int, int, object]]:
    """(field number, wire type, value) of one serialized message (varint, 64-bit, length-delimited and 32-bit wire types)."""
    pos = 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        num, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]; pos += 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            v = buf[pos:pos + n]; pos += n
        elif wt == 5:
            v = buf[pos:pos + 4]; pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield num, wt, v


def _parse_entry(buf: bytes) -> Dict[str, object]:
    """BundleEntryProto: 1 dtype, 2 shape (TensorShapeProto: 2 dim {1 size}), 3 shard_id, 4 offset, 5 size, 6 crc32c, 7 slices."""
    e = {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "slices": 0}
    for num, wt, v in _fields(buf):
        if num == 1:
            e["dtype"] = v
        elif num == 2:
            for n2, _, dim in _fields(v):
                if n2 == 2:
                    size = 0
                    for n3, _, s in _fields(dim):
                        if n3 == 1:
                            size = s
                    e["shape"].append(size)
        elif num == 3:
            e["shard_id"] = v
        elif num == 4:
            e["offset"] = v
        elif num == 5:
            e["size"] = v
        elif num == 6:
            e["crc32c"] = struct.unpack("<I", v)[0]
        elif num == 7:
            e["slices"] += 1
    return e


# ------------------------------------------------------------------------------------------------ LevelDB table
def _block_entries(block: bytes) -> List[Tuple[bytes, bytes]]:
    """Entries of one table block: (shared, unshared, value_len) varints + key delta + value, then the restart array."""
    n_restarts = struct.unpack("<I", block[-4:])[0]
    end = len(block) - 4 - 4 * n_restarts
    out, pos, key = [], 0, b""
    while pos < end:
        shared, pos = _varint(block, pos)
        unshared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + block[pos:pos + unshared]
        pos += unshared
        out.append((key, block[pos:pos + vlen]))
        pos += vlen
    return out


def _read_block(data: bytes, offset: int, size: int) -> bytes:
    kind = data[offset + size]
    if kind != 0:
        raise ValueError("index block is compressed (type %d); only uncompressed tensor-bundle indexes are supported" % kind)
    return data[offset:offset + size]


def read_table(data: bytes) -> List[Tuple[bytes, bytes]]:
    """All (key, value) pairs of a LevelDB-format table, in key order."""
    if len(data) < 48 or struct.unpack("<Q", data[-8:])[0] != TABLE_MAGIC:
        raise ValueError("not a tensor-bundle index (bad table magic)")
    footer = data[-48:]
    _, pos = _varint(footer, 0)          # metaindex handle: offset, size
    _, pos = _varint(footer, pos)
    ioff, pos = _varint(footer, pos)     # index handle
    isize, pos = _varint(footer, pos)
    out = []
    for _, handle in _block_entries(_read_block(data, ioff, isize)):
        boff, p = _varint(handle, 0)
        bsize, _ = _varint(handle, p)
        out.extend(_block_entries(_read_block(data, boff, bsize)))
    return out


# ------------------------------------------------------------------------------------------------ reader
def list_variables(prefix: str) -> Dict[str, Tuple[int, Tuple[int, ...]]]:
    """{variable name: (dtype enum, shape)} of a checkpoint."""
    with open(prefix + ".index", "rb") as f:
        data = f.read()
    return {k.decode(): (e["dtype"], tuple(e["shape"])) for k, e in ((k, _parse_entry(v)) for k, v in read_table(data) if k != b"")}


def read_checkpoint(prefix: str, wanted: Optional[Iterable[str]] = None) -> Dict[str, np.ndarray]:
    """{name: array} for the variables in ``wanted`` (all of them if None); float tensors come back as float32."""
    with open(prefix + ".index", "rb") as f:
        data = f.read()
    entries = read_table(data)
    num_shards = 1
    for k, v in entries:
        if k == b"":
            for num, _, val in _fields(v):       # BundleHeaderProto: 1 num_shards, 2 endianness (0 = little)
                if num == 1:
                    num_shards = val
                elif num == 2 and val != 0:
                    raise ValueError("big-endian tensor bundles are not supported")
    want = set(wanted) if wanted is not None else None
    shards: Dict[int, object] = {}
    out: Dict[str, np.ndarray] = {}
    try:
        for k, v in entries:
            name = k.decode()
            if k == b"" or (want is not None and name not in want):
                continue
            e = _parse_entry(v)
            if e["slices"]:
                raise ValueError("%s is a partitioned variable (slices); not supported" % name)
            if e["dtype"] not in _DTYPES:
                raise ValueError("%s has unsupported dtype enum %d" % (name, e["dtype"]))
            sid = e["shard_id"]
            if sid not in shards:
                shards[sid] = open("%s.data-%05d-of-%05d" % (prefix, sid, num_shards), "rb")
            fd = shards[sid]
            fd.seek(e["offset"])
            raw = fd.read(e["size"])
            dt = np.dtype(_DTYPES[e["dtype"]])
            n = int(np.prod(e["shape"])) if e["shape"] else 1
            if len(raw) != e["size"] or e["size"] != n * dt.itemsize:
                raise ValueError("%s: %d bytes on disk, shape %s needs %d" % (name, len(raw), e["shape"], n * dt.itemsize))
            arr = np.frombuffer(raw, dtype=dt).reshape(e["shape"])
            out[name] = np.asarray(arr, np.float32) if dt.kind == "f" else arr.copy()
    finally:
        for fd in shards.values():
            fd.close()
    if want is not None:
        missing = sorted(want - set(out))
        if missing:
            raise KeyError("checkpoint %s lacks %d variables, e.g. %s" % (prefix, len(missing), missing[:3]))
    return out


def load_model_params(prefix: str, cfg, feat_dim: int) -> Dict[str, np.ndarray]:
    """The variables ``cfg`` needs (arch.enumerate_variables), validated by shape; everything else in the checkpoint (Momentum slots,
    global_step, the classifier's projection kernel — tf_train_tdnn.py:229-237, tf_projection.py:180) is ignored."""
    from . import arch
    specs = arch.enumerate_variables(cfg, feat_dim).specs
    params = read_checkpoint(prefix, [s.name for s in specs])
    for s in specs:
        if tuple(params[s.name].shape) != tuple(s.shape):
            raise ValueError("%s: checkpoint shape %s, model %s at %d-dim features needs %s"
                             % (s.name, params[s.name].shape, cfg.model_id, feat_dim, s.shape))
    return params


# ------------------------------------------------------------------------------------------------ writer (tests, array-held weights)
def _block(entries: List[Tuple[bytes, bytes]], restart_interval: int = 16) -> bytes:
    out, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(entries):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _entry_proto(dtype: int, shape, shard_id: int, offset: int, size: int, crc: int) -> bytes:
    shape_msg = b"".join(b"\x12" + _put_varint(len(d)) + d for d in (b"\x08" + _put_varint(int(s)) for s in shape))
    out = b"\x08" + _put_varint(dtype) + b"\x12" + _put_varint(len(shape_msg)) + shape_msg
    if shard_id:
        out += b"\x18" + _put_varint(shard_id)
    if offset:
        out += b"\x20" + _put_varint(offset)
    out += b"\x28" + _put_varint(size) + b"\x35" + struct.pack("<I", crc & 0xFFFFFFFF)
    return out


def write_checkpoint(prefix: str, tensors: Dict[str, np.ndarray], block_entries: int = 64) -> None:
    """Write ``tensors`` as a one-shard tensor bundle (uncompressed index blocks).  The crc32c fields are filled with zlib's CRC-32
    (this module's reader does not verify them; TensorFlow would)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    items, offset = [], 0
    with open(prefix + ".data-00000-of-00001", "wb") as fd:
        for name in sorted(tensors, key=lambda s: s.encode()):
            a = np.asarray(tensors[name]).copy(order="C")          # (ascontiguousarray would turn a scalar into shape (1,))
            dt = {np.dtype("float32"): DT_FLOAT, np.dtype("float64"): DT_DOUBLE, np.dtype("int32"): DT_INT32, np.dtype("int64"): DT_INT64}[a.dtype]
            raw = a.astype(a.dtype.newbyteorder("<")).tobytes()
            fd.write(raw)
            items.append((name.encode(), _entry_proto(dt, a.shape, 0, offset, len(raw), zlib.crc32(raw))))
            offset += len(raw)
    header = b"\x08\x01" + b"\x1a\x02\x08\x01"                 # num_shards = 1, version { producer: 1 }
    entries = [(b"", header)] + items
    out, index = bytearray(), []
    for i in range(0, len(entries), block_entries):
        chunk = entries[i:i + block_entries]
        blk = _block(chunk)
        index.append((chunk[-1][0], _put_varint(len(out)) + _put_varint(len(blk))))
        out += blk + b"\x00" + struct.pack("<I", 0)             # no compression, crc unchecked
    meta = _block([])
    meta_handle = _put_varint(len(out)) + _put_varint(len(meta))
    out += meta + b"\x00" + struct.pack("<I", 0)
    iblk = _block(index, restart_interval=1)
    index_handle = _put_varint(len(out)) + _put_varint(len(iblk))
    out += iblk + b"\x00" + struct.pack("<I", 0)
    footer = meta_handle + index_handle
    out += footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))
