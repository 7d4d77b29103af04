"""TensorFlow tensor-bundle reader (ckpt_loader.py) against the published format: hand-assembled table bytes, the
protobuf wire format cross-checked with the TensorShapeProto that ships with tensorboard, and a round trip through the writer."""
import struct

import numpy as np
import pytest

from voxsrc2020_speaker_verification_b200 import arch, ckpt_loader as ck


def test_table_bytes_assembled_by_hand():
    """One data block with prefix-compressed keys ('ab' then 'abc' sharing 2 bytes, then 'b'), restart array of one, an index
    block pointing at it, an empty metaindex block and the 48-byte footer — written out byte by byte, not with the module's writer."""
    data_block = (bytes([0, 2, 2]) + b"ab" + b"v1" +            # shared 0, unshared 2, value 2
                  bytes([2, 1, 3]) + b"c" + b"v22" +            # shared 2 -> key 'abc'
                  bytes([0, 1, 1]) + b"b" + b"w" +              # (a restart would normally be here; shared 0 is legal anywhere)
                  struct.pack("<II", 0, 1))                     # restarts [0], count 1
    trailer = b"\x00" + b"\x00\x00\x00\x00"                      # uncompressed, crc ignored
    meta_block = struct.pack("<II", 0, 1)
    meta_off = len(data_block) + 5
    index_entry_value = bytes([0, len(data_block)])              # handle: offset 0, size (both < 128: one-byte varints)
    index_block = bytes([0, 1, len(index_entry_value)]) + b"b" + index_entry_value + struct.pack("<II", 0, 1)
    index_off = meta_off + len(meta_block) + 5
    footer = bytes([meta_off, len(meta_block), index_off, len(index_block)])
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", 0xDB4775248B80FB57)
    table = data_block + trailer + meta_block + trailer + index_block + trailer + footer
    assert ck.read_table(table) == [(b"ab", b"v1"), (b"abc", b"v22"), (b"b", b"w")]
    with pytest.raises(ValueError):
        ck.read_table(table[:-1] + b"\x00")                      # bad magic
    snappy = bytearray(table)
    snappy[len(data_block)] = 1                                  # block type 1 = snappy
    with pytest.raises(ValueError):
        ck.read_table(bytes(snappy))


def test_entry_proto_matches_tensorboard_shape_proto():
    pb = pytest.importorskip("tensorboard.compat.proto.tensor_shape_pb2")
    shape = pb.TensorShapeProto()
    for d in (3, 3, 24, 72):
        shape.dim.add().size = d
    raw = ck._entry_proto(ck.DT_FLOAT, (3, 3, 24, 72), 0, 4096, 3 * 3 * 24 * 72 * 4, 0xDEADBEEF)
    e = ck._parse_entry(raw)
    assert (e["dtype"], e["shape"], e["offset"], e["size"], e["crc32c"]) == (1, [3, 3, 24, 72], 4096, 62208, 0xDEADBEEF)
    assert b"\x12" + bytes([len(shape.SerializeToString())]) + shape.SerializeToString() in raw     # field 2 = the same TensorShapeProto bytes
    assert ck._varint(ck._put_varint(300), 0) == (300, 2) and ck._put_varint(300) == b"\xac\x02"


@pytest.mark.parametrize("model_id,fd", [("tdnn", 40), ("res2net50_w24_s4_c32", 80), ("dpn68", 80)])
def test_round_trip_through_many_blocks(tmp_path, model_id, fd):
    cfg = arch.get_config(model_id)
    rng = np.random.default_rng(1)
    tensors = {s.name: rng.standard_normal(s.shape).astype(np.float32) for s in arch.enumerate_variables(cfg, fd).specs}
    extra = {"global_step": np.array(98765, np.int64), "conv2d/kernel/Momentum": np.zeros(tensors["conv2d/kernel"].shape, np.float32),
             "cm_linear_voxsrc2020/kernel": rng.standard_normal((256, 5994)).astype(np.float32)}     # training-only variables
    prefix = str(tmp_path / "model.ckpt-98765")
    ck.write_checkpoint(prefix, {**tensors, **extra}, block_entries=5)
    listed = ck.list_variables(prefix)
    assert set(listed) == set(tensors) | set(extra)
    assert listed["global_step"] == (ck.DT_INT64, ()) and listed["dense/kernel"][1] == tensors["dense/kernel"].shape
    got = ck.load_model_params(prefix, cfg, fd)
    assert set(got) == set(tensors)
    for k, v in tensors.items():
        np.testing.assert_array_equal(got[k], v)
    assert int(ck.read_checkpoint(prefix, ["global_step"])["global_step"]) == 98765
    with pytest.raises(KeyError):
        ck.read_checkpoint(prefix, ["no/such/variable"])
    with pytest.raises(ValueError):                              # right names, wrong feature dimension
        ck.load_model_params(prefix, cfg, fd + 8 if cfg.family != arch.FAMILY_TDNN else 2 * fd)
