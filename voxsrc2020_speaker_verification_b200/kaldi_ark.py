"""Kaldi ark/scp wire formats for the extraction and scoring stages (buffered, not byte-at-a-time).

Formats follow what the reference reads and writes through its vendored kaldi_io.py:
  * float matrix   ``key ' ' \\0B 'FM ' \\x04 <i32 rows> \\x04 <i32 cols> <f32 row-major>``  (kaldi_io.py:437-454)
  * compressed     ``\\0B 'CM ' <f32 min> <f32 range> <i32 rows> <i32 cols>`` + per-column 4 x uint16
                   percentiles + column-major uint8                                         (kaldi_io.py:471-504)
  * float vector   ``key ' ' \\0B 'FV ' \\x04 <i32 dim> <f32…>``                               (kaldi_io.py:304-334)
  * scp            ``key <path>[:<offset of \\0B>]`` — what ``copy-vector ark:- ark,scp:…`` emits for the
                   extractor's output (reference tf_extract.py:65) [ext: Kaldi table I/O].
"""
from __future__ import annotations

import io
import os
import struct
from typing import BinaryIO, Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np


class UnknownMatrixHeader(Exception):
    pass


class UnknownVectorHeader(Exception):
    pass


def _read_exact(fd: BinaryIO, n: int) -> bytes:
    buf = fd.read(n)
    if len(buf) != n:
        raise EOFError("truncated ark: wanted %d bytes, got %d" % (n, len(buf)))
    return buf


def _read_key(fd: BinaryIO) -> Optional[str]:
    """Key = bytes up to the first space (kaldi_io.py:123-141).  Uses peek() when available so the
    common case costs one call instead of one read per byte."""
    chunks = []
    while True:
        peeked = fd.peek(256) if hasattr(fd, "peek") else b""
        if peeked:
            sp = peeked.find(b" ")
            if sp >= 0:
                chunks.append(_read_exact(fd, sp + 1)[:-1])
                break
            chunks.append(_read_exact(fd, len(peeked)))
            continue
        ch = fd.read(1)
        if ch == b"" or ch == b" ":
            break
        chunks.append(ch)
    key = b"".join(chunks).decode("utf-8").strip()
    return key if key else None


def _open(path_or_fd, mode="rb"):
    if hasattr(path_or_fd, "read") or hasattr(path_or_fd, "write"):
        return path_or_fd, False
    path = path_or_fd
    if path.startswith("ark:") or path.startswith("scp:"):
        path = path.split(":", 1)[1]
    fd = open(path, mode)
    if "r" in mode and not hasattr(fd, "peek"):
        fd = io.BufferedReader(fd)
    return fd, True


# ------------------------------------------------------------------ matrices
def _decode_compressed(fd: BinaryIO, header: str) -> np.ndarray:
    """CompressedMatrix → float32 [rows, cols].  'CM ' as in kaldi_io.py:471-504; 'CM2'/'CM3' are the
    header-less 16/8-bit variants Kaldi writes for short matrices [ext]."""
    gmin, grange, rows, cols = struct.unpack("<ffii", _read_exact(fd, 16))
    gmin, grange = np.float32(gmin), np.float32(grange)
    if header == "CM ":
        colh = np.frombuffer(_read_exact(fd, cols * 8), dtype="<u2").reshape(cols, 4)
        colh = (colh.astype(np.float32) * grange * np.float32(1.52590218966964e-05) + gmin).astype(np.float32)
        data = np.frombuffer(_read_exact(fd, cols * rows), dtype=np.uint8).reshape(cols, rows)
        p0, p25, p75, p100 = (colh[:, i:i + 1] for i in range(4))
        d = data.astype(np.float32)
        lo = p0 + (p25 - p0) / np.float32(64.0) * d
        mid = p25 + (p75 - p25) / np.float32(128.0) * (d - np.float32(64.0))
        hi = p75 + (p100 - p75) / np.float32(63.0) * (d - np.float32(192.0))
        mat = np.where(data <= 64, lo, np.where(data > 192, hi, mid)).astype(np.float32)
        return np.ascontiguousarray(mat.T)
    if header == "CM2":
        data = np.frombuffer(_read_exact(fd, 2 * rows * cols), dtype="<u2").reshape(rows, cols)
        return (data.astype(np.float32) * (grange * np.float32(1.0 / 65535.0)) + gmin).astype(np.float32)
    if header == "CM3":
        data = np.frombuffer(_read_exact(fd, rows * cols), dtype=np.uint8).reshape(rows, cols)
        return (data.astype(np.float32) * (grange * np.float32(1.0 / 255.0)) + gmin).astype(np.float32)
    raise UnknownMatrixHeader(header)


def read_mat(fd: BinaryIO) -> np.ndarray:
    """One binary matrix at the current position (after the key)."""
    flag = _read_exact(fd, 2)
    if flag != b"\0B":
        raise UnknownMatrixHeader("only binary Kaldi matrices are supported, got %r" % flag)
    header = _read_exact(fd, 3).decode("latin1")
    if header.startswith("CM"):
        return _decode_compressed(fd, header)
    if header == "FM ":
        dt, size = "<f4", 4
    elif header == "DM ":
        dt, size = "<f8", 8
    else:
        raise UnknownMatrixHeader("The header contained %r" % header)
    s1, rows, s2, cols = struct.unpack("<bibi", _read_exact(fd, 10))
    if s1 != 4 or s2 != 4:
        raise UnknownMatrixHeader("bad int-size markers %d %d" % (s1, s2))
    buf = _read_exact(fd, rows * cols * size)
    return np.frombuffer(buf, dtype=dt).reshape(rows, cols)


def read_mat_raw(fd: BinaryIO):
    """Like ``read_mat`` but leaves 'CM ' records undecoded for the device decoder (svx_decode_compressed):
    returns ('CM ', payload bytes starting at the global header, rows, cols) or ('dense', ndarray)."""
    flag = _read_exact(fd, 2)
    if flag != b"\0B":
        raise UnknownMatrixHeader("only binary Kaldi matrices are supported, got %r" % flag)
    header = _read_exact(fd, 3).decode("latin1")
    if header == "CM ":
        head = _read_exact(fd, 16)
        _, _, rows, cols = struct.unpack("<ffii", head)
        return "CM ", head + _read_exact(fd, cols * 8 + cols * rows), rows, cols
    fd.seek(-5, 1)
    m = read_mat(fd)
    return "dense", m, m.shape[0], m.shape[1]


def peek_mat_shape(fd: BinaryIO) -> Tuple[int, int]:
    """(rows, cols) of the binary matrix at the current position, reading only its header (the position is restored)."""
    pos = fd.tell()
    try:
        if _read_exact(fd, 2) != b"\0B":
            raise UnknownMatrixHeader("only binary Kaldi matrices are supported")
        header = _read_exact(fd, 3).decode("latin1")
        if header == "CM ":
            _, _, rows, cols = struct.unpack("<ffii", _read_exact(fd, 16))
        elif header in ("FM ", "DM "):
            s1, rows, s2, cols = struct.unpack("<bibi", _read_exact(fd, 10))
        else:
            raise UnknownMatrixHeader("The header contained %r" % header)
        return rows, cols
    finally:
        fd.seek(pos)


def scp_shapes(path: str) -> List[Tuple[str, int, int]]:
    """(key, rows, cols) of every record a script file addresses — headers only (what a rank needs to balance the work)."""
    out, cur_path, cur_fd = [], None, None
    try:
        for key, rx, off in read_scp(path):
            if rx != cur_path:
                if cur_fd is not None:
                    cur_fd.close()
                cur_fd, cur_path = open(rx, "rb"), rx
            if off is None:
                raise ValueError("scp entry %s has no byte offset" % key)
            cur_fd.seek(off)
            out.append((key,) + peek_mat_shape(cur_fd))
    finally:
        if cur_fd is not None:
            cur_fd.close()
    return out


def read_mat_ark_raw(path_or_fd):
    """(key, kind, payload, rows, cols) for every record of a matrix ark (see ``read_mat_raw``)."""
    fd, own = _open(path_or_fd)
    try:
        key = _read_key(fd)
        while key:
            yield (key,) + read_mat_raw(fd)
            key = _read_key(fd)
    finally:
        if own:
            fd.close()


def read_mat_ark(path_or_fd) -> Iterator[Tuple[str, np.ndarray]]:
    """(key, matrix) pairs of a matrix ark, in file order (kaldi_io.py:367-385)."""
    fd, own = _open(path_or_fd)
    try:
        key = _read_key(fd)
        while key:
            yield key, read_mat(fd)
            key = _read_key(fd)
    finally:
        if own:
            fd.close()


def read_scp(path: str) -> List[Tuple[str, str, Optional[int]]]:
    """Script file lines ``key path[:offset]`` → (key, path, offset)."""
    if path.startswith("scp:"):
        path = path[4:]
    out = []
    with open(path, "r") as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            key, rx = line.split(None, 1)
            off = None
            head, sep, tail = rx.rpartition(":")
            if sep and tail.isdigit():
                rx, off = head, int(tail)
            out.append((key, rx, off))
    return out


def read_mat_scp(path: str) -> Iterator[Tuple[str, np.ndarray]]:
    """(key, matrix) pairs addressed by a script file; keeps one ark open at a time."""
    cur_path, cur_fd = None, None
    try:
        for key, rx, off in read_scp(path):
            if rx != cur_path:
                if cur_fd is not None:
                    cur_fd.close()
                cur_fd, cur_path = open(rx, "rb"), rx
            if off is not None:
                cur_fd.seek(off)
            yield key, read_mat(cur_fd)
    finally:
        if cur_fd is not None:
            cur_fd.close()


def read_mat_scp_raw(path: str, only: Optional[Sequence[int]] = None):
    """(key, kind, payload, rows, cols) addressed by a script file; 'CM ' records stay compressed (see ``read_mat_raw``).
    ``only``: line indices to read (a rank's share), in that order."""
    cur_path, cur_fd = None, None
    try:
        entries = read_scp(path)
        if only is not None:
            entries = [entries[i] for i in only]
        for key, rx, off in entries:
            if rx != cur_path:
                if cur_fd is not None:
                    cur_fd.close()
                cur_fd, cur_path = open(rx, "rb"), rx
            if off is not None:
                cur_fd.seek(off)
            yield (key,) + read_mat_raw(cur_fd)
    finally:
        if cur_fd is not None:
            cur_fd.close()


def write_mat(fd: BinaryIO, m: np.ndarray, key: str = "") -> int:
    """Write one 'FM ' record; returns the byte offset of its \\0B flag (for scp lines)."""
    m = np.ascontiguousarray(m, dtype=np.float32)
    if key:
        fd.write((key + " ").encode("utf-8"))
    off = fd.tell()
    fd.write(b"\0BFM ")
    fd.write(struct.pack("<bibi", 4, m.shape[0], 4, m.shape[1]))
    fd.write(m.tobytes())
    return off


# ------------------------------------------------------------------ vectors
def read_vec_flt(fd: BinaryIO) -> np.ndarray:
    flag = _read_exact(fd, 2)
    if flag != b"\0B":
        raise UnknownVectorHeader("only binary Kaldi vectors are supported, got %r" % flag)
    header = _read_exact(fd, 3).decode("latin1")
    if header == "FV ":
        dt, size = "<f4", 4
    elif header == "DV ":
        dt, size = "<f8", 8
    else:
        raise UnknownVectorHeader("The header contained %r" % header)
    if _read_exact(fd, 1) != b"\4":
        raise UnknownVectorHeader("bad int-size marker")
    (dim,) = struct.unpack("<i", _read_exact(fd, 4))
    if dim == 0:
        return np.array([], dtype=np.float32)
    return np.frombuffer(_read_exact(fd, dim * size), dtype=dt)


def read_vec_flt_ark(path_or_fd) -> Iterator[Tuple[str, np.ndarray]]:
    """(key, vector) pairs of a vector ark, in file order (kaldi_io.py:249-264)."""
    fd, own = _open(path_or_fd)
    try:
        key = _read_key(fd)
        while key:
            yield key, read_vec_flt(fd)
            key = _read_key(fd)
    finally:
        if own:
            fd.close()


def read_vec_ark_matrix(path: str) -> Tuple[List[str], np.ndarray]:
    """Whole vector ark → (keys, float32 [n, dim]) with one bulk parse per record."""
    keys, rows = [], []
    for k, v in read_vec_flt_ark(path):
        keys.append(k)
        rows.append(np.asarray(v, dtype=np.float32))
    if not rows:
        return keys, np.zeros((0, 0), np.float32)
    return keys, np.stack(rows)


def write_vec_flt(fd: BinaryIO, v: np.ndarray, key: str = "") -> int:
    """Write one 'FV ' record exactly as kaldi_io.write_vec_flt does (kaldi_io.py:322-332); returns
    the byte offset of the \\0B flag."""
    v = np.ascontiguousarray(v, dtype=np.float32)
    if key:
        fd.write((key + " ").encode("utf-8"))
    off = fd.tell()
    fd.write(b"\0BFV \4")
    fd.write(struct.pack("<i", v.shape[0]))
    fd.write(v.tobytes())
    return off


class VectorArkScpWriter:
    """Stand-in for ``ark:| copy-vector ark:- ark,scp:W.ark,W.scp`` (reference tf_extract.py:65)."""

    def __init__(self, wspec: str):
        self.ark_path = os.path.abspath(wspec + ".ark")
        self.scp_path = wspec + ".scp"
        self._ark = open(self.ark_path, "wb")
        self._scp = open(self.scp_path, "w")

    def write(self, key: str, vec: np.ndarray) -> None:
        off = write_vec_flt(self._ark, vec, key)
        self._scp.write("%s %s:%d\n" % (key, self.ark_path, off))

    def close(self) -> None:
        self._ark.close()
        self._scp.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
