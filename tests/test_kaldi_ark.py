"""Kaldi ark/scp codec against what the reference's kaldi_io.py decoded (tests/golden/io) and round trips."""
import io
import os

import numpy as np
import pytest

from voxsrc2020_speaker_verification_b200 import kaldi_ark


def test_float_and_compressed_matrices_match_reference(golden_dir):
    g = os.path.join(golden_dir, "io")
    ref = np.load(os.path.join(g, "ref_feats.npz"))
    for tag in ("fm", "cm"):
        got = dict(kaldi_ark.read_mat_ark(os.path.join(g, "feats_%s.ark" % tag)))
        assert list(got) == ["uttA", "spk/uttB.wav", "uttC"]
        for k, m in got.items():
            np.testing.assert_array_equal(m, ref["%s:%s" % (tag, k)])


def test_vector_ark_matches_reference_writer(golden_dir):
    g = os.path.join(golden_dir, "score")
    keys, mat = kaldi_ark.read_vec_ark_matrix(os.path.join(g, "test.ark"))
    assert len(keys) == 1100 and mat.shape == (1100, 64)
    buf = io.BytesIO()
    for k, v in zip(keys, mat):
        kaldi_ark.write_vec_flt(buf, v, k)
    assert buf.getvalue() == open(os.path.join(g, "test.ark"), "rb").read()     # byte-identical records


def test_scp_round_trip_and_cat(tmp_path):
    rng = np.random.default_rng(0)
    vecs = {"a/b.wav": rng.standard_normal(256).astype(np.float32), "c": rng.standard_normal(256).astype(np.float32)}
    with kaldi_ark.VectorArkScpWriter(str(tmp_path / "xvector.1")) as w:
        for k, v in vecs.items():
            w.write(k, v)
    with kaldi_ark.VectorArkScpWriter(str(tmp_path / "xvector.2")) as w:
        w.write("d", vecs["c"] * 2)
    # shards are concatenated with `cat` (eval_inference_model.sh:38-39): records are self-delimiting
    cat = tmp_path / "xvector.ark"
    cat.write_bytes((tmp_path / "xvector.1.ark").read_bytes() + (tmp_path / "xvector.2.ark").read_bytes())
    got = dict(kaldi_ark.read_vec_flt_ark(str(cat)))
    assert list(got) == ["a/b.wav", "c", "d"]
    np.testing.assert_array_equal(got["a/b.wav"], vecs["a/b.wav"])
    for key, path, off in kaldi_ark.read_scp(str(tmp_path / "xvector.1.scp")):
        with open(path, "rb") as f:
            f.seek(off)
            np.testing.assert_array_equal(kaldi_ark.read_vec_flt(f), vecs[key])


def test_mat_scp_with_offsets(tmp_path):
    rng = np.random.default_rng(1)
    mats = {"u1": rng.standard_normal((30, 40)).astype(np.float32), "u2": rng.standard_normal((26, 40)).astype(np.float32)}
    ark = tmp_path / "feats.ark"
    with open(ark, "wb") as f, open(tmp_path / "feats.scp", "w") as scp:
        for k, m in mats.items():
            off = kaldi_ark.write_mat(f, m, k)
            scp.write("%s %s:%d\n" % (k, ark, off))
    got = dict(kaldi_ark.read_mat_scp(str(tmp_path / "feats.scp")))
    for k in mats:
        np.testing.assert_array_equal(got[k], mats[k])


def test_empty_and_truncated(tmp_path):
    p = tmp_path / "empty.ark"
    p.write_bytes(b"")
    assert list(kaldi_ark.read_vec_flt_ark(str(p))) == []
    q = tmp_path / "bad.ark"
    q.write_bytes(b"utt \0BFV \x04\x10\x00\x00\x00abc")
    with pytest.raises(EOFError):
        list(kaldi_ark.read_vec_flt_ark(str(q)))
    r = tmp_path / "hdr.ark"
    r.write_bytes(b"utt \0BXX \x04")
    with pytest.raises(kaldi_ark.UnknownVectorHeader):
        list(kaldi_ark.read_vec_flt_ark(str(r)))
