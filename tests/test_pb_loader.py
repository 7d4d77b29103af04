"""Frozen-graph weight source: write a .pb shaped like the reference's export, read it back, infer the model."""
import numpy as np
import pytest

from voxsrc2020_speaker_verification_b200 import arch, pb_loader


def fake_params(cfg, fd, seed=0):
    rng = np.random.default_rng(seed)
    return {s.name: rng.standard_normal(s.shape).astype(np.float32) for s in arch.enumerate_variables(cfg, fd).specs}


@pytest.mark.parametrize("model_id,fd", [("tdnn", 40), ("res2net50_w8_s6_c16", 80), ("dpn68", 40), ("res2net50_w24_s4_c32", 80)])
def test_round_trip_and_inference(tmp_path, model_id, fd):
    cfg = arch.get_config(model_id)
    params = fake_params(cfg, fd)
    path = str(tmp_path / "m.pb")
    pb_loader.write_pb(path, params, cfg, fd)
    consts, in_shape = pb_loader.read_pb(path)
    assert in_shape is not None and fd in in_shape
    for k, v in params.items():
        np.testing.assert_array_equal(consts[k], v)
    got_cfg, got_fd = pb_loader.infer_model(consts, in_shape, cfg.expand_dim)
    assert (got_cfg.model_id, got_fd) == (cfg.model_id, fd)


def test_model_prefix_is_stripped_and_unknown_graph_rejected(tmp_path):
    cfg = arch.get_config("tdnn")
    params = {"model/" + k: v for k, v in fake_params(cfg, 40).items()}
    path = str(tmp_path / "m.pb")
    pb_loader.write_pb(path, params, cfg, 40)
    consts, in_shape = pb_loader.read_pb(path)
    assert "conv2d/kernel" in consts
    bad = dict(consts)
    bad["conv2d_2/kernel"] = np.zeros((3, 1, 512, 511), np.float32)
    with pytest.raises(ValueError):
        pb_loader.infer_model(bad, in_shape, 2)
