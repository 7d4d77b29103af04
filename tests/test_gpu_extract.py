"""Parity of the CUDA extractor (through the C ABI) with the CPU oracle on seeded inputs and weights.

Tolerance (BASELINE.json north_star): cosine similarity >= 0.9999 per utterance.  fp16 operands meet it on all
three networks; bf16 operands are checked against the looser bound their 8-bit mantissa allows.
"""
import numpy as np
import pytest
import torch

from oracle import cmn_oracle, net_oracle
from voxsrc2020_speaker_verification_b200 import arch

pytestmark = pytest.mark.gpu

COS_TOL = 0.9999
# Cosine is blind to a gain error, and the chunk rule averages UN-normalised chunk embeddings, so magnitude is bounded too:
# relative L2 (cos >= 0.9999 alone allows 1.41e-2) and the ratio of the norms.
REL_L2_TOL = 1.6e-2
NORM_TOL = 5e-3


def cosines(a, b):
    return (a * b).sum(1) / np.linalg.norm(a, axis=1) / np.linalg.norm(b, axis=1)


def check_close(got, want, cos_tol=COS_TOL, rel_tol=REL_L2_TOL, norm_tol=NORM_TOL):
    cos = cosines(got, want)
    rel = np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)
    ratio = np.linalg.norm(got, axis=1) / np.linalg.norm(want, axis=1)
    assert cos.min() >= cos_tol, (cos, rel, ratio)
    assert rel.max() <= rel_tol, (cos, rel, ratio)
    assert np.abs(ratio - 1).max() <= norm_tol, (cos, rel, ratio)


_cache = {}


def model(model_id, feat_dim, precision="fp16"):
    from voxsrc2020_speaker_verification_b200.extractor import Extractor
    key = (model_id, feat_dim, precision)
    if key not in _cache:
        cfg = arch.get_config(model_id)
        pkey = (model_id, feat_dim)
        if pkey not in _cache:
            _cache[pkey] = net_oracle.init_params(cfg, feat_dim, seed=4321)
        ex = Extractor(model_id, feat_dim, precision=precision).load_params(_cache[pkey])
        _cache[key] = (cfg, _cache[pkey], ex)
    return _cache[key]


def oracle_segments(cfg, params, utts):
    return np.stack([net_oracle.forward(cfg, params, u[None])[0] for u in utts])


def run_segments(ex, utts):
    feats = torch.from_numpy(np.concatenate(utts, 0)).cuda()
    offs = np.zeros(len(utts) + 1, np.int32)
    offs[1:] = np.cumsum([u.shape[0] for u in utts])
    out = ex.run_segments(feats, offs)
    torch.cuda.synchronize()
    return out.cpu().numpy()


CASES = [
    ("tdnn", 40, [61, 25, 320, 100]),
    ("tdnn", 80, [47, 200]),
    ("res2net50_w24_s4_c32", 80, [200, 57, 25, 26]),
    ("res2net50_w24_s4_c32", 40, [64, 33]),
    ("res2net50_w8_s6_c16", 80, [40, 121]),
    ("res2net50_w24_s4_c64", 40, [48]),
    ("res2net50_w8_s6_c16_att", 40, [48, 75, 25]),      # attentive statistics pooling (models.py:273-303)
    ("dpn68", 80, [64, 57, 25, 26, 200]),
    ("dpn68", 40, [50, 31]),
]


@pytest.mark.parametrize("model_id,feat_dim,lens", CASES)
@pytest.mark.parametrize("path", ["tensor", "tensor_2d_only", "simple"])
def test_segments_match_oracle(model_id, feat_dim, lens, path):
    """tensor: flat tcgen05 kernel (stride-1 convs) + 2-D tiled tcgen05 kernel (stride-2 convs); tensor_2d_only: the flat kernel
    switched off, so every conv the 2-D tiled kernel can take runs there; simple: CUDA-core kernel everywhere."""
    cfg, params, ex = model(model_id, feat_dim)
    ex.set_option("force_simple", 1 if path == "simple" else 0)
    ex.set_option("no_flat", 1 if path == "tensor_2d_only" else 0)
    rng = np.random.default_rng(1234)
    utts = [net_oracle.synth_feats(rng, 1, t, feat_dim)[0] for t in lens]
    got = run_segments(ex, utts)
    ex.set_option("force_simple", 0)
    ex.set_option("no_flat", 0)
    want = oracle_segments(cfg, params, utts)
    assert np.isfinite(got).all()
    check_close(got, want)
    assert ex.last_launches > 0


def test_batch_independence():
    """A segment's embedding must not depend on its neighbours in the tall image (zero rows isolate them)."""
    cfg, params, ex = model("res2net50_w24_s4_c32", 80)
    rng = np.random.default_rng(7)
    utts = [net_oracle.synth_feats(rng, 1, t, 80)[0] for t in (90, 41, 200, 33)]
    together = run_segments(ex, utts)
    alone = np.concatenate([run_segments(ex, [u]) for u in utts])
    np.testing.assert_allclose(together, alone, rtol=0, atol=1e-5)


@pytest.mark.parametrize("model_id,feat_dim", [("tdnn", 40), ("res2net50_w8_s6_c16", 80)])
def test_chunk_rule(model_id, feat_dim):
    """tf_extract.py:96-111: 1000-frame chunks, tail < 25 dropped, length-weighted mean; T=1024 vs 1025."""
    cfg, params, ex = model(model_id, feat_dim)
    rng = np.random.default_rng(99)
    utts = [net_oracle.synth_feats(rng, 1, t, feat_dim)[0] for t in (1024, 1025, 300, 2030)]
    got = ex.extract(utts)
    want = np.stack([net_oracle.extract_utterance(cfg, params, u) for u in utts])
    check_close(got, want)
    # the 24-frame tail of the 1024-frame utterance is dropped: identical to its first 1000 frames
    first = ex.extract([utts[0][:1000]])
    np.testing.assert_allclose(got[0], first[0], atol=1e-5)


def test_short_utterance_fails_loudly():
    _, _, ex = model("tdnn", 40)
    from voxsrc2020_speaker_verification_b200.lib import SvxError
    with pytest.raises(SvxError):
        ex.extract([np.zeros((24, 40), np.float32)])


# bf16 operands (the dtype the north star names; fp16 is the default because it meets cosine >= 0.9999): bounds from the measured
# values of profiles/r02_parity_probe.txt — TDNN 0.9998, Res2Net-50 0.9994-0.9997, DPN-68 0.9975-0.9983 on the damped synthetic
# weights.  Every BLOCK is at the bf16 rounding floor (tests/test_gpu_blocks.py); the whole-network figure is that floor amplified
# by a random network's conditioning.
@pytest.mark.parametrize("model_id,feat_dim,lens,cos_tol,rel_tol,norm_tol",
                         [("tdnn", 40, [80, 200], 0.9995, 3.5e-2, 1e-2),
                          ("res2net50_w24_s4_c32", 80, [200, 57], 0.999, 5e-2, 1e-2),
                          ("dpn68", 80, [200, 57], 0.996, 1e-1, 2e-2)])
def test_bf16_precision_mode(model_id, feat_dim, lens, cos_tol, rel_tol, norm_tol):
    cfg, params, ex = model(model_id, feat_dim, "bf16")
    rng = np.random.default_rng(5)
    utts = [net_oracle.synth_feats(rng, 1, t, feat_dim)[0] for t in lens]
    check_close(run_segments(ex, utts), oracle_segments(cfg, params, utts), cos_tol, rel_tol, norm_tol)


def test_res2net200_full_utterances_with_chunk_rule():
    """BASELINE configs[3]: res2net200_w8_s6_c16 (66 bottleneck blocks, 8-channel splits) on 300 / 1025 / 3000-frame utterances
    through the chunk rule.  Every block is at the fp16 rounding floor (~5e-4 relative, test_gpu_blocks.py); 66 of them on random
    weights amplify it to cosine 0.9988-0.9998 (profiles/r02_parity_probe.txt) — the bound here is that measured figure, not the
    0.9999 the 16-block networks meet."""
    cfg, params, ex = model("res2net200_w8_s6_c16", 80)
    rng = np.random.default_rng(77)
    utts = [net_oracle.synth_feats(rng, 1, t, 80)[0] for t in (300, 1025, 3000)]
    got = ex.extract(utts)
    want = np.stack([net_oracle.extract_utterance(cfg, params, u) for u in utts])
    check_close(got, want, 0.998, 7e-2, 1e-2)


def test_full_depth_attentive_model_and_long_dpn():
    """res2net101_w24_s4_c32_att (33 blocks + attentive statistics pooling, res2net_model.py:264-268) and dpn68 at 599 / 600 frames
    (odd / even lengths through three TF-SAME stride-2 stages)."""
    cfg, params, ex = model("res2net101_w24_s4_c32_att", 80)
    rng = np.random.default_rng(78)
    utts = [net_oracle.synth_feats(rng, 1, t, 80)[0] for t in (200, 57)]
    check_close(run_segments(ex, utts), oracle_segments(cfg, params, utts))
    cfg, params, ex = model("dpn68", 80)
    utts = [net_oracle.synth_feats(rng, 1, t, 80)[0] for t in (599, 600)]
    check_close(run_segments(ex, utts), oracle_segments(cfg, params, utts))


def test_extract_bucketed_order():
    cfg, params, ex = model("tdnn", 40)
    rng = np.random.default_rng(3)
    utts = [net_oracle.synth_feats(rng, 1, t, 40)[0] for t in (200, 30, 120, 55, 400)]
    a = ex.extract_bucketed(utts, max_frames=300)
    b = ex.extract(utts)
    np.testing.assert_allclose(a, b, atol=1e-5)


def test_many_spans_per_cta():
    """Enough pixels that every persistent CTA walks several spans (ring wrap-around, slot reuse, both TMEM buffers)."""
    cfg, params, ex = model("res2net50_w24_s4_c32", 80)
    rng = np.random.default_rng(21)
    feats = net_oracle.synth_feats(rng, 40, 120, 80)
    utts = [feats[i] for i in range(40)]
    a = run_segments(ex, utts)
    want = oracle_segments(cfg, params, utts[-3:])
    check_close(a[-3:], want)
    b = run_segments(ex, utts[-3:])
    np.testing.assert_allclose(a[-3:], b, atol=1e-5)


def test_full_size_properties():
    """BASELINE config 2 shape (batch 256 x 200 frames x 80-d): finite, deterministic, and equal to the
    same utterances run in smaller groups."""
    cfg, params, ex = model("res2net50_w24_s4_c32", 80)
    rng = np.random.default_rng(11)
    feats = net_oracle.synth_feats(rng, 256, 200, 80)
    utts = [feats[i] for i in range(256)]
    a = run_segments(ex, utts)
    b = run_segments(ex, utts)
    assert np.isfinite(a).all()
    np.testing.assert_array_equal(a, b)
    c = run_segments(ex, utts[:7])
    np.testing.assert_allclose(a[:7], c, atol=1e-5)
    want = oracle_segments(cfg, params, utts[:2])
    assert cosines(a[:2], want).min() >= COS_TOL


def test_cmvn_sliding_on_device_matches_host_restatement():
    """svx_cmvn_sliding vs the CMN oracle (Kaldi's SlidingWindowCmn rule pinned by hand-computed vectors in
    tests/test_oracle_cmn.py; reference tf_extract.py:63 pipe), windows shorter and longer than the utterance, the centred
    mode the reference uses and the causal mode with min_window, then end to end through extract(cmvn=True)."""
    cfg, params, ex = model("tdnn", 40)
    rng = np.random.default_rng(17)
    lens = [1, 25, 149, 300, 301, 777]
    utts = [(rng.standard_normal((t, 40)) * 3 + rng.standard_normal(40)).astype(np.float32) for t in lens]
    offs = np.zeros(len(utts) + 1, np.int32)
    offs[1:] = np.cumsum(lens)
    dev = torch.from_numpy(np.concatenate(utts, 0)).cuda()
    ex.cmvn_sliding(dev, offs)
    got = dev.cpu().numpy()
    want = np.concatenate([cmn_oracle.apply_cmvn_sliding(u) for u in utts], 0)
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-6)
    for window, min_window in ((300, 100), (100, 30)):          # causal windows: the first frames look ahead to min_window
        dev = torch.from_numpy(np.concatenate(utts, 0)).cuda()
        ex.cmvn_sliding(dev, offs, cmn_window=window, center=False, min_window=min_window)
        want = np.concatenate([cmn_oracle.apply_cmvn_sliding(u, window, False, min_window) for u in utts], 0)
        np.testing.assert_allclose(dev.cpu().numpy(), want, rtol=0, atol=2e-6)
    a = ex.extract(utts[1:], cmvn=True)
    b = ex.extract([cmn_oracle.apply_cmvn_sliding(u) for u in utts[1:]])
    np.testing.assert_allclose(a, b, atol=2e-4)


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
@pytest.mark.parametrize("n_utt,frames", [(3, 57), (40, 120), (256, 200)])
def test_fused_chain_equals_separate_convs(precision, n_utt, frames):
    """res2_chain.cu (the three hierarchical 3x3 convs of a stride-1 Res2Net block in one launch, running sums in shared memory)
    against the same convs as separate flat-kernel launches: same MMA order and rounding points, so the embeddings are
    bit-identical — from 3 short utterances (bands shorter than the pipeline) to the full BASELINE batch (225 tiles per CTA)."""
    cfg, params, ex = model("res2net50_w24_s4_c32", 80, precision)
    rng = np.random.default_rng(31)
    feats = net_oracle.synth_feats(rng, n_utt, frames, 80)
    utts = [feats[i] for i in range(n_utt)]
    ex.set_option("no_chain", 1)
    sep = run_segments(ex, utts)
    n_sep = ex.last_launches
    ex.set_option("no_chain", 0)
    fused = run_segments(ex, utts)
    assert ex.last_launches == n_sep - 2 * 3            # the three 24-channel blocks of stage 1: one launch instead of three, each
    np.testing.assert_array_equal(fused, sep)


@pytest.mark.parametrize("model_id,feat_dim,precision,n_utt,frames", [
    ("res2net50_w24_s4_c32", 80, "fp16", 5, 57), ("res2net50_w24_s4_c32", 80, "bf16", 64, 200), ("res2net50_w24_s4_c32", 80, "fp16", 256, 200),
    ("res2net50_w24_s4_c64", 40, "fp16", 9, 48), ("dpn68", 80, "fp16", 12, 120), ("tdnn", 40, "fp16", 33, 320)])
def test_pair_gemm_equals_flat_kernel(model_id, feat_dim, precision, n_utt, frames):
    """conv_pair.cu (deep 1x1 convs as a cta_group::2 GEMM, one M = 256 x N <= 256 tile per CTA pair) against the same convs on the
    flat kernel: same K order and epilogue expression, so the embeddings are bit-identical; from a handful of pixel blocks (fewer
    tiles than CTA pairs) to the full BASELINE batch.  The networks cover residual, planar-split and two-destination epilogues."""
    cfg, params, ex = model(model_id, feat_dim, precision)
    rng = np.random.default_rng(37)
    feats = net_oracle.synth_feats(rng, n_utt, frames, feat_dim)
    utts = [feats[i] for i in range(n_utt)]
    ex.set_option("no_pair_s2", 1)          # the stride-2 convs accumulate in a different order on the pair kernel: compared below with a tolerance
    ex.set_option("no_pair", 1)
    flat = run_segments(ex, utts)
    ex.set_option("no_pair", 0)
    paired = run_segments(ex, utts)
    ex.set_option("no_pair_s2", 0)
    np.testing.assert_array_equal(paired, flat)


@pytest.mark.parametrize("model_id,feat_dim,precision,lens", [
    ("res2net50_w24_s4_c32", 80, "fp16", [200, 57, 25, 26, 131]), ("res2net50_w24_s4_c32", 40, "bf16", [64, 33, 200]),
    ("res2net50_w24_s4_c64", 40, "fp16", [48, 99]), ("res2net50_w8_s6_c16", 80, "fp16", [40, 121])])
def test_pair_stride2_tiles_match_umma_kernel(model_id, feat_dim, precision, lens):
    """Stride-2 convs of the Res2Net down-sampling blocks on conv_pair.cu (2-D tiles over four parity-phase views, haloed boxes,
    shifted descriptors) against conv_umma.cu (one box per tap): same products, other summation order (phase-major instead of
    tap-major), so 16-bit roundings flip here and there through the 40 layers behind them — the two paths agree to 5e-3 relative L2
    and the pair path is as close to the fp32 oracle as the other one; odd and even segment lengths, several segments per call
    (gap rows, zero column)."""
    cfg, params, ex = model(model_id, feat_dim, precision)
    rng = np.random.default_rng(41)
    utts = [net_oracle.synth_feats(rng, 1, n, feat_dim)[0] for n in lens]
    ex.set_option("no_pair_s2", 1)
    ref = run_segments(ex, utts)
    n_ref = ex.last_launches
    ex.set_option("no_pair_s2", 0)
    got = run_segments(ex, utts)
    assert ex.last_launches == n_ref
    if precision == "bf16":     # 8-bit mantissa: every flipped rounding is 8x larger
        check_close(got, ref, cos_tol=0.9995, rel_tol=4e-2, norm_tol=1e-2)
    else:
        check_close(got, ref, cos_tol=0.99999, rel_tol=5e-3, norm_tol=1e-3)
    want = oracle_segments(cfg, params, utts)
    rel = lambda a: (np.linalg.norm(a - want, axis=1) / np.linalg.norm(want, axis=1)).max()
    assert rel(got) <= 1.25 * rel(ref) + (8e-3 if precision == "bf16" else 1e-3), (rel(got), rel(ref))
