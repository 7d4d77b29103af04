"""The network oracle (oracle/net_oracle.py).  The reference pins nothing for this path (no TF here, no golden
vectors), so the restatement rules themselves are tested against naive NumPy loops, the architecture tables
against the reference README's parameter counts, and the outputs against a committed regression pin."""
import os

import numpy as np
import pytest
import torch

from oracle import net_oracle as no
from voxsrc2020_speaker_verification_b200 import arch


def naive_conv_nhwc(x, k, stride, pads, dil=(1, 1), groups=1):
    """x [H,W,C], k [kh,kw,C/g,O] (HWIO), explicit pads ((pt,pb),(pl,pr)) → [Ho,Wo,O] by direct loops."""
    x = np.pad(x, (pads[0], pads[1], (0, 0)))
    kh, kw, cg, o = k.shape
    H, W, _ = x.shape
    ho = (H - (kh - 1) * dil[0] - 1) // stride + 1
    wo = (W - (kw - 1) * dil[1] - 1) // stride + 1
    og = o // groups
    out = np.zeros((ho, wo, o), np.float64)
    for i in range(ho):
        for j in range(wo):
            for oc in range(o):
                g = oc // og
                acc = 0.0
                for r in range(kh):
                    for s in range(kw):
                        acc += np.dot(x[i * stride + r * dil[0], j * stride + s * dil[1], g * cg:(g + 1) * cg].astype(np.float64),
                                      k[r, s, :, oc].astype(np.float64))
                out[i, j, oc] = acc
    return out


def torch_conv(x_hwc, k, stride, padding, dil=(1, 1), groups=1):
    x = torch.from_numpy(x_hwc).permute(2, 0, 1)[None]
    y = no._conv_raw(x, torch.from_numpy(k), stride, padding, dil, groups)
    return y[0].permute(1, 2, 0).numpy()


@pytest.mark.parametrize("H,W", [(6, 8), (7, 5)])
def test_same_stride2_padding_rule(H, W):
    """TF SAME, stride 2, 3x3: pad (0,1) on even extents, (1,1) on odd ones (DPN, dpn_model.py:43)."""
    rng = np.random.default_rng(H * 10 + W)
    x = rng.standard_normal((H, W, 4)).astype(np.float32)
    k = rng.standard_normal((3, 3, 2, 4)).astype(np.float32)
    pads = tuple((0, 1) if n % 2 == 0 else (1, 1) for n in (H, W))
    np.testing.assert_allclose(torch_conv(x, k, 2, "SAME", groups=2), naive_conv_nhwc(x, k, 2, pads, groups=2), atol=1e-5)


def test_fixed_padding_stride2_and_projection_sampling():
    """Res2Net: stride-2 3x3 = pad (1,1) + VALID whatever the extent; 1x1 stride 2 samples even indices."""
    rng = np.random.default_rng(3)
    for H, W in ((6, 8), (7, 5)):
        x = rng.standard_normal((H, W, 3)).astype(np.float32)
        k = rng.standard_normal((3, 3, 3, 2)).astype(np.float32)
        xt = no.fixed_padding(torch.from_numpy(x).permute(2, 0, 1)[None], 3)
        got = no._conv_raw(xt, torch.from_numpy(k), 2, "VALID")[0].permute(1, 2, 0).numpy()
        np.testing.assert_allclose(got, naive_conv_nhwc(x, k, 2, ((1, 1), (1, 1))), atol=1e-5)
        k1 = rng.standard_normal((1, 1, 3, 2)).astype(np.float32)
        got1 = torch_conv(x, k1, 2, "VALID")
        np.testing.assert_allclose(got1, x[::2, ::2] @ k1[0, 0], atol=1e-5)


def test_dilated_same_conv_tdnn():
    rng = np.random.default_rng(4)
    x = rng.standard_normal((20, 1, 6)).astype(np.float32)
    for k_, d in ((5, 1), (3, 2), (3, 3)):
        k = rng.standard_normal((k_, 1, 6, 4)).astype(np.float32)
        p = d * (k_ - 1) // 2
        np.testing.assert_allclose(torch_conv(x, k, 1, "SAME", (d, 1)), naive_conv_nhwc(x, k, 1, ((p, p), (0, 0)), (d, 1)), atol=1e-5)


def test_avgpool_divisor_and_stats_pool_layout():
    x = torch.arange(2 * 5 * 4, dtype=torch.float32).reshape(1, 2, 5, 4)
    xp = no.fixed_padding(x, 3)
    y = torch.nn.functional.avg_pool2d(xp, 3, 2, 0)
    assert y.shape == (1, 2, 3, 2)
    np.testing.assert_allclose(y[0, 0, 0, 0].item(), (0 + 1 + 4 + 5) / 9.0)          # divisor 9 including the zero pad
    t = torch.randn(2, 3, 7, 4)
    p = no.flatten_nhwc(no.stats_pool(t)).numpy()                                      # index = w*2C + {c | C+c}
    tn = t.numpy()
    for w in range(4):
        np.testing.assert_allclose(p[:, w * 6:w * 6 + 3], tn[:, :, :, w].mean(2), atol=1e-6)
        np.testing.assert_allclose(p[:, w * 6 + 3:w * 6 + 6], np.sqrt(tn[:, :, :, w].var(2) + 1e-5), atol=1e-6)


def test_parameter_counts_match_reference_readme():
    """README.md:184-185,198-199,239-244 (printed to 0.1 M)."""
    want = {("tdnn", 40): 3.5, ("dpn68", 40): 13.9, ("res2net50_w24_s4_c64", 40): 26.9, ("res2net50_w24_s4_c32", 80): 17.7,
            ("res2net50_w8_s6_c16", 80): 4.8}
    for (mid, fd), m in want.items():
        assert abs(arch.param_count(arch.get_config(mid), fd) / 1e6 - m) < 0.1, (mid, fd)     # printed to 0.1 M; c64@40 is 26.83 M vs the README 26.9
    assert arch.param_count(arch.get_config("tdnn"), 40) == 3510272


def test_variable_names_follow_tf_numbering():
    gv = arch.enumerate_variables(arch.get_config("res2net50_w24_s4_c32"), 80)
    names = gv.names()
    assert names[0] == "conv2d/kernel" and names[1] == "batch_normalization/moving_mean"
    assert "conv2d_52/kernel" in names and "conv2d_53/kernel" not in names          # conv2d … conv2d_52
    assert "batch_normalization_38/moving_mean" in names and "batch_normalization_39/moving_mean" not in names
    assert "conv2d_3/batch_normalization_2/moving_variance" in names                  # BNs inside the hierarchical scope
    assert gv.shapes()["conv2d_3/kernel"] == (3, 3, 24, 72)
    d = arch.enumerate_variables(arch.get_config("dpn68"), 80).names()
    assert "conv2d_70/kernel" in d and "conv2d_71/kernel" not in d and "batch_normalization_73/moving_mean" in d
    t = arch.enumerate_variables(arch.get_config("tdnn"), 40).names()
    assert "conv2d_4/kernel" in t and "batch_normalization_6/moving_mean" in t


def test_chunk_plan_edge_cases():
    assert arch.chunk_plan(24) == []
    assert arch.chunk_plan(25) == [(0, 25)]
    assert arch.chunk_plan(1000) == [(0, 1000)]
    assert arch.chunk_plan(1024) == [(0, 1000)]                   # 24-frame tail dropped
    assert arch.chunk_plan(1025) == [(0, 1000), (1000, 25)]
    assert arch.chunk_plan(2500) == [(0, 1000), (1000, 1000), (2000, 500)]


@pytest.mark.parametrize("model_id", ["tdnn", "res2net50_w8_s6_c16", "dpn68"])
def test_regression_pin(golden_dir, model_id):
    g = np.load(os.path.join(golden_dir, "net", model_id + ".npz"))
    cfg = arch.get_config(model_id)
    params = no.init_params(cfg, int(g["feat_dim"]), seed=4321)
    x = no.synth_feats(np.random.default_rng(1234), 2, int(g["frames"]), int(g["feat_dim"]))
    y = no.forward(cfg, params, x)
    cos = (y * g["emb"]).sum(1) / np.linalg.norm(y, axis=1) / np.linalg.norm(g["emb"], axis=1)
    assert cos.min() > 0.99999


def test_batch_equals_single_and_extract_utterance():
    cfg = arch.get_config("tdnn")
    params = no.init_params(cfg, 40, seed=1)
    x = no.synth_feats(np.random.default_rng(0), 3, 40, 40)
    np.testing.assert_allclose(no.forward(cfg, params, x), np.concatenate([no.forward(cfg, params, x[i:i + 1]) for i in range(3)]), atol=2e-5)
    long = no.synth_feats(np.random.default_rng(1), 1, 1030, 40)[0]
    e = no.extract_utterance(cfg, params, long)
    a, b = no.forward(cfg, params, long[None, :1000])[0], no.forward(cfg, params, long[None, 1000:1030])[0]
    np.testing.assert_allclose(e, (a * 1000 + b * 30) / 1030, atol=1e-5)
    with pytest.raises(ZeroDivisionError):
        no.extract_utterance(cfg, params, long[:24])


def test_att_stats_pool_matches_naive_loops():
    """reference models.py:273-303 restated with explicit loops: concat(x, tiled mean, tiled std) -> 1x1 -> tanh -> 1x1 ->
    softmax over time -> weighted mean / sqrt(weighted second moment - mean^2 + eps)."""
    import torch
    from oracle import net_oracle
    rng = np.random.default_rng(5)
    n, c, h, w, a = 2, 6, 7, 3, 4
    x = rng.standard_normal((n, h, w, c)).astype(np.float32)            # NHWC like the reference
    k1 = rng.standard_normal((1, 1, 3 * c, a)).astype(np.float32)
    k2 = rng.standard_normal((1, 1, a, c)).astype(np.float32)
    ctx = net_oracle._Ctx({"att_stats_pool/conv2d/kernel": k1, "att_stats_pool/conv2d_1/kernel": k2})
    got = net_oracle.att_stats_pool(ctx, torch.from_numpy(x).permute(0, 3, 1, 2), a).permute(0, 2, 3, 1).numpy()   # [n,1,w,2c]
    want = np.zeros((n, 1, w, 2 * c), np.float64)
    for i in range(n):
        for j in range(w):
            xs = x[i, :, j, :].astype(np.float64)                       # [h, c]
            mean, std = xs.mean(0), np.sqrt(xs.var(0) + 1e-5)
            logits = np.zeros((h, c))
            for t in range(h):
                att_in = np.concatenate([xs[t], mean, std])
                logits[t] = np.tanh(att_in @ k1[0, 0].astype(np.float64)) @ k2[0, 0].astype(np.float64)
            wts = np.exp(logits - logits.max(0))
            wts /= wts.sum(0)
            wm = (xs * wts).sum(0)
            wss = (xs * xs * wts).sum(0)
            want[i, 0, j, :c] = wm
            want[i, 0, j, c:] = np.sqrt(wss - wm * wm + 1e-5)
    np.testing.assert_allclose(got, want, rtol=2e-4, atol=2e-5)


def test_att_models_enumerate_attention_kernels():
    from voxsrc2020_speaker_verification_b200 import arch
    cfg = arch.get_config("res2net200_w24_s4_c32_att")
    shapes = arch.enumerate_variables(cfg, 80).shapes()
    assert shapes["att_stats_pool/conv2d/kernel"] == (1, 1, 3 * 1024, 128)
    assert shapes["att_stats_pool/conv2d_1/kernel"] == (1, 1, 128, 1024)
    names = arch.enumerate_variables(cfg, 80).names()
    assert names.index("att_stats_pool/conv2d_1/kernel") < names.index("dense/kernel")
