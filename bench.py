#!/usr/bin/env python
"""bench.py — headline benchmark: Res2Net50 (w24_s4_c32) embeddings/s on 200-frame 80-d FBANK, batch 256 per GPU,
plus the AS-norm scoring job of BASELINE config 5 (trials/s), on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  torchrun --nnodes=1 --nproc-per-node N … bench.py --gpus N …        (N > 1, one rank per GPU)

A step = one pass of the extraction hot path over one batch of 256 synthetic utterances per GPU (weak scaling).
value    = embeddings/s with features already resident in HBM (device in, device out);
e2e      = the same through the public host API (pinned host features in, host embeddings out, copies timed);
roofline = the tcgen05 conv kernels (conv_flat_kernel for stride-1 layers, conv_umma_kernel for stride-2 layers): algorithmic conv
           FLOPs of their launches / their CUDA-event time, against the measured sustained 16-bit tensor peak of MEASURED_PEAKS.json;
           the "hbm" view divides the layer-wise algorithmic bytes (212 MB per utterance) by the same time; "traffic"
           is the DRAM bytes ncu counted for the same launches (profiles/r02_traffic.json);
cpu_baseline = the oracle port (PyTorch CPU fp32 restatement of the reference graph, batch 1 like tf_extract.py) on this box's
           host cores, on a bounded sample;
asnorm   = BASELINE config 5 (145 160 x 256 test rows vs 5994-speaker cohort, top-300, 579 818 trials) with its own roofline
           (tensor view of the 445.5-GFLOP cohort GEMM, HBM view of SURVEY §8d's 156 MB + 2080 B per trial) and cpu_baseline (the
           NumPy port of the reference's snorm.py functions, pinned to the reference by tests/golden/score);
bf16     = the headline step again with bf16 operands (the dtype BASELINE.json's north_star names; fp16 is the default because it
           meets the cosine >= 0.9999 bar);
configs  = the other BASELINE configs on the same library: C1 tdnn batch 64 + 1000 cosine trials, C3 dpn68 200-600 frames,
           C4 res2net200_w8_s6_c16 300-3000 frames — C3 / C4 on a FIXED utterance set, frame-balanced over the ranks, including the
           NCCL all-gather of the embeddings (strong scaling).
--impl reference times the CPU ports alone (TensorFlow 1.x, which the reference's extraction needs, cannot be installed).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MODEL_ID, FEAT_DIM, FRAMES, BATCH = "res2net50_w24_s4_c32", 80, 200, 256
METRIC = "Res2Net50 embeddings/s (200fr 80-d) at 1/2/4/8 B200; asnorm trials/s"
FLOPS_PER_UTT = 22.367e9            # 2 x conv/dense MACs, 200 frames x 80 bins
BYTES_PER_UTT = 212e6                # layer-wise fused activation traffic, 16-bit activations
SCORE_N, SCORE_C, SCORE_D, SCORE_TOPK, SCORE_TRIALS = 145160, 5994, 256, 300, 579818
SCORE_GEMM_FLOPS = 2.0 * SCORE_N * SCORE_C * SCORE_D                                   # 445.5 GFLOP
SCORE_BYTES = SCORE_N * SCORE_D * 4 + SCORE_C * SCORE_D * 4 + SCORE_N * 8 + 2080.0 * SCORE_TRIALS   # 156 MB + 2080 B per trial
# other BASELINE configs (SURVEY §8d algorithmic FLOPs)
C1 = dict(model="tdnn", feat_dim=40, frames=320, batch=64, trials=1000, flops_per_utt=1.7448e9)
C3 = dict(model="dpn68", feat_dim=80, lo=200, hi=600, n=1024, flops_per_frame=0.11848e9, flops_per_utt=8.5e6)
C4 = dict(model="res2net200_w8_s6_c16", feat_dim=80, lo=300, hi=3000, n=256, flops_per_frame=0.097573e9, flops_per_utt=3.9e6)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------- CPU legs (oracle ports)
def cpu_oracle_rate(seconds_budget: float = 12.0, threads=None):
    """Extraction oracle on the host cores, reference semantics (one utterance per graph evaluation, tf_extract.py:84)."""
    import numpy as np
    import torch
    from oracle import net_oracle
    from voxsrc2020_speaker_verification_b200 import arch
    if threads:
        torch.set_num_threads(threads)
    cfg = arch.get_config(MODEL_ID)
    params = net_oracle.init_params(cfg, FEAT_DIM, seed=4321, calib_frames=48, calib_batch=4)
    x = net_oracle.synth_feats(np.random.default_rng(1234), 64, FRAMES, FEAT_DIM)
    net_oracle.forward(cfg, params, x[:1])
    t0 = time.perf_counter()
    n = 0
    while n < 1024 and (time.perf_counter() - t0 < seconds_budget or n < 2):
        net_oracle.forward(cfg, params, x[n % 64:n % 64 + 1])
        n += 1
    dt = time.perf_counter() - t0
    return n / dt, n, dt, torch.get_num_threads()


def score_inputs(rng, n=SCORE_N, c=SCORE_C, d=SCORE_D, trials=SCORE_TRIALS):
    import numpy as np

    def unit(k):
        v = rng.standard_normal((k, d), dtype=np.float32)
        return v / np.linalg.norm(v, axis=1, keepdims=True)
    x = unit(n)
    cohort = ((unit(c) + unit(c) + unit(c)) / np.float32(3)).astype(np.float32)      # speaker means of 3 unit vectors: |c| ~ 0.58
    i1 = rng.integers(0, n, trials).astype(np.int32)
    i2 = rng.integers(0, n, trials).astype(np.int32)
    return x.astype(np.float32), cohort, i1, i2


def cpu_score_rate(rows: int = 6144, trials: int = 200000):
    """The NumPy port of the reference's scoring functions (oracle/score_oracle.py, pinned to snorm.py by tests/golden/score) on a
    bounded sample of config 5: `rows` test rows against the full cohort, `trials` trials.  The whole job is extrapolated linearly
    (both phases are linear in their unit): t_job = N / rows_per_s + T / trials_per_s."""
    import numpy as np
    from oracle import score_oracle
    rng = np.random.default_rng(99)
    x, cohort, i1, i2 = score_inputs(rng, n=rows, trials=trials)
    t0 = time.perf_counter()
    mean, std = score_oracle.cohort_mean_std_arrays(x, cohort, SCORE_TOPK)
    t1 = time.perf_counter()
    score_oracle.trial_scores_arrays(x, i1, i2, mean, std)
    t2 = time.perf_counter()
    rows_per_s, trials_per_s = rows / (t1 - t0), trials / (t2 - t1)
    t_job = SCORE_N / rows_per_s + SCORE_TRIALS / trials_per_s
    return {"value": SCORE_TRIALS / t_job, "unit": "trials/s", "cores": os.cpu_count() or 1, "kind": "port",
            "sample": "%d of %d test rows vs the full %d-speaker cohort (%.1f s) + %d of %d trials (%.1f s), NumPy port of snorm.py:83-131; "
                      "job time extrapolated linearly" % (rows, SCORE_N, SCORE_C, t1 - t0, trials, SCORE_TRIALS, t2 - t1),
            "rows_per_s": rows_per_s, "trials_per_s": trials_per_s, "job_seconds_extrapolated": t_job}


def run_reference(args, rank, world):
    if rank != 0:
        return
    import numpy as np
    import torch
    from oracle import net_oracle
    from voxsrc2020_speaker_verification_b200 import arch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = arch.get_config(MODEL_ID)
    params = net_oracle.init_params(cfg, FEAT_DIM, seed=4321, calib_frames=48, calib_batch=4)
    sample = 4                                       # utterances per step (bounded sample of the 256-utterance batch)
    x = net_oracle.synth_feats(np.random.default_rng(1234), sample, FRAMES, FEAT_DIM)
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        for i in range(sample):
            net_oracle.forward(cfg, params, x[i:i + 1])   # batch 1, as tf_extract.py:84
        if it >= args.warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    value = sample * len(times) / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "embeddings/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "res2net50_w24_s4_c32 80-d FBANK 200 frames; %d utterances per step at batch 1 on host cores "
                                   "(PyTorch-CPU port of the reference graph; TensorFlow 1.x is not installable here)" % sample},
            "cpu_baseline": {"value": value, "unit": "embeddings/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": "%d utterances x %d steps, batch 1" % (sample, len(times))},
            "e2e": {"value": value, "unit": "embeddings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if not args.no_scoring:
        line["asnorm"] = cpu_score_rate()
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-scoring", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1 / C3 / C4 legs and the bf16 headline")
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from oracle import net_oracle          # cpu_baseline leg + synthetic weights/features generator only
    from voxsrc2020_speaker_verification_b200 import arch, dist as svdist
    from voxsrc2020_speaker_verification_b200.extractor import Extractor
    from voxsrc2020_speaker_verification_b200.scoring import Scorer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200; there is no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pk = peaks()
    warm = max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        """ms for `steps` calls: CUDA events on the current stream, barrier + synchronize on both sides, max over ranks."""
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def conv_roofline(ex, run):
        """CUDA events around every tcgen05 conv launch of one pass (option time_convs): (ms, algorithmic FLOPs)."""
        ex.set_option("time_convs", 1)
        best, flops = None, 0.0
        for _ in range(3):
            run()
            torch.cuda.synchronize()
            m, flops = ex.conv_time()
            best = m if best is None else min(best, m)
        ex.set_option("time_convs", 0)
        return best, flops

    # ---- headline: res2net50_w24_s4_c32, batch 256 x 200 x 80 per GPU
    cfg = arch.get_config(MODEL_ID)
    params = net_oracle.init_params(cfg, FEAT_DIM, seed=4321, calib_frames=48, calib_batch=4)
    ex = Extractor(MODEL_ID, FEAT_DIM, device=local, precision=args.precision).load_params(params)
    feats_h = torch.from_numpy(net_oracle.synth_feats(np.random.default_rng(1234 + rank), BATCH, FRAMES, FEAT_DIM)
                               .reshape(BATCH * FRAMES, FEAT_DIM)).pin_memory()
    feats_d = feats_h.cuda(non_blocking=True)
    offs = (np.arange(BATCH + 1) * FRAMES).astype(np.int32)
    out_d = torch.empty((BATCH, ex.embed_dim), dtype=torch.float32, device="cuda")
    out_h = torch.empty((BATCH, ex.embed_dim), dtype=torch.float32).pin_memory()

    sampler = ClockSampler(local)
    sampler.start()
    ms_dev = timed(lambda: ex.extract_packed(feats_d, offs, out_d), args.steps, warm)
    clocks = sampler.stop()
    launches_per_step = ex.last_launches
    # end to end through the host API: pinned host features in, host embeddings out, every step
    ms_e2e = timed(lambda: ex.extract_packed(feats_h, offs, out_h), args.steps, warm)
    conv_ms_step, conv_flops = conv_roofline(ex, lambda: ex.extract_packed(feats_d, offs, out_d))
    achieved = conv_flops / (conv_ms_step * 1e-3) / 1e12 if conv_ms_step > 0 else 0.0
    hbm_achieved = BATCH * BYTES_PER_UTT / (conv_ms_step * 1e-3) / 1e9 if conv_ms_step > 0 else 0.0
    traffic = None
    for name in ("r02_traffic.json", "r01_traffic.json"):
        tp = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get("conv_dram_bytes_per_step")
            break

    # ---- the same step with bf16 operands
    bf16 = None
    if not args.no_configs and args.precision == "fp16":
        exb = Extractor(MODEL_ID, FEAT_DIM, device=local, precision="bf16").load_params(params)
        ms_b = timed(lambda: exb.extract_packed(feats_d, offs, out_d), max(3, args.steps // 2), 3) / max(3, args.steps // 2)
        cms, cfl = conv_roofline(exb, lambda: exb.extract_packed(feats_d, offs, out_d))
        bf16 = {"value": world * BATCH / (ms_b * 1e-3), "unit": "embeddings/s", "ms_per_step": ms_b, "dtype": "bf16",
                "conv_tflops": cfl / (cms * 1e-3) / 1e12, "frac_of_sustained_peak": cfl / (cms * 1e-3) / 1e12 / pk["bf16_tflops_sustained"],
                "note": "bf16 operands run the same kernels at the same rate; they miss cosine >= 0.9999 on the synthetic weights "
                        "(0.9994-0.9997, profiles/r02_parity_probe.txt), which is why fp16 is the default"}
        del exb

    # ---- other BASELINE configs
    configs = None
    if not args.no_configs:
        configs = {}
        # C1: tdnn, batch 64 x 320 x 40 per GPU + 1000 cosine trials among the 64 embeddings (replicas, like the headline)
        c = C1
        cf1 = arch.get_config(c["model"])
        ex1 = Extractor(c["model"], c["feat_dim"], device=local, precision=args.precision).load_params(
            net_oracle.init_params(cf1, c["feat_dim"], seed=4321, calib_frames=64, calib_batch=4))
        f1 = torch.from_numpy(net_oracle.synth_feats(np.random.default_rng(7 + rank), c["batch"], c["frames"], c["feat_dim"])
                              .reshape(-1, c["feat_dim"])).cuda()
        o1 = (np.arange(c["batch"] + 1) * c["frames"]).astype(np.int32)
        e1 = torch.empty((c["batch"], ex1.embed_dim), dtype=torch.float32, device="cuda")
        sc1 = Scorer(local)
        rng1 = np.random.default_rng(8)
        t1a = torch.from_numpy(rng1.integers(0, c["batch"], c["trials"]).astype(np.int32)).cuda()
        t1b = torch.from_numpy(rng1.integers(0, c["batch"], c["trials"]).astype(np.int32)).cuda()

        def c1_step():
            ex1.extract_packed(f1, o1, e1)
            sc1.trial_scores(sc1.l2norm(e1), t1a, t1b)
        ms1 = timed(c1_step, 20, 5) / 20
        cms, cfl = conv_roofline(ex1, lambda: ex1.extract_packed(f1, o1, e1))
        configs["C1"] = {"workload": "tdnn 40-d, 320 frames, batch %d per GPU + %d cosine trials (BASELINE configs[0])" % (c["batch"], c["trials"]),
                         "value": world * c["batch"] / (ms1 * 1e-3), "unit": "embeddings/s", "ms_per_step": ms1, "scaling": "weak",
                         "tflops": world * c["batch"] * c["flops_per_utt"] / (ms1 * 1e-3) / 1e12,
                         "frac_of_sustained_peak": c["batch"] * c["flops_per_utt"] / (ms1 * 1e-3) / 1e12 / pk["bf16_tflops_sustained"],
                         "conv_tflops": cfl / (cms * 1e-3) / 1e12,
                         "note": "0.11 GFLOP per launch sequence: launch-latency bound at this batch"}
        del ex1

        def strong_leg(c, tag):
            """Fixed utterance set, frame-balanced over the ranks, device-resident features, NCCL all-gather of the embeddings."""
            cf = arch.get_config(c["model"])
            exs = Extractor(c["model"], c["feat_dim"], device=local, precision=args.precision).load_params(
                net_oracle.init_params(cf, c["feat_dim"], seed=4321, calib_frames=48, calib_batch=2))
            rng = np.random.default_rng(31)
            lens = rng.integers(c["lo"], c["hi"] + 1, c["n"]).tolist()
            parts = svdist.balance_by_frames(lens, world)
            mine = parts[rank]
            tot = int(sum(lens[i] for i in mine))
            base = net_oracle.synth_feats(np.random.default_rng(32 + rank), 1, min(tot, 40000), c["feat_dim"])[0]
            fs = torch.from_numpy(np.resize(base, (tot, c["feat_dim"]))).cuda()          # synthetic frames, tiled to the share's size
            of = np.zeros(len(mine) + 1, np.int32)
            of[1:] = np.cumsum([lens[i] for i in mine])
            loc = torch.empty((len(mine), exs.embed_dim), dtype=torch.float32, device="cuda")

            def step():
                exs.extract_packed(fs, of, loc)
                if world > 1:
                    svdist.gather_sharded(loc, parts)
            steps = 3
            ms = timed(step, steps, 2) / steps
            cms, cfl = conv_roofline(exs, lambda: exs.extract_packed(fs, of, loc))
            frames = float(sum(lens))
            flops = frames * c["flops_per_frame"] + len(lens) * c["flops_per_utt"]
            layerwise_note = {}
            return {"workload": "%s 80-d, %d utterances of %d-%d frames (%.0f k frames, fixed set), frame-balanced over %d GPU(s)%s (BASELINE %s)"
                                % (c["model"], c["n"], c["lo"], c["hi"], frames / 1e3, world,
                                   " + NCCL all-gather of the embeddings" if world > 1 else "", tag),
                    "value": len(lens) / (ms * 1e-3), "unit": "embeddings/s", "frames_per_s": frames / (ms * 1e-3), "ms_per_pass": ms,
                    "scaling": "strong", "tflops": flops / (ms * 1e-3) / 1e12,
                    "frac_of_sustained_peak": flops / (ms * 1e-3) / 1e12 / (world * pk["bf16_tflops_sustained"]),
                    "conv_tflops_rank0": cfl / (cms * 1e-3) / 1e12, "launches_per_pass": int(exs.last_launches), **layerwise_note}
        configs["C3"] = strong_leg(C3, "configs[2]")
        configs["C4"] = strong_leg(C4, "configs[3]")

    # ---- scoring job (BASELINE config 5).  N > 1: the test rows are sharded (cohort replicated, all-gather of two floats per row)
    # — the layout that scales; the cohort-row-sharded layout the north star names (NCCL exchange of top-k candidates) is timed
    # beside it.
    score = None
    if not args.no_scoring:
        rng = np.random.default_rng(99)
        sc = Scorer(local)
        xh, ch, i1h, i2h = score_inputs(rng)
        x, cohort = torch.from_numpy(xh).cuda(), torch.from_numpy(ch).cuda()
        i1, i2 = torch.from_numpy(i1h).cuda(), torch.from_numpy(i2h).cuda()
        lo, hi = rank * SCORE_TRIALS // world, (rank + 1) * SCORE_TRIALS // world
        stats_ms = {}

        def job(stats_fn):
            def f():
                mean, std = stats_fn()
                sc.trial_scores(x, i1[lo:hi], i2[lo:hi], mean, std)
            return f
        reps = 5
        if world > 1:
            ms_rows = timed(job(lambda: svdist.rows_sharded_cohort_mean_std(sc, x, cohort, SCORE_TOPK)), reps, 2) / reps
            ms_coh = timed(job(lambda: svdist.sharded_cohort_mean_std(sc, x, cohort, SCORE_TOPK)), reps, 2) / reps
            ms_job = ms_rows
        else:
            ms_job = timed(job(lambda: sc.cohort_mean_std(x, cohort, SCORE_TOPK)), reps, 2) / reps
        ms_stats = timed(lambda: sc.cohort_mean_std(x, cohort, SCORE_TOPK) if world == 1
                         else svdist.rows_sharded_cohort_mean_std(sc, x, cohort, SCORE_TOPK), reps, 1) / reps
        fused_rows, handed_back = sc.last_path()
        sc.set_option("fused", 0)
        ms_unfused = timed(lambda: sc.cohort_mean_std(x, cohort, SCORE_TOPK), 2, 1) / 2 if world == 1 else None
        sc.set_option("fused", 1)
        gemm_tf = SCORE_GEMM_FLOPS / (ms_stats * 1e-3) / 1e12
        score = {"value": SCORE_TRIALS / (ms_job * 1e-3), "unit": "trials/s", "ms_per_job": ms_job, "ms_statistics": ms_stats,
                 "workload": "%d x %d-d test rows vs %d-speaker cohort, top-%d, %d trials%s" %
                             (SCORE_N, SCORE_D, SCORE_C, SCORE_TOPK, SCORE_TRIALS,
                              "; test rows sharded over the ranks, cohort replicated, all-gather of the statistics" if world > 1 else ""),
                 "kernel": "asnorm_fused_kernel (split-bf16 tcgen05 GEMM, test operand in tensor memory, one-term histogram pass + exact "
                           "selection pass in the epilogue, scores never written) + trial_scores_kernel",
                 "rows_finished_by_fused_kernel": fused_rows, "rows_handed_to_unfused_kernels": handed_back,
                 "roofline": {"bound": "tensor", "achieved": gemm_tf, "peak": world * pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                              "frac": gemm_tf / (world * pk["bf16_tflops_sustained"]), "traffic": None,
                              "note": "algorithmic FLOPs of ONE fp32-equivalent cohort GEMM (445.5 GFLOP) over the statistics time; the kernel "
                                      "issues 4x that (histogram pass on the xh.ch term alone, exact pass on the three split-bf16 terms)",
                              "issued_tflops": 4.0 * gemm_tf,
                              "hbm": {"bound": "hbm", "achieved": SCORE_BYTES / (ms_job * 1e-3) / 1e9, "peak": world * pk["hbm_gbs"], "unit": "GB/s",
                                      "frac": SCORE_BYTES / (ms_job * 1e-3) / 1e9 / (world * pk["hbm_gbs"]),
                                      "algorithmic_bytes_per_job": SCORE_BYTES,
                                      "note": "SURVEY §8d: 156 MB of operands/statistics + 2080 B per trial; the statistics phase is tensor/"
                                              "ingest-bound (arithmetic intensity ~2900 FLOP/B), only the trial gather is HBM/L2-bound"}}}
        if ms_unfused is not None:
            score["unfused_ms_statistics"] = ms_unfused
        if world > 1:
            score["cohort_rows_sharded"] = {"value": SCORE_TRIALS / (ms_coh * 1e-3), "unit": "trials/s", "ms_per_job": ms_coh,
                                            "note": "the layout BASELINE.json's north_star names: cohort rows sharded, NCCL all-to-all of per-rank "
                                                    "top-k candidates, merge on the owner of each test row — every rank still streams all test rows"}
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            score["cpu_baseline"] = cpu_score_rate()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, n, dt, threads = cpu_oracle_rate(12.0)
        cpu = {"value": rate, "unit": "embeddings/s", "cores": threads, "kind": "port",
               "sample": "%d utterances of 200 frames at batch 1 (%.1f s)" % (n, dt)}

    if rank == 0:
        step_ms = ms_dev / args.steps
        e2e_ms = ms_e2e / args.steps
        line = {
            "metric": METRIC, "value": world * BATCH / (step_ms * 1e-3), "unit": "embeddings/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": "res2net50_w24_s4_c32 embedding extraction, 80-dim FBANK, 200 frames, batch %d per GPU "
                                   "(BASELINE configs[1]); random-init calibrated weights" % BATCH,
                       "l2": "activations of one step (>5 GB) exceed the 126 MB L2; no explicit flush needed",
                       "parallelism": "utterance-sharded replicas x%d" % world},
            "clocks": clocks,
            "e2e": {"value": world * BATCH / (e2e_ms * 1e-3), "unit": "embeddings/s",
                    "h2d_bytes_per_step": BATCH * FRAMES * FEAT_DIM * 4 + (BATCH + 1) * 4, "d2h_bytes_per_step": BATCH * ex.embed_dim * 4},
            "gpu_launches": int(launches_per_step) * args.steps,
            "roofline": {"bound": "tensor", "kernel": "conv_flat_kernel + conv_pair_kernel + res2_chain_kernel + conv_umma_kernel (all tcgen05 conv launches of one step)",
                         "achieved": achieved, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                         "frac": achieved / pk["bf16_tflops_sustained"], "traffic": traffic,
                         "peak_source": pk["source"] + " sustained (kernels timed inside a long step)",
                         "conv_ms_per_step": conv_ms_step, "conv_share_of_step": conv_ms_step / step_ms,
                         "whole_step_tflops": BATCH * FLOPS_PER_UTT / (step_ms * 1e-3) / 1e12,
                         "hbm": {"bound": "hbm", "achieved": hbm_achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                 "frac": hbm_achieved / pk["hbm_gbs"], "algorithmic_bytes_per_step": BATCH * BYTES_PER_UTT,
                                 "note": "layer-wise execution is HBM-bound (arithmetic intensity 106 FLOP/B < ridge ~210)"}},
            "cpu_baseline": cpu,
            "asnorm": score,
            "bf16": bf16,
            "configs": configs,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
