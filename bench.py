#!/usr/bin/env python
"""bench.py — headline benchmark: Res2Net50 (w24_s4_c32) embeddings/s on 200-frame 80-d FBANK, batch 256 per GPU,
plus the AS-norm scoring job of BASELINE config 5 (trials/s), on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  torchrun --nnodes=1 --nproc-per-node N … bench.py --gpus N …        (N > 1, one rank per GPU)

A step = one pass of the extraction hot path over one batch of 256 synthetic utterances per GPU (weak scaling).
value  = embeddings/s with features already resident in HBM (device in, device out);
e2e    = the same through the public host API (pinned host features in, host embeddings out, copies timed);
roofline = the tcgen05 conv kernels (conv_flat_kernel for stride-1 layers, conv_umma_kernel for stride-2 layers): algorithmic
           conv FLOPs of their launches / their CUDA-event time, against the measured sustained 16-bit tensor peak of
           MEASURED_PEAKS.json; the "hbm" view divides the layer-wise algorithmic bytes (212 MB per utterance)
           by the same time; "traffic" is the DRAM bytes ncu counted for the same launches (profiles/r01_traffic.json);
cpu_baseline = the oracle port (PyTorch CPU fp32 restatement of the reference graph, batch 1 like tf_extract.py)
           on this box's host cores, on a bounded sample.
--impl reference times that CPU port alone (TensorFlow 1.x, which the reference needs, cannot be installed).
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


def cpu_oracle_rate(seconds_budget: float = 20.0, threads=None):
    """Oracle port on the host cores, reference semantics (one utterance per graph evaluation)."""
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
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-scoring", action="store_true")
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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pk = peaks()
    cfg = arch.get_config(MODEL_ID)
    params = net_oracle.init_params(cfg, FEAT_DIM, seed=4321, calib_frames=48, calib_batch=4)
    ex = Extractor(MODEL_ID, FEAT_DIM, device=local, precision=args.precision).load_params(params)
    feats_h = torch.from_numpy(net_oracle.synth_feats(np.random.default_rng(1234 + rank), BATCH, FRAMES, FEAT_DIM)
                               .reshape(BATCH * FRAMES, FEAT_DIM)).pin_memory()
    feats_d = feats_h.cuda(non_blocking=True)
    offs = (np.arange(BATCH + 1) * FRAMES).astype(np.int32)
    out_d = torch.empty((BATCH, ex.embed_dim), dtype=torch.float32, device="cuda")
    out_h = torch.empty((BATCH, ex.embed_dim), dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
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

    sampler = ClockSampler(local)
    # ---- resident-input throughput (value)
    sampler.start()
    ms_dev = timed(lambda: ex.extract_packed(feats_d, offs, out_d), args.steps, max(args.warmup, 3))
    clocks = sampler.stop()
    launches_per_step = ex.last_launches
    # ---- end to end through the host API (e2e): pinned host features in, host embeddings out, every step
    ms_e2e = timed(lambda: ex.extract_packed(feats_h, offs, out_h), args.steps, max(args.warmup, 3))
    # ---- conv kernel roofline: CUDA events around every tcgen05 conv launch inside a timed step
    ex.set_option("time_convs", 1)
    conv_ms, conv_flops = [], 0.0
    for _ in range(3):
        ex.extract_packed(feats_d, offs, out_d)
        torch.cuda.synchronize()
        m, conv_flops = ex.conv_time()
        conv_ms.append(m)
    ex.set_option("time_convs", 0)
    conv_ms_step = min(conv_ms)
    achieved = conv_flops / (conv_ms_step * 1e-3) / 1e12 if conv_ms_step > 0 else 0.0
    hbm_achieved = BATCH * BYTES_PER_UTT / (conv_ms_step * 1e-3) / 1e9 if conv_ms_step > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get("conv_dram_bytes_per_step")

    # ---- scoring job (BASELINE config 5), test rows sharded across ranks' replicas is the natural layout; the
    # cohort-row-sharded layout with an NCCL all-gather of top-k candidates is svdist.sharded_cohort_mean_std
    score = None
    if not args.no_scoring:
        rng = np.random.default_rng(99)
        sc = Scorer(local)
        x = torch.from_numpy(rng.standard_normal((SCORE_N, SCORE_D), dtype=np.float32)).cuda()
        x = sc.l2norm(x)
        c3 = torch.from_numpy(rng.standard_normal((3, SCORE_C, SCORE_D), dtype=np.float32)).cuda()
        cohort = (sc.l2norm(c3[0]) + sc.l2norm(c3[1]) + sc.l2norm(c3[2])) / 3.0
        i1 = torch.from_numpy(rng.integers(0, SCORE_N, SCORE_TRIALS).astype(np.int32)).cuda()
        i2 = torch.from_numpy(rng.integers(0, SCORE_N, SCORE_TRIALS).astype(np.int32)).cuda()

        def score_step():
            if world > 1:
                mean, std = svdist.sharded_cohort_mean_std(sc, x, cohort, SCORE_TOPK)
            else:
                mean, std = sc.cohort_mean_std(x, cohort, SCORE_TOPK)
            lo = rank * SCORE_TRIALS // world
            hi = (rank + 1) * SCORE_TRIALS // world
            sc.trial_scores(x, i1[lo:hi], i2[lo:hi], mean, std)
        ms_score = timed(score_step, 3, 2)
        ms_score_rows = None
        if world > 1:       # the natural layout beside it: test rows sharded, cohort replicated, all-gather of [n, 2] statistics
            def score_step_rows():
                mean, std = svdist.rows_sharded_cohort_mean_std(sc, x, cohort, SCORE_TOPK)
                lo = rank * SCORE_TRIALS // world
                hi = (rank + 1) * SCORE_TRIALS // world
                sc.trial_scores(x, i1[lo:hi], i2[lo:hi], mean, std)
            ms_score_rows = timed(score_step_rows, 3, 2)
        score = {"value": SCORE_TRIALS / (ms_score / 3 * 1e-3), "unit": "trials/s", "ms_per_job": ms_score / 3,
                 "workload": "%d x %d-d test rows vs %d-speaker cohort, top-%d, %d trials%s" %
                             (SCORE_N, SCORE_D, SCORE_C, SCORE_TOPK, SCORE_TRIALS,
                              "; cohort row-sharded, NCCL all-to-all of per-rank top-k candidates, merge on the owner of each test row" if world > 1 else "")}
        if ms_score_rows is not None:
            score["test_rows_sharded"] = {"value": SCORE_TRIALS / (ms_score_rows / 3 * 1e-3), "unit": "trials/s", "ms_per_job": ms_score_rows / 3,
                                          "note": "natural layout: test rows sharded, cohort replicated, all-gather of the statistics"}

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
            "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": "res2net50_w24_s4_c32 embedding extraction, 80-dim FBANK, 200 frames, batch %d per GPU "
                                   "(BASELINE configs[1]); random-init calibrated weights" % BATCH,
                       "l2": "activations of one step (>5 GB) exceed the 126 MB L2; no explicit flush needed",
                       "parallelism": "utterance-sharded replicas x%d" % world},
            "clocks": clocks,
            "e2e": {"value": world * BATCH / (e2e_ms * 1e-3), "unit": "embeddings/s",
                    "h2d_bytes_per_step": BATCH * FRAMES * FEAT_DIM * 4 + (BATCH + 1) * 4, "d2h_bytes_per_step": BATCH * ex.embed_dim * 4},
            "gpu_launches": int(launches_per_step) * args.steps,
            "roofline": {"bound": "tensor", "kernel": "conv_flat_kernel + conv_umma_kernel (all tcgen05 conv launches of one step)",
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
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
