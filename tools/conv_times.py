"""Per-conv-launch device times of one extraction pass (debug library: SVX_LIB=.../libsvx_dbg.so SVX_CONV_TIMES=1), plus the
step time.  Usage: python tools/conv_times.py [--model M --feat-dim F --frames T --batch B --precision fp16]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import net_oracle  # synthetic weights / features generator only
from voxsrc2020_speaker_verification_b200 import arch
from voxsrc2020_speaker_verification_b200.extractor import Extractor

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="res2net50_w24_s4_c32")
ap.add_argument("--feat-dim", type=int, default=80)
ap.add_argument("--frames", type=int, default=200)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--precision", default="fp16")
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
cfg = arch.get_config(a.model)
params = net_oracle.init_params(cfg, a.feat_dim, seed=4321, calib_frames=48, calib_batch=4)
ex = Extractor(a.model, a.feat_dim, precision=a.precision).load_params(params)
feats = torch.from_numpy(net_oracle.synth_feats(np.random.default_rng(0), a.batch, a.frames, a.feat_dim).reshape(-1, a.feat_dim)).cuda()
offs = (np.arange(a.batch + 1) * a.frames).astype(np.int32)
out = torch.empty((a.batch, ex.embed_dim), device="cuda")
for _ in range(3):
    ex.extract_packed(feats, offs, out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    ex.extract_packed(feats, offs, out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
print("step %.3f ms  %.0f emb/s" % (ms, a.batch / ms * 1e3))
ex.set_option("time_convs", 1)
ex.extract_packed(feats, offs, out)
torch.cuda.synchronize()
m, fl = ex.conv_time()
print("conv %.3f ms  %.1f TFLOP/s" % (m, fl / m / 1e9))
