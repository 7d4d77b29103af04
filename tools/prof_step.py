"""One or more plain extraction passes of the headline workload, for ncu captures (no timing is reported)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voxsrc2020_speaker_verification_b200 import lib as _lib
if os.environ.get("SVX_LIB"):
    _lib.LIB_PATH = os.environ["SVX_LIB"]
from oracle import net_oracle  # synthetic weights / features generator only
from voxsrc2020_speaker_verification_b200 import arch
from voxsrc2020_speaker_verification_b200.extractor import Extractor

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="res2net50_w24_s4_c32")
ap.add_argument("--feat-dim", type=int, default=80)
ap.add_argument("--frames", type=int, default=200)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--passes", type=int, default=2)
a = ap.parse_args()
cfg = arch.get_config(a.model)
params = net_oracle.init_params(cfg, a.feat_dim, seed=4321, calib_frames=48, calib_batch=4)
ex = Extractor(a.model, a.feat_dim).load_params(params)
feats = torch.from_numpy(net_oracle.synth_feats(np.random.default_rng(0), a.batch, a.frames, a.feat_dim).reshape(-1, a.feat_dim)).cuda()
offs = (np.arange(a.batch + 1) * a.frames).astype(np.int32)
out = torch.empty((a.batch, ex.embed_dim), device="cuda")
for _ in range(a.passes):
    ex.extract_packed(feats, offs, out)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
