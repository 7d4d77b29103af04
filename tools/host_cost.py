"""Host-side cost of one extraction call (time until the call returns, before the device has finished)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import net_oracle
from voxsrc2020_speaker_verification_b200 import arch
from voxsrc2020_speaker_verification_b200.extractor import Extractor
for model, batch, frames in (("res2net50_w24_s4_c32", 256, 200), ("res2net200_w8_s6_c16", 32, 1500)):
    cfg = arch.get_config(model)
    params = net_oracle.init_params(cfg, 80, seed=4321, calib_frames=48, calib_batch=4)
    ex = Extractor(model, 80).load_params(params)
    feats = torch.randn(batch * frames, 80, device="cuda")
    offs = (np.arange(batch + 1) * frames).astype(np.int32)
    out = torch.empty((batch, ex.embed_dim), device="cuda")
    for _ in range(2):
        ex.extract_packed(feats, offs, out)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); ex.extract_packed(feats, offs, out); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("%s: launches %d, host returns after %.2f ms, device done after %.2f ms -> %.1f us per launch on the host" %
          (model, ex.last_launches, (t1 - t0) * 1e3, (t2 - t0) * 1e3, (t1 - t0) * 1e6 / ex.last_launches), flush=True)
