"""Debug: run the headline model at growing batch sizes with a sync after every op; print the plan of the failing op."""
import os, sys
os.environ["SVX_SYNC_EACH"] = "1"
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import net_oracle
from voxsrc2020_speaker_verification_b200 import arch
from voxsrc2020_speaker_verification_b200.extractor import Extractor
model, fd = (sys.argv[1] if len(sys.argv) > 1 else "res2net50_w24_s4_c32"), 80
cfg = arch.get_config(model)
params = net_oracle.init_params(cfg, fd, seed=4321, calib_frames=48, calib_batch=4)
ex = Extractor(model, fd).load_params(params)
for batch in ([int(x) for x in sys.argv[2:]] or [2, 8, 32, 128, 256]):
    feats = torch.from_numpy(net_oracle.synth_feats(np.random.default_rng(0), batch, 200, fd).reshape(-1, fd)).cuda()
    offs = (np.arange(batch + 1) * 200).astype(np.int32)
    try:
        out = ex.run_segments(feats, offs)
        torch.cuda.synchronize()
        print("batch", batch, "ok", float(out.abs().mean()), flush=True)
    except Exception as e:
        print("batch", batch, "FAILED:", e, flush=True)
        break
