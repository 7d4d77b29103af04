"""Throughput of the non-headline BASELINE configs (TDNN, DPN, deep / attentive Res2Nets) for the record."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import net_oracle
from voxsrc2020_speaker_verification_b200 import arch
from voxsrc2020_speaker_verification_b200.extractor import Extractor


def bench(model, fd, batch, frames, reps=5):
    cfg = arch.get_config(model)
    params = net_oracle.init_params(cfg, fd, seed=4321, calib_frames=48, calib_batch=4)
    ex = Extractor(model, fd).load_params(params)
    lens = [frames] * batch if isinstance(frames, int) else list(np.random.default_rng(0).integers(frames[0], frames[1] + 1, batch))
    lens = sorted(int(x) for x in lens)
    feats = torch.randn(sum(lens), fd, device="cuda")
    offs = np.zeros(batch + 1, np.int32); offs[1:] = np.cumsum(lens)
    out = torch.empty((batch, ex.embed_dim), device="cuda")
    for _ in range(2):
        ex.extract_packed(feats, offs, out)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        ex.extract_packed(feats, offs, out)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
    print("%-28s feat %d batch %d frames %s: %.2f ms/step, %.0f emb/s, %.0f frames/s" % (model, fd, batch, frames, dt * 1e3, batch / dt, sum(lens) / dt), flush=True)


which = sys.argv[1:] or ["tdnn", "dpn68", "res2net200_w8_s6_c16", "res2net200_w24_s4_c32_att"]
if "tdnn" in which:
    bench("tdnn", 40, 64, 320); bench("tdnn", 40, 1024, 320)
if "dpn68" in which:
    bench("dpn68", 80, 128, (200, 600))
if "res2net200_w8_s6_c16" in which:
    bench("res2net200_w8_s6_c16", 80, 32, (300, 3000))
if "res2net200_w24_s4_c32_att" in which:
    bench("res2net200_w24_s4_c32_att", 80, 64, 200)
