"""Experiment: does tcgen05.mma read a K-major swizzled A tile correctly when the descriptor start address is shifted by
whole rows (not aligned to the swizzle repeat), with / without the descriptor base_offset field?  TDNN only (W = 1)."""
import os, sys, subprocess
code = r'''
import sys, os, numpy as np, torch
sys.path.insert(0, ".")
from voxsrc2020_speaker_verification_b200 import lib
lib.LIB_PATH = os.environ["SVX_LIB"]
from oracle import net_oracle
from voxsrc2020_speaker_verification_b200 import arch
from voxsrc2020_speaker_verification_b200.extractor import Extractor
fd = int(os.environ["FD"])
cfg = arch.get_config("tdnn")
params = net_oracle.init_params(cfg, fd, seed=4321)
ex = Extractor("tdnn", fd).load_params(params)
rng = np.random.default_rng(1)
utts = [net_oracle.synth_feats(rng, 1, t, fd)[0] for t in (300, 61)]
feats = torch.from_numpy(np.concatenate(utts, 0)).cuda()
offs = np.zeros(3, np.int32); offs[1:] = np.cumsum([u.shape[0] for u in utts])
got = ex.run_segments(feats, offs).cpu().numpy()
want = np.stack([net_oracle.forward(cfg, params, u[None])[0] for u in utts])
print("cos", (got * want).sum(1) / np.linalg.norm(got, axis=1) / np.linalg.norm(want, axis=1))
'''
for libname in sorted(os.listdir("voxsrc2020_speaker_verification_b200")):
    if not libname.startswith("libsvx_exp"):
        continue
    for fd in (40, 24):
        env = dict(os.environ, SVX_LIB=os.path.abspath("voxsrc2020_speaker_verification_b200/" + libname), FD=str(fd))
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        print(libname, "feat_dim", fd, (r.stdout.strip() or r.stderr.strip()[-300:]))
