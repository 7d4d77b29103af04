"""Debug: run one small batch twice (CUDA-core convs vs the tensor-core path), dump every op's output tensor and
report the first op whose outputs differ.  usage: dbg_dump.py [model] [feat_dim] [lens...]"""
import glob, os, subprocess, sys
import numpy as np
model = sys.argv[1] if len(sys.argv) > 1 else "res2net50_w24_s4_c32"
fd = int(sys.argv[2]) if len(sys.argv) > 2 else 80
lens = [int(x) for x in sys.argv[3:]] or [200, 57]
code = r'''
import numpy as np, torch, sys, os
sys.path.insert(0, ".")
from oracle import net_oracle
from voxsrc2020_speaker_verification_b200 import arch
from voxsrc2020_speaker_verification_b200.extractor import Extractor
model, fd, lens = %r, %d, %r
cfg = arch.get_config(model)
params = net_oracle.init_params(cfg, fd, seed=4321)
ex = Extractor(model, fd).load_params(params)
rng = np.random.default_rng(1234)
utts = [net_oracle.synth_feats(rng, 1, t, fd)[0] for t in lens]
feats = torch.from_numpy(np.concatenate(utts, 0)).cuda()
offs = np.zeros(len(utts) + 1, np.int32); offs[1:] = np.cumsum([u.shape[0] for u in utts])
out = ex.run_segments(feats, offs); torch.cuda.synchronize()
print("done", float(out.abs().mean()))
''' % (model, fd, lens)
for tag, extra in (("a", {"SVX_FORCE_SIMPLE": "1"}), ("b", {})):
    d = "/tmp/dump_" + tag
    os.makedirs(d, exist_ok=True)
    for f in glob.glob(d + "/*"):
        os.remove(f)
    env = dict(os.environ, SVX_DUMP_DIR=d, **extra)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(tag, r.stdout.strip(), r.stderr.strip()[-600:])
fa = sorted(glob.glob("/tmp/dump_a/*.bin"))
shown = 0
for pa in fa:
    pb = pa.replace("dump_a", "dump_b")
    name = os.path.basename(pa)[:-4]
    parts = dict((p[0], int(p[1:])) for p in name.split("_")[1:] if p[0] in "ktrwcn" and p[1:].lstrip("-").isdigit())
    off = int([p for p in name.split("_") if p.startswith("off")][0][3:])
    a = np.fromfile(pa, np.float16).reshape(parts["r"], parts["w"], parts["c"]).astype(np.float32)
    if not os.path.exists(pb):
        print(name, "missing in b"); continue
    b = np.fromfile(pb, np.float16).reshape(parts["r"], parts["w"], parts["c"]).astype(np.float32)
    sl = slice(off, off + parts["n"])
    d = np.abs(a[:, :, sl] - b[:, :, sl])
    tol = 2e-2 + 1e-2 * np.abs(a[:, :, sl])
    bad = np.argwhere(d > tol)
    print(name, "maxdiff %.4f" % d.max(), "maxabs %.2f" % np.abs(a[:, :, sl]).max(), "nbad", len(bad))
    if len(bad) and shown < 2:
        shown += 1
        rows = np.unique(bad[:, 0]); cols = np.unique(bad[:, 1]); ch = np.unique(bad[:, 2])
        print("  bad rows", rows[:40], "... n", len(rows)); print("  bad cols", cols); print("  bad ch", ch)
        r, c, k = bad[0]
        print("  first bad", bad[0], "a", a[r, c, off + k], "b", b[r, c, off + k])
