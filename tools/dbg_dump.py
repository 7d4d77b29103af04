"""Debug: run one small Res2Net batch twice (staged mask A vs B), dump every op's output tensor, report first difference."""
import glob, os, subprocess, sys
import numpy as np
code = r'''
import numpy as np, torch, sys, os
sys.path.insert(0, ".")
from oracle import net_oracle
from voxsrc2020_speaker_verification_b200 import arch
from voxsrc2020_speaker_verification_b200.extractor import Extractor
cfg = arch.get_config("res2net50_w24_s4_c32")
params = net_oracle.init_params(cfg, 80, seed=4321)
ex = Extractor("res2net50_w24_s4_c32", 80).load_params(params)
rng = np.random.default_rng(1234)
utts = [net_oracle.synth_feats(rng, 1, t, 80)[0] for t in (200, 57)]
feats = torch.from_numpy(np.concatenate(utts, 0)).cuda()
offs = np.zeros(len(utts) + 1, np.int32); offs[1:] = np.cumsum([u.shape[0] for u in utts])
out = ex.run_segments(feats, offs); torch.cuda.synchronize()
print("done", float(out.abs().mean()))
'''
for tag, mask in (("a", "0"), ("b", sys.argv[1] if len(sys.argv) > 1 else "32")):
    d = "/tmp/dump_" + tag
    os.makedirs(d, exist_ok=True)
    for f in glob.glob(d + "/*"):
        os.remove(f)
    env = dict(os.environ, SVX_DUMP_DIR=d, SVX_STAGED_MASK=mask)
    print(subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True).stdout.strip())
fa = sorted(glob.glob("/tmp/dump_a/*.bin"))
shown = 0
for pa in fa:
    pb = pa.replace("dump_a", "dump_b")
    name = os.path.basename(pa)[:-4]
    parts = dict((p[0], int(p[1:])) for p in name.split("_")[1:] if p[0] in "ktrwcn" and p[1:].lstrip("-").isdigit())
    off = int([p for p in name.split("_") if p.startswith("off")][0][3:])
    a = np.fromfile(pa, np.float16).reshape(parts["r"], parts["w"], parts["c"]).astype(np.float32)
    b = np.fromfile(pb, np.float16).reshape(parts["r"], parts["w"], parts["c"]).astype(np.float32)
    sl = slice(off, off + parts["n"])
    d = np.abs(a[:, :, sl] - b[:, :, sl])
    bad = np.argwhere(d > 1e-3)
    print(name, "maxdiff %.4f" % d.max(), "nbad", len(bad))
    if len(bad) and shown < 2:
        shown += 1
        rows = np.unique(bad[:, 0]); cols = np.unique(bad[:, 1]); ch = np.unique(bad[:, 2])
        print("  bad rows", rows[:40], "... n", len(rows)); print("  bad cols", cols); print("  bad ch", ch)
        r, c, k = bad[0]
        print("  first bad", bad[0], "a", a[r, c, off + k], "b", b[r, c, off + k])
