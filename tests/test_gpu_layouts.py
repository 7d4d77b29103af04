"""The alternative code paths of the flat conv kernel give the same embeddings as the default one.

The switches exist only in the debug build of the library (libsvx_dbg.so, `python -m voxsrc2020_speaker_verification_b200.build
--debug`; the production libsvx.so reads no environment variable), which the subprocess selects through SVX_LIB; the test is
skipped when that build is not in the tree.  The switches are read once per process, so every variant runs in its own subprocess:
the default path is covered by test_gpu_extract.py; here: unpadded concat / planar tensors (SVX_NO_YPAD), TMA-only epilogue (SVX_NO_DIRECT), two TMEM buffers (SVX_NO_TMEM4).
All of them are checked against the CPU oracle with the tolerance of BASELINE.json (cosine >= 0.9999 per utterance).
"""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SNIPPET = r"""
import json, sys
import numpy as np, torch
sys.path.insert(0, %r)
from oracle import net_oracle
from voxsrc2020_speaker_verification_b200 import arch
from voxsrc2020_speaker_verification_b200.extractor import Extractor
model_id, fd, lens = "res2net50_w24_s4_c32", 80, [96, 57]
cfg = arch.get_config(model_id)
params = net_oracle.init_params(cfg, fd, seed=4321)
ex = Extractor(model_id, fd).load_params(params)
rng = np.random.default_rng(7)
utts = [net_oracle.synth_feats(rng, 1, n, fd)[0] for n in lens]
feats = torch.from_numpy(np.concatenate(utts, 0)).cuda()
offs = np.zeros(len(utts) + 1, np.int32); offs[1:] = np.cumsum(lens)
got = ex.run_segments(feats, offs).cpu().numpy()
ref = np.stack([net_oracle.forward(cfg, params, u[None])[0] for u in utts])
cos = (got * ref).sum(1) / np.linalg.norm(got, axis=1) / np.linalg.norm(ref, axis=1)
print("RESULT " + json.dumps({"cos": cos.tolist()}))
""" % ROOT


@pytest.mark.parametrize("env", [{"SVX_NO_YPAD": "1"}, {"SVX_NO_DIRECT": "1"}, {"SVX_NO_TMEM4": "1"}],
                         ids=["unpadded_splits", "tma_epilogue", "two_tmem_buffers"])
def test_variant_matches_oracle(env):
    dbg = os.path.join(ROOT, "voxsrc2020_speaker_verification_b200", "libsvx_dbg.so")
    if not os.path.exists(dbg):
        pytest.skip("debug library not built")
    e = dict(os.environ)
    e.update(env)
    e["SVX_LIB"] = dbg
    r = subprocess.run([sys.executable, "-c", SNIPPET], capture_output=True, text=True, env=e, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1]
    cos = json.loads(line[len("RESULT "):])["cos"]
    assert min(cos) >= 0.9999, cos
