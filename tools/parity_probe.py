"""Measured whole-network parity (cosine, relative L2, norm ratio) of configurations that had no GPU test in round 1."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import net_oracle
from voxsrc2020_speaker_verification_b200 import arch
from voxsrc2020_speaker_verification_b200.extractor import Extractor

def probe(model_id, fd, lens, precision="fp16", damp=4.0):
    cfg = arch.get_config(model_id)
    t0 = time.time()
    params = net_oracle.init_params(cfg, fd, seed=4321, damp=damp)
    ex = Extractor(model_id, fd, precision=precision).load_params(params)
    rng = np.random.default_rng(77)
    utts = [net_oracle.synth_feats(rng, 1, t, fd)[0] for t in lens]
    got = ex.extract(utts)
    t1 = time.time()
    want = np.stack([net_oracle.extract_utterance(cfg, params, u) for u in utts])
    cos = (got * want).sum(1) / np.linalg.norm(got, axis=1) / np.linalg.norm(want, axis=1)
    rel = np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)
    ratio = np.linalg.norm(got, axis=1) / np.linalg.norm(want, axis=1)
    print("%-28s %-5s damp %.0f lens %s: cos %s rel %s norm-ratio %s  (oracle %.1f s)" % (
        model_id, precision, damp, lens, np.round(cos, 6), np.round(rel, 4), np.round(ratio, 4), time.time() - t1), flush=True)

probe("res2net200_w8_s6_c16", 80, [300, 1025, 3000])
probe("res2net101_w24_s4_c32_att", 80, [200, 57])
probe("dpn68", 80, [599, 600])
probe("res2net50_w24_s4_c32", 80, [200, 57], "bf16")
probe("dpn68", 80, [200, 57], "bf16")
probe("res2net50_w24_s4_c32", 80, [200, 57], "fp16")
probe("res2net50_w24_s4_c32", 80, [200, 57], "fp16", damp=1.0)
probe("res2net50_w24_s4_c32", 80, [200, 57], "bf16", damp=1.0)
probe("dpn68", 80, [200, 57], "fp16", damp=1.0)
probe("tdnn", 40, [320, 100], "fp16", damp=1.0)
