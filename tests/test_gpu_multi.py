"""Two ranks on two GPUs over NCCL (skipped on a one-GPU box): the distributed extraction entry point
(`torchrun … -m voxsrc2020_speaker_verification_b200.tf_extract --distributed`: frame-balanced shares of ONE scp, all-gather of the
embeddings, rank 0 writes) must produce the ark the single-process stage produces, and both multi-GPU layouts of the cohort
statistics must equal the single-GPU result.  Reference layout: eval_inference_model.sh:29-39."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import net_oracle
from voxsrc2020_speaker_verification_b200 import arch, kaldi_ark, pb_loader

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _two_gpus():
    return torch.cuda.is_available() and torch.cuda.device_count() >= 2


def _torchrun(args, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port)] + args
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)


@pytest.mark.skipif(not _two_gpus(), reason="needs two GPUs")
def test_distributed_extraction_equals_single_process(tmp_path):
    from voxsrc2020_speaker_verification_b200 import tf_extract
    cfg = arch.get_config("res2net50_w8_s6_c16")
    params = net_oracle.init_params(cfg, 40, seed=4321, calib_frames=48, calib_batch=4)
    pb = str(tmp_path / "m.pb")
    pb_loader.write_pb(pb, params, cfg, 40)
    rng = np.random.default_rng(5)
    ark, scp = str(tmp_path / "f.ark"), str(tmp_path / "feats.1.scp")
    lens = [300, 25, 1999, 1000, 57, 640, 41, 2500, 33]
    with open(ark, "wb") as fa, open(scp, "w") as fs:
        for i, t in enumerate(lens):
            key = "utt%02d" % i
            off = kaldi_ark.write_mat(fa, (rng.standard_normal((t, 40)) * 2 + 1).astype(np.float32), key)
            fs.write("%s %s:%d\n" % (key, ark, off))
    single = str(tmp_path / "single")
    assert tf_extract.main(["--pb-file", pb, "--expand-dim", "3", "--rspec", scp[:-4], "--wspec", single]) == 0
    multi = str(tmp_path / "multi")
    r = _torchrun(["-m", "voxsrc2020_speaker_verification_b200.tf_extract", "--distributed", "--pb-file", pb, "--expand-dim", "3",
                   "--rspec", scp[:-4], "--wspec", multi], 29611)
    assert r.returncode == 0, r.stderr[-3000:]
    a = list(kaldi_ark.read_vec_flt_ark(single + ".ark"))
    b = list(kaldi_ark.read_vec_flt_ark(multi + ".ark"))
    assert [k for k, _ in a] == [k for k, _ in b] == ["utt%02d" % i for i in range(len(lens))]
    np.testing.assert_allclose(np.stack([v for _, v in b]), np.stack([v for _, v in a]), rtol=0, atol=2e-5)
    assert [k for k, _, _ in kaldi_ark.read_scp(multi + ".scp")] == [k for k, _ in a]


@pytest.mark.skipif(not _two_gpus(), reason="needs two GPUs")
def test_sharded_cohort_statistics_equal_single_gpu():
    r = _torchrun([os.path.join(ROOT, "tools", "check_dist_score.py")], 29612)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    assert "OK" in r.stdout
