"""Drop-in for the reference's ``tensorflow/eval_inference_model.sh`` (the stage that strings the hot path together).

    python -m voxsrc2020_speaker_verification_b200.eval_inference_model MODEL.pb [EXPANSION_DIM] \
        [--data-root ../data] [--num-gpus 8] [--eer-script /path/to/eer_minDCF.py]

Same inputs and outputs as the reference script (eval_inference_model.sh:19-60):

  * for ``voxceleb2_dev`` and ``voxceleb1``: one extraction process per GPU on ``<data-root>/<set>/<N>-split/feats.<i>.scp``
    (``CUDA_VISIBLE_DEVICES=i-1``, no communication, like the reference), written to
    ``<model>_embeddings/<set>/xvector.<i>.{ark,scp}``, then the shards concatenated to ``xvector.ark``
  * for the T / E / H trial lists: ``cosine_<x>.txt`` and ``snorm_<x>.txt`` next to the voxceleb1 embeddings
  * optionally the reference's own ``eer_minDCF.py`` on every score file (it runs on them unchanged).

Differences that do not change any output: the three scoring jobs run one after another on one GPU (each takes
milliseconds) instead of three CPU processes, and the Kaldi pipes are not needed (``tf_extract`` reads the scp and applies
the sliding CMN itself).
"""
from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys

DATASETS = ("voxceleb2_dev", "voxceleb1")       # eval_inference_model.sh:27
TESTSETS = ("T", "E", "H")                      # eval_inference_model.sh:42


def extract_dataset(model: str, expansion_dim: int, data_root: str, dataset: str, out_dir: str, num_gpus: int, extra=()) -> None:
    os.makedirs(out_dir, exist_ok=True)
    procs = []
    for i in range(1, num_gpus + 1):             # eval_inference_model.sh:29-36
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=str(i - 1))
        cmd = [sys.executable, "-m", "voxsrc2020_speaker_verification_b200.tf_extract", "--pb-file", model, "--expand-dim", str(expansion_dim),
               "--rspec", os.path.join(data_root, dataset, "%d-split" % num_gpus, "feats.%d" % i),
               "--wspec", os.path.join(out_dir, "xvector.%d" % i)] + list(extra)
        procs.append(subprocess.Popen(cmd, env=env))
    codes = [p.wait() for p in procs]            # :37
    if any(codes):
        raise RuntimeError("extraction failed for %s: exit codes %s" % (dataset, codes))
    with open(os.path.join(out_dir, "xvector.ark"), "wb") as out:    # :38-39 — records are self-delimiting, so cat is enough
        for i in range(1, num_gpus + 1):
            with open(os.path.join(out_dir, "xvector.%d.ark" % i), "rb") as f:
                shutil.copyfileobj(f, out)


def main(argv=None) -> int:
    p = argparse.ArgumentParser()
    p.add_argument("model", help="frozen .pb (export_inference_model.sh)")
    p.add_argument("expansion_dim", nargs="?", type=int, default=2)
    p.add_argument("--data-root", default="../data")
    p.add_argument("--num-gpus", type=int, default=8, help="global_config.sh:19")
    p.add_argument("--topk", type=int, default=400)
    p.add_argument("--eer-script", default=None, help="path of the reference's eer_minDCF.py to run on the score files")
    p.add_argument("--skip-extraction", action="store_true")
    a, extra = p.parse_known_args(argv)
    out_root = os.path.abspath(a.model[:-3] if a.model.endswith(".pb") else a.model) + "_embeddings"     # :25
    if not a.skip_extraction:
        for ds in DATASETS:
            extract_dataset(a.model, a.expansion_dim, a.data_root, ds, os.path.join(out_root, ds), a.num_gpus, extra)
    from .scoring import score_files
    vox1 = os.path.join(out_root, "voxceleb1")
    for t in TESTSETS:                           # :42-51
        trial = os.path.join(a.data_root, "voxceleb1_trials", "list_test_%s.txt" % t)
        if not os.path.exists(trial):
            continue
        score_files(trial, os.path.join(vox1, "xvector.ark"), os.path.join(vox1, "cosine_%s.txt" % t),
                    cohort_ark=os.path.join(out_root, "voxceleb2_dev", "xvector.ark"),
                    cohort_spk2utt=os.path.join(a.data_root, "voxceleb2_dev", "spk2utt"),
                    snorm_score=os.path.join(vox1, "snorm_%s.txt" % t), topk=a.topk)
    if a.eer_script:
        for t in TESTSETS:                       # :53-60
            trial = os.path.join(a.data_root, "voxceleb1_trials", "list_test_%s.txt" % t)
            for kind in ("cosine", "snorm"):
                score = os.path.join(vox1, "%s_%s.txt" % (kind, t))
                if os.path.exists(trial) and os.path.exists(score):
                    subprocess.run([sys.executable, a.eer_script, "--trial", trial, "--score", score], check=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
