"""Drop-in for the reference's ``tensorflow/tf_extract.py`` stage (same flags, same files in and out).

    python -m voxsrc2020_speaker_verification_b200.tf_extract --pb-file M.pb --expand-dim {2,3} --rspec P --wspec Q

reads ``P.scp`` (Kaldi script file of FBANK matrices, compressed or not), applies what the reference's pipe
``apply-cmvn-sliding --norm-vars=false --center=true --cmn-window=300`` applies (tf_extract.py:63), extracts one
embedding per utterance with the reference's chunk rule (tf_extract.py:96-111) and writes ``Q.ark`` + ``Q.scp``
like ``copy-vector ark:- ark,scp:Q.ark,Q.scp`` (tf_extract.py:65).  One process per GPU, selected by
CUDA_VISIBLE_DEVICES as in eval_inference_model.sh:29-36.
"""
from __future__ import annotations

import argparse
import io
import sys

import numpy as np
import torch

from . import kaldi_ark
from .extractor import Extractor


def _load_model(pb_file: str, expand_dim: int, device: int, precision: str, model_id, feat_dim) -> Extractor:
    """``--pb-file`` names a frozen graph (the reference's contract) or, as an addition, a checkpoint prefix ``…/model.ckpt-N``
    (then --model-id and --feat-dim are required: a checkpoint does not describe its graph)."""
    import os
    if not pb_file.endswith(".pb") and os.path.exists(pb_file + ".index"):
        if model_id is None or feat_dim is None:
            raise ValueError("a checkpoint prefix needs --model-id and --feat-dim")
        ex = Extractor.from_checkpoint(pb_file, model_id, feat_dim, device=device, precision=precision)
        if expand_dim is not None and expand_dim != ex.cfg.expand_dim:
            raise ValueError("--expand-dim %d does not fit model %s (needs %d)" % (expand_dim, model_id, ex.cfg.expand_dim))
        return ex
    return Extractor.from_pb(pb_file, expand_dim, device=device, precision=precision, model_id=model_id, feat_dim=feat_dim)


def run_distributed(pb_file: str, expand_dim: int, rspec: str, wspec: str, precision: str = "fp16", max_frames: int = 60000,
                    cmvn: bool = True, model_id=None, feat_dim=None) -> int:
    """One scp, all GPUs of the node (launched by torchrun, one rank per GPU): every rank reads the record HEADERS, takes a
    frame-balanced share of the utterances (dist.balance_by_frames), extracts it on its own GPU and the embeddings are
    all-gathered over NCCL; rank 0 writes ``wspec.ark/.scp`` in the scp's order.  Replaces the reference's static N-way scp split
    + N processes + ``cat`` (eval_inference_model.sh:29-39) when the utterance lengths are uneven."""
    import os
    import torch.distributed as tdist
    from . import dist as svdist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    own_group = not tdist.is_initialized()
    if own_group:
        tdist.init_process_group("nccl", device_id=torch.device("cuda", local))
    try:
        ex = _load_model(pb_file, expand_dim, local, precision, model_id, feat_dim)
        shapes = kaldi_ark.scp_shapes(rspec + ".scp")
        for key, rows, cols in shapes:
            if rows < 25:
                raise ZeroDivisionError("utterance %s has %d frames (< 25)" % (key, rows))
            if cols != ex.feat_dim:
                raise ValueError("utterance %s has %d-dim features, the model takes %d" % (key, cols, ex.feat_dim))

        def extract_local(indices):
            out = torch.empty((len(indices), ex.embed_dim), dtype=torch.float32, device=torch.device("cuda", local))
            recs = [(kind, payload, rows) for _, kind, payload, rows, _ in kaldi_ark.read_mat_scp_raw(rspec + ".scp", only=indices)]
            pos = 0
            while pos < len(recs):                      # launch sequences of <= 4 * max_frames frames, like run()
                end, frames = pos, 0
                while end < len(recs) and (end == pos or frames + recs[end][2] <= 4 * max_frames):
                    frames += recs[end][2]
                    end += 1
                chunk = recs[pos:end]
                if all(kind == "CM " for kind, _, _ in chunk):
                    emb = extract_compressed(ex, chunk, max_frames, cmvn)
                else:
                    mats = [kaldi_ark._decode_compressed(io.BytesIO(p), "CM ") if kind == "CM " else np.asarray(p, np.float32) for kind, p, _ in chunk]
                    emb = ex.extract_bucketed(mats, max_frames=max_frames, cmvn=cmvn)
                out[pos:end] = torch.from_numpy(emb).to(out.device)
                pos = end
            return out

        emb = svdist.extract_sharded(extract_local, [r for _, r, _ in shapes])
        if tdist.get_rank() == 0:
            host = emb.cpu().numpy()
            with kaldi_ark.VectorArkScpWriter(wspec) as writer:
                for (key, _, _), e in zip(shapes, host):
                    writer.write(key, e)
        tdist.barrier()
        return len(shapes)
    finally:
        if own_group:
            tdist.destroy_process_group()


def run(pb_file: str, expand_dim: int, rspec: str, wspec: str, precision: str = "fp16", device: int = 0,
        max_frames: int = 60000, cmvn: bool = True, model_id=None, feat_dim=None) -> int:
    ex = _load_model(pb_file, expand_dim, device, precision, model_id, feat_dim)
    n_done = 0
    with kaldi_ark.VectorArkScpWriter(wspec) as writer:
        keys, recs, frames = [], [], 0

        def flush():
            nonlocal keys, recs, frames, n_done
            if not keys:
                return
            if all(kind == "CM " for kind, _, _ in recs):
                emb = extract_compressed(ex, recs, max_frames, cmvn)      # decode + CMN + network, all on the device
            else:
                mats = [kaldi_ark._decode_compressed(io.BytesIO(p), "CM ") if kind == "CM " else np.asarray(p, np.float32)
                        for kind, p, _ in recs]
                emb = ex.extract_bucketed(mats, max_frames=max_frames, cmvn=cmvn)   # sliding CMN runs on the device
            for k, e in zip(keys, emb):          # written in input order, like the reference
                writer.write(k, e)
            n_done += len(keys)
            keys, recs, frames = [], [], 0

        for key, kind, payload, rows, cols in kaldi_ark.read_mat_scp_raw(rspec + ".scp"):
            if rows < 25:
                # the reference divides by zero here (tf_extract.py:102,111); fail as loudly
                raise ZeroDivisionError("utterance %s has %d frames (< 25)" % (key, rows))
            if cols != ex.feat_dim:
                raise ValueError("utterance %s has %d-dim features, the model takes %d" % (key, cols, ex.feat_dim))
            keys.append(key)
            recs.append((kind, payload, rows))
            frames += rows
            if frames >= 4 * max_frames:
                flush()
        flush()
    return n_done


def extract_compressed(ex: Extractor, recs, max_frames: int, cmvn: bool) -> np.ndarray:
    """Compressed FBANK records → embeddings without a host-side decode: length-sorted groups of ≤ max_frames frames are
    decoded (svx_decode_compressed), mean-normalised (svx_cmvn_sliding) and embedded on the device."""
    order = np.argsort([r for _, _, r in recs], kind="stable")
    out = np.zeros((len(recs), ex.embed_dim), np.float32)
    group, frames = [], 0
    for idx in list(order) + [None]:
        if idx is not None and (not group or frames + recs[idx][2] <= max_frames):
            group.append(int(idx))
            frames += recs[idx][2]
            continue
        if group:
            rows = [recs[i][2] for i in group]
            dev = ex.decode_compressed([recs[i][1] for i in group], rows)
            offs = np.zeros(len(group) + 1, np.int32)
            np.cumsum(rows, out=offs[1:])
            if cmvn:
                ex.cmvn_sliding(dev, offs)
            res = torch.empty((len(group), ex.embed_dim), dtype=torch.float32, device=dev.device)
            ex.extract_packed(dev, offs, res)
            out[group] = res.cpu().numpy()
        group, frames = ([int(idx)], recs[idx][2]) if idx is not None else ([], 0)
    return out


def main(argv=None) -> int:
    p = argparse.ArgumentParser()
    p.add_argument("--pb-file", dest="pb_file", default="", help="protobuffer model file")
    p.add_argument("--expand-dim", dest="expand_dim", default=2, type=int, help="expansion dimension of the input feature")
    p.add_argument("--rspec", dest="rspec", default="/tmp/fbank", help='source fbank scp path specification, without ".scp" suffix')
    p.add_argument("--wspec", dest="wspec", default="/tmp/xvector", help='destination xvector scp path specification, without ".scp" suffix')
    # additions (all optional; defaults reproduce the reference)
    p.add_argument("--precision", default="fp16", choices=["fp16", "bf16"])
    p.add_argument("--no-cmvn", dest="cmvn", action="store_false", help="features are already mean-normalised")
    p.add_argument("--max-frames", type=int, default=60000, help="frames per GPU launch sequence")
    p.add_argument("--model-id", default=None)
    p.add_argument("--feat-dim", type=int, default=None)
    p.add_argument("--distributed", action="store_true",
                   help="under torchrun: all ranks share ONE scp (frame-balanced), NCCL all-gather of the embeddings, rank 0 writes")
    a = p.parse_args(argv)
    if a.distributed:
        n = run_distributed(a.pb_file, a.expand_dim, a.rspec, a.wspec, a.precision, a.max_frames, a.cmvn, a.model_id, a.feat_dim)
        return 0
    n = run(a.pb_file, a.expand_dim, a.rspec, a.wspec, a.precision, 0, a.max_frames, a.cmvn, a.model_id, a.feat_dim)
    print("extracted %d embeddings → %s.ark" % (n, a.wspec), file=sys.stderr)
    return 0


if __name__ == "__main__":
    sys.exit(main())
