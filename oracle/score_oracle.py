"""ORACLE (test infrastructure, not product code): NumPy restatement of the reference scoring stage.

Each function follows reference tensorflow/snorm.py line by line in fp32 NumPy.  It is PINNED: the
fixtures under tests/golden/ were produced by importing the reference's own snorm.py / kaldi_io.py in
the build container (oracle/gen_golden.py) and tests/test_oracle_score.py checks this restatement
against them bit-for-bit.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np


def l2norm(x: np.ndarray, axis: int = 0, keepdims: bool = True) -> np.ndarray:
    """reference snorm.py:23-25."""
    return x / np.linalg.norm(x, axis=axis, keepdims=keepdims)


def normalise_xvectors(vectors: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """reference snorm.py:28-33 (read_xvector without the ark reader): L2-normalise on load."""
    return {utt: l2norm(np.asarray(vec), axis=0) for utt, vec in vectors.items()}


def read_spk2utt(path: str) -> Dict[str, List[str]]:
    """reference snorm.py:36-42."""
    spk2utt = {}
    for line in open(path, "r"):
        f = line.strip().split()
        spk2utt[f[0]] = f[1:]
    return spk2utt


def read_speaker_xvector(xvectors: Dict[str, np.ndarray], spk2utt: Dict[str, Sequence[str]]) -> Dict[str, np.ndarray]:
    """reference snorm.py:45-67 — per speaker: re-normalise rows, mean over utterances, NOT re-normalised.
    Speakers appear in the order their first utterance appears in ``xvectors``."""
    utt_to_spk = {}
    for spk, utts in spk2utt.items():
        for utt in utts:
            utt_to_spk[utt] = spk
    speaker = {}
    for utt, vec in xvectors.items():
        if utt in utt_to_spk:
            speaker.setdefault(utt_to_spk[utt], []).append(vec)
    for spk, vecs in speaker.items():
        m = l2norm(np.array(vecs), axis=1)
        speaker[spk] = np.mean(m, axis=0)
    return speaker


def get_cohort_mean_std(trial_xvectors: Dict[str, np.ndarray], cohort_xvectors: Dict[str, np.ndarray],
                        topk: int = 400) -> Tuple[Dict[str, np.float32], Dict[str, np.float32]]:
    """reference snorm.py:83-110 — 1024-row blocks; full descending sort; mean and population std of
    the first ``topk`` columns (the whole row when topk exceeds the cohort size)."""
    utt = list(trial_xvectors.keys())
    trial_matrix = np.array(list(trial_xvectors.values()))
    cohort_matrix_t = np.transpose(np.array(list(cohort_xvectors.values())))
    utt_to_mean, utt_to_std = {}, {}
    sub = 1024
    for i in range(0, len(trial_matrix), sub):
        e = min(i + sub, len(trial_matrix))
        s = np.matmul(trial_matrix[i:e, :], cohort_matrix_t)
        top = (-1 * np.sort(-s, axis=1))[:, :topk]
        mean = np.mean(top, axis=1)
        std = np.std(top, axis=1)
        for j in range(i, e):
            utt_to_mean[utt[j]], utt_to_std[utt[j]] = mean[j % 1024], std[j % 1024]
    return utt_to_mean, utt_to_std


def parse_trials(path: str) -> List[Tuple[str, str]]:
    """reference snorm.py:115-116 — the last two whitespace fields of every line."""
    out = []
    for line in open(path, "r"):
        u1, u2 = line.strip().split()[-2:]
        out.append((u1, u2))
    return out


def get_cosine_score(trial_xvectors: Dict[str, np.ndarray], trial_path: str):
    """reference snorm.py:113-120."""
    return [(u1, u2, np.dot(trial_xvectors[u1], trial_xvectors[u2])) for u1, u2 in parse_trials(trial_path)]


def get_asnorm1_score(utt_to_mean, utt_to_std, scores):
    """reference snorm.py:123-131."""
    return [(u1, u2, 0.5 * ((s - utt_to_mean[u1]) / utt_to_std[u1] + (s - utt_to_mean[u2]) / utt_to_std[u2]))
            for (u1, u2, s) in scores]


def write_scores(path: str, scores) -> None:
    """reference snorm.py:164-166,180-182 — ``print(utt1, utt2, score)``."""
    with open(path, "w") as f:
        for (u1, u2, s) in scores:
            print(u1, u2, s, file=f)


# ----------------------------------------------------------------- array-level forms used by the benches
def cohort_mean_std_arrays(x: np.ndarray, cohort: np.ndarray, topk: int):
    """Same arithmetic as get_cohort_mean_std on plain arrays ([n,D] unit rows, [c,D] cohort)."""
    n = x.shape[0]
    mean = np.empty(n, np.float32)
    std = np.empty(n, np.float32)
    ct = np.transpose(cohort)
    for i in range(0, n, 1024):
        e = min(i + 1024, n)
        s = np.matmul(x[i:e], ct)
        top = (-1 * np.sort(-s, axis=1))[:, :topk]
        mean[i:e] = np.mean(top, axis=1)
        std[i:e] = np.std(top, axis=1)
    return mean, std


def trial_scores_arrays(x: np.ndarray, idx1: np.ndarray, idx2: np.ndarray, mean: np.ndarray, std: np.ndarray):
    """Vectorised form of get_cosine_score + get_asnorm1_score for index-pair trial lists."""
    cos = np.einsum("ij,ij->i", x[idx1], x[idx2]).astype(np.float32)
    sn = 0.5 * ((cos - mean[idx1]) / std[idx1] + (cos - mean[idx2]) / std[idx2])
    return cos, sn.astype(np.float32)
