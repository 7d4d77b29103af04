"""Cosine + adaptive symmetric normalisation (AS-norm1) scoring on the GPU.

Same function names, arguments and return types as the reference's tensorflow/snorm.py so that code written
against it keeps working:

    read_xvector, read_spk2utt, read_speaker_xvector, get_cohort_xvector, get_projection_weight,
    get_cohort_mean_std(trial_xvectors, cohort_xvectors, topk=400), get_cosine_score(xvectors, trial),
    get_asnorm1_score(utt_to_mean, utt_to_std, scores)

plus ``Scorer``, the array-level engine they are built on (device tensors in, device tensors out).
All arithmetic runs in libsvx CUDA kernels; there is no NumPy fallback.
"""
from __future__ import annotations

import ctypes
import pickle
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import kaldi_ark, lib


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p()


class Scorer:
    """One per GPU.  Tensors are fp32 / int32 CUDA tensors on ``device``."""

    def __init__(self, device: int = 0):
        self.device = int(device)
        self._lib = lib.load()
        self._h = ctypes.c_void_p()
        lib.check(self._lib.svx_scorer_create(self.device, ctypes.byref(self._h)))
        self.launches = 0

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set_option(self, key: str, value: int) -> None:
        """"fused" = 0: cohort statistics through the unfused kernels (the fused kernel is the default)."""
        lib.check(self._lib.svx_scorer_set_option(self._h, key.encode(), int(value)))

    def last_path(self) -> Tuple[int, int]:
        """(rows finished by the fused kernel, rows handed to the unfused kernels) of the last statistics call."""
        a, b = ctypes.c_longlong(), ctypes.c_longlong()
        lib.check(self._lib.svx_scorer_last_path(self._h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def l2norm(self, x: torch.Tensor) -> torch.Tensor:
        """snorm.l2norm per row (snorm.py:23-25,32)."""
        x = x.contiguous()
        out = torch.empty_like(x)
        lib.check(self._lib.svx_l2norm_rows(_ptr(x), _ptr(out), x.shape[0], x.shape[1], self._stream()))
        self.launches += 1
        return out

    def group_means(self, unit_rows: torch.Tensor, group, n_groups: int) -> torch.Tensor:
        """read_speaker_xvector (snorm.py:45-67): mean of the unit rows of each group, not re-normalised.
        ``group[i]`` = group of row i (< 0: unused).  Members are summed in row order (the order the reference appends
        them) and divided by the count, like np.mean(matrix, axis=0): deterministic and bit-identical to it."""
        g = np.asarray(group.cpu() if isinstance(group, torch.Tensor) else group, dtype=np.int64)
        used = np.nonzero(g >= 0)[0]
        order = used[np.argsort(g[used], kind="stable")].astype(np.int32)          # CSR: rows of group 0, then group 1, …
        offs = np.zeros(n_groups + 1, np.int32)
        np.cumsum(np.bincount(g[used], minlength=n_groups), out=offs[1:])
        dv = unit_rows.device
        out = torch.empty((n_groups, unit_rows.shape[1]), dtype=torch.float32, device=dv)
        rows_d, order_d, offs_d = unit_rows.contiguous(), torch.from_numpy(order).to(dv), torch.from_numpy(offs).to(dv)   # held until the launch
        lib.check(self._lib.svx_group_means(_ptr(rows_d), rows_d.shape[1], _ptr(order_d), _ptr(offs_d), _ptr(out), n_groups, self._stream()))
        self.launches += 1
        return out

    def cohort_mean_std(self, test: torch.Tensor, cohort: torch.Tensor, topk: int = 400) -> Tuple[torch.Tensor, torch.Tensor]:
        """get_cohort_mean_std (snorm.py:83-110) on arrays: test [n,D] unit rows, cohort [c,D]."""
        test, cohort = test.contiguous(), cohort.contiguous()
        n, d = test.shape
        mean = torch.empty(n, dtype=torch.float32, device=test.device)
        std = torch.empty(n, dtype=torch.float32, device=test.device)
        lib.check(self._lib.svx_asnorm_stats(self._h, _ptr(test), n, _ptr(cohort), cohort.shape[0], d, int(topk),
                                             _ptr(mean), _ptr(std), self._stream()))
        self.launches += int(self._lib.svx_scorer_last_launches(self._h))
        return mean, std

    def cohort_topk_values(self, test: torch.Tensor, cohort_shard: torch.Tensor, topk: int) -> torch.Tensor:
        """Per-shard candidates for the cohort-row-sharded layout: [n, topk] largest dot products (unordered)."""
        test, cohort_shard = test.contiguous(), cohort_shard.contiguous()
        n, d = test.shape
        vals = torch.empty((n, topk), dtype=torch.float32, device=test.device)
        lib.check(self._lib.svx_cohort_topk_values(self._h, _ptr(test), n, _ptr(cohort_shard), cohort_shard.shape[0], d, int(topk),
                                                   _ptr(vals), self._stream()))
        self.launches += int(self._lib.svx_scorer_last_launches(self._h))
        return vals

    def topk_stats(self, vals: torch.Tensor, topk: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Mean / population std of the topk largest entries of every row of ``vals`` [n, m]."""
        vals = vals.contiguous()
        n, m = vals.shape
        mean = torch.empty(n, dtype=torch.float32, device=vals.device)
        std = torch.empty(n, dtype=torch.float32, device=vals.device)
        lib.check(self._lib.svx_topk_stats(_ptr(vals), m, n, m, int(topk), _ptr(mean), _ptr(std), self._stream()))
        self.launches += 1
        return mean, std

    def trial_scores(self, emb: torch.Tensor, idx1: torch.Tensor, idx2: torch.Tensor, mean: Optional[torch.Tensor] = None,
                     std: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """get_cosine_score + get_asnorm1_score (snorm.py:113-131) for index-pair trials."""
        emb, idx1, idx2 = emb.contiguous(), idx1.contiguous(), idx2.contiguous()      # named, so that they live until the launch
        t = idx1.shape[0]
        cos = torch.empty(t, dtype=torch.float32, device=emb.device)
        sn = torch.empty(t, dtype=torch.float32, device=emb.device) if mean is not None else None
        lib.check(self._lib.svx_trial_scores(_ptr(emb), emb.shape[1], _ptr(idx1), _ptr(idx2), t,
                                             _ptr(mean), _ptr(std), _ptr(cos), _ptr(sn), self._stream()))
        self.launches += 1
        return cos, sn

    def close(self):
        if self._h:
            self._lib.svx_scorer_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default: Dict[int, Scorer] = {}


def default_scorer(device: int = 0) -> Scorer:
    if device not in _default:
        _default[device] = Scorer(device)
    return _default[device]


def _dev(device: int) -> torch.device:
    return torch.device("cuda", device)


# ---------------------------------------------------------------------------------------------------------
# snorm.py-compatible API (dicts of numpy vectors in, dicts / lists out)
def l2norm(x, axis=0, keepdims=True):
    """snorm.py:23-25 for one vector or a matrix of row vectors (axis = last)."""
    x = np.asarray(x, dtype=np.float32)
    rows = x.reshape(1, -1) if x.ndim == 1 else x
    if x.ndim == 2 and axis not in (1, -1):
        raise ValueError("matrix l2norm is row-wise (axis=1), as the reference uses it")
    out = default_scorer().l2norm(torch.from_numpy(np.ascontiguousarray(rows)).to(_dev(0)))
    return out.cpu().numpy().reshape(x.shape)


def read_xvector(xvector_ark: str, device: int = 0) -> Dict[str, np.ndarray]:
    """snorm.py:28-33 — ark → {utt: unit vector}; normalisation runs on the GPU."""
    keys, mat = kaldi_ark.read_vec_ark_matrix(xvector_ark)
    if not keys:
        return {}
    unit = default_scorer(device).l2norm(torch.from_numpy(mat).to(_dev(device))).cpu().numpy()
    return {k: unit[i] for i, k in enumerate(keys)}


def read_spk2utt(spk2utt_file: str) -> Dict[str, List[str]]:
    """snorm.py:36-42."""
    spk2utt = {}
    with open(spk2utt_file, "r") as f:
        for line in f:
            fields = line.strip().split()
            if fields:
                spk2utt[fields[0]] = fields[1:]
    return spk2utt


def read_speaker_xvector(xvectors: Dict[str, np.ndarray], spk2utt: Dict[str, Sequence[str]], device: int = 0) -> Dict[str, np.ndarray]:
    """snorm.py:45-67 — per speaker: rows re-normalised, mean over utterances, not re-normalised.
    Speakers are ordered by the first appearance of one of their utterances in ``xvectors``."""
    utt_to_spk = {}
    for spk, utts in spk2utt.items():
        for utt in utts:
            utt_to_spk[utt] = spk
    spk_index: Dict[str, int] = {}
    rows, group = [], []
    for utt, vec in xvectors.items():
        spk = utt_to_spk.get(utt)
        if spk is None:
            continue
        group.append(spk_index.setdefault(spk, len(spk_index)))
        rows.append(vec)
    if not rows:
        return {}
    sc = default_scorer(device)
    x = sc.l2norm(torch.from_numpy(np.ascontiguousarray(np.stack(rows), dtype=np.float32)).to(_dev(device)))
    means = sc.group_means(x, torch.tensor(group, dtype=torch.int32, device=_dev(device)), len(spk_index)).cpu().numpy()
    return {spk: means[i] for spk, i in spk_index.items()}


def get_cohort_xvector(cohort_ark: str, cohort_spk2utt: str, device: int = 0) -> Dict[str, np.ndarray]:
    """snorm.py:70-74."""
    return read_speaker_xvector(read_xvector(cohort_ark, device), read_spk2utt(cohort_spk2utt), device)


def get_projection_weight(weight_matrix_pkl: str, device: int = 0) -> Dict[int, np.ndarray]:
    """snorm.py:77-80 — cohort from the classifier's projection matrix, rows L2-normalised."""
    with open(weight_matrix_pkl, "rb") as f:
        w = np.asarray(pickle.load(f), dtype=np.float32)
    unit = default_scorer(device).l2norm(torch.from_numpy(np.ascontiguousarray(w)).to(_dev(device))).cpu().numpy()
    return {i: unit[i] for i in range(len(unit))}


def get_cohort_mean_std(trial_xvectors: Dict[str, np.ndarray], cohort_xvectors: Dict, topk: int = 400, device: int = 0):
    """snorm.py:83-110 → ({utt: mean}, {utt: std}) as np.float32 scalars."""
    utts = list(trial_xvectors.keys())
    x = torch.from_numpy(np.ascontiguousarray(np.array(list(trial_xvectors.values()), dtype=np.float32))).to(_dev(device))
    c = torch.from_numpy(np.ascontiguousarray(np.array(list(cohort_xvectors.values()), dtype=np.float32))).to(_dev(device))
    mean, std = default_scorer(device).cohort_mean_std(x, c, topk)
    mean, std = mean.cpu().numpy(), std.cpu().numpy()
    return dict(zip(utts, mean)), dict(zip(utts, std))


def parse_trials(trial: str) -> List[Tuple[str, str]]:
    """snorm.py:115-116 — the last two whitespace-separated fields of every line."""
    pairs = []
    with open(trial, "r") as f:
        for line in f:
            u1, u2 = line.strip().split()[-2:]
            pairs.append((u1, u2))
    return pairs


def _index_trials(xvectors: Dict[str, np.ndarray], pairs):
    index = {k: i for i, k in enumerate(xvectors.keys())}
    idx1 = np.fromiter((index[a] for a, _ in pairs), dtype=np.int32, count=len(pairs))   # KeyError on unknown utt, as the reference
    idx2 = np.fromiter((index[b] for _, b in pairs), dtype=np.int32, count=len(pairs))
    return idx1, idx2


def get_cosine_score(trial_xvectors: Dict[str, np.ndarray], trial: str, device: int = 0):
    """snorm.py:113-120 → [(utt1, utt2, np.float32 score)] in trial-file order."""
    pairs = parse_trials(trial)
    if not pairs:
        return []
    idx1, idx2 = _index_trials(trial_xvectors, pairs)
    x = torch.from_numpy(np.ascontiguousarray(np.array(list(trial_xvectors.values()), dtype=np.float32))).to(_dev(device))
    cos, _ = default_scorer(device).trial_scores(x, torch.from_numpy(idx1).to(_dev(device)), torch.from_numpy(idx2).to(_dev(device)))
    cos = cos.cpu().numpy()
    return [(a, b, s) for (a, b), s in zip(pairs, cos)]


def get_asnorm1_score(utt_to_mean, utt_to_std, scores, device: int = 0):
    """snorm.py:123-131 — 0.5*((s-m1)/s1 + (s-m2)/s2) per trial (elementwise on the GPU)."""
    if not scores:
        return []
    dv = _dev(device)
    s = torch.tensor([float(x[2]) for x in scores], dtype=torch.float32, device=dv)
    m1 = torch.tensor([float(utt_to_mean[x[0]]) for x in scores], dtype=torch.float32, device=dv)
    s1 = torch.tensor([float(utt_to_std[x[0]]) for x in scores], dtype=torch.float32, device=dv)
    m2 = torch.tensor([float(utt_to_mean[x[1]]) for x in scores], dtype=torch.float32, device=dv)
    s2 = torch.tensor([float(utt_to_std[x[1]]) for x in scores], dtype=torch.float32, device=dv)
    out = (0.5 * ((s - m1) / s1 + (s - m2) / s2)).cpu().numpy()
    return [(a, b, v) for (a, b, _), v in zip(scores, out)]


def write_scores(path: str, scores) -> None:
    """snorm.py:164-166 / :180-182 — ``print(utt1, utt2, score)`` with score an np.float32."""
    with open(path, "w") as f:
        f.write("".join("%s %s %s\n" % (a, b, str(np.float32(s))) for a, b, s in scores))


def score_files(trial: str, test_ark: str, cosine_score: str, test_spk2utt: Optional[str] = None,
                cohort_ark: Optional[str] = None, cohort_spk2utt: Optional[str] = None, weight_matrix: Optional[str] = None,
                snorm_score: Optional[str] = None, topk: int = 400, device: int = 0) -> None:
    """The whole of snorm.py's ``__main__`` (snorm.py:155-182) with embeddings resident on the GPU between
    the stages: one upload of the test matrix, index-pair trials, one gather kernel for both score files."""
    sc = default_scorer(device)
    dv = _dev(device)
    keys, mat = kaldi_ark.read_vec_ark_matrix(test_ark)
    x = sc.l2norm(torch.from_numpy(mat).to(dv))
    names = list(dict.fromkeys(keys))
    if len(names) != len(keys):                       # duplicate keys: the dict keeps the last vector at the first position
        last = {k: i for i, k in enumerate(keys)}
        x = x[torch.tensor([last[k] for k in names], device=dv)]
    index = {k: i for i, k in enumerate(names)}
    if test_spk2utt is not None:                      # snorm.py:157-160
        spk2utt = read_spk2utt(test_spk2utt)
        utt_to_spk = {u: s for s, us in spk2utt.items() for u in us}
        spk_index: Dict[str, int] = {}
        group = np.full(len(names), -1, np.int32)
        for i, k in enumerate(names):
            s = utt_to_spk.get(k)
            if s is not None:
                group[i] = spk_index.setdefault(s, len(spk_index))
        if spk_index:
            means = sc.group_means(x, torch.from_numpy(group).to(dv), len(spk_index))
            extra_rows = []
            for s, g in spk_index.items():            # dict.update: existing keys are overwritten in place
                if s in index:
                    x[index[s]] = means[g]
                else:
                    index[s] = len(names)
                    names.append(s)
                    extra_rows.append(g)
            if extra_rows:
                x = torch.cat([x, means[torch.tensor(extra_rows, device=dv)]], dim=0)
    pairs = parse_trials(trial)
    idx1 = torch.from_numpy(np.fromiter((index[a] for a, _ in pairs), np.int32, len(pairs))).to(dv)
    idx2 = torch.from_numpy(np.fromiter((index[b] for _, b in pairs), np.int32, len(pairs))).to(dv)
    mean = std = None
    if snorm_score is not None:
        if cohort_ark is not None and cohort_spk2utt is not None:
            ckeys, cmat = kaldi_ark.read_vec_ark_matrix(cohort_ark)
            cdict_order = list(dict.fromkeys(ckeys))
            clast = {k: i for i, k in enumerate(ckeys)}
            cx = sc.l2norm(torch.from_numpy(cmat).to(dv))
            spk2utt = read_spk2utt(cohort_spk2utt)
            utt_to_spk = {u: s for s, us in spk2utt.items() for u in us}
            spk_index = {}
            rows, group = [], []
            for k in cdict_order:
                s = utt_to_spk.get(k)
                if s is not None:
                    rows.append(clast[k])
                    group.append(spk_index.setdefault(s, len(spk_index)))
            cohort = sc.group_means(cx[torch.tensor(rows, device=dv)], torch.tensor(group, dtype=torch.int32, device=dv),
                                    len(spk_index))
        elif weight_matrix is not None:
            with open(weight_matrix, "rb") as f:
                w = np.asarray(pickle.load(f), dtype=np.float32)
            cohort = sc.l2norm(torch.from_numpy(np.ascontiguousarray(w)).to(dv))
        else:
            raise ValueError("Can not compute snorm scores: no cohort vectors provided")   # snorm.py:174
        mean, std = sc.cohort_mean_std(x, cohort, topk)
    cos, sn = sc.trial_scores(x, idx1, idx2, mean, std)
    cos = cos.cpu().numpy()
    with open(cosine_score, "w") as f:
        f.write("".join("%s %s %s\n" % (a, b, str(s)) for (a, b), s in zip(pairs, cos)))
    if sn is not None:
        sn = sn.cpu().numpy()
        with open(snorm_score, "w") as f:
            f.write("".join("%s %s %s\n" % (a, b, str(s)) for (a, b), s in zip(pairs, sn)))
