"""ctypes binding of libsvx.so (include/svx.h).  There is no CPU fallback: a missing library or a missing
B200 raises."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_float, c_int, c_int32, c_int64, c_longlong, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SVX_LIB") or os.path.join(HERE, "libsvx.so")   # SVX_LIB: debug override (A/B of two builds)

PRECISION_FP16, PRECISION_BF16 = 0, 1


class SvxError(RuntimeError):
    pass


class ModelConfigStruct(ctypes.Structure):
    """Mirror of ``svx_model_config``."""
    _fields_ = [
        ("family", c_int32), ("feat_dim", c_int32), ("embed_dim", c_int32),
        ("tdnn_layers", c_int32), ("tdnn_filters", c_int32 * 8), ("tdnn_kernels", c_int32 * 8), ("tdnn_dilations", c_int32 * 8),
        ("num_filters", c_int32 * 4), ("width", c_int32 * 4), ("split", c_int32), ("block_sizes", c_int32 * 4),
        ("block_strides", c_int32 * 4),
        ("init_features", c_int32), ("bw", c_int32), ("k_r", c_int32), ("cardinality", c_int32), ("k_sec", c_int32 * 4),
        ("inc_sec", c_int32 * 4),
        ("att_pool", c_int32), ("att_dim", c_int32),
    ]


# every symbol include/svx.h declares: name → (restype, argtypes)
SYMBOLS = {
    "svx_version": (c_int, []),
    "svx_build_flags": (c_int, []),
    "svx_last_error": (c_char_p, []),
    "svx_extractor_create": (c_int, [POINTER(ModelConfigStruct), c_int, c_int, POINTER(c_void_p)]),
    "svx_extractor_destroy": (c_int, [c_void_p]),
    "svx_extractor_num_tensors": (c_int, [c_void_p]),
    "svx_extractor_tensor_info": (c_int, [c_void_p, c_int, POINTER(c_char_p), POINTER(c_int), POINTER(c_int64)]),
    "svx_extractor_set_tensor": (c_int, [c_void_p, c_char_p, c_void_p, c_int, POINTER(c_int64)]),
    "svx_extractor_finalize": (c_int, [c_void_p]),
    "svx_extractor_embed_dim": (c_int, [c_void_p]),
    "svx_extractor_set_option": (c_int, [c_void_p, c_char_p, c_int]),
    "svx_extractor_run_segments": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "svx_extractor_extract": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "svx_extractor_set_dump_dir": (c_int, [c_void_p, c_char_p]),
    "svx_extractor_last_launches": (c_longlong, [c_void_p]),
    "svx_extractor_conv_time": (c_int, [c_void_p, POINTER(ctypes.c_double), POINTER(ctypes.c_double)]),
    "svx_cmvn_sliding": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "svx_decode_compressed": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "svx_scorer_create": (c_int, [c_int, POINTER(c_void_p)]),
    "svx_scorer_destroy": (c_int, [c_void_p]),
    "svx_l2norm_rows": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "svx_group_means": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "svx_asnorm_stats": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "svx_cohort_topk_values": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "svx_topk_stats": (c_int, [c_void_p, c_int, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "svx_trial_scores": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "svx_scorer_last_launches": (c_longlong, [c_void_p]),
    "svx_eer_min_dcf": (c_int, [c_void_p, c_void_p, c_int64, ctypes.c_double, ctypes.c_double, ctypes.c_double, c_void_p, c_void_p]),
    "svx_scorer_set_option": (c_int, [c_void_p, c_char_p, c_int]),
    "svx_scorer_last_path": (c_int, [c_void_p, POINTER(c_longlong), POINTER(c_longlong)]),
}

_lib = None


def load() -> ctypes.CDLL:
    """dlopen libsvx.so and bind every declared symbol.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SvxError("libsvx.so is missing at %s — build it with `python -m voxsrc2020_speaker_verification_b200.build` "
                       "(there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().svx_last_error()
        raise SvxError(msg.decode("utf-8", "replace") if msg else "libsvx call failed with status %d" % status)
