"""The C-ABI library loads on a CPU-only box, exports every symbol include/svx.h declares, and refuses to run
without a GPU instead of falling back to anything."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__
    __graft_entry__.build()
    from voxsrc2020_speaker_verification_b200 import lib
    return lib


def header_functions():
    src = open(os.path.join(ROOT, "include", "svx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(svx_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(built):
    names = header_functions()
    assert len(names) >= 20
    dll = ctypes.CDLL(built.LIB_PATH)
    for n in names:
        assert hasattr(dll, n), "libsvx.so does not export %s" % n
    assert sorted(built.SYMBOLS) == names, "lib.py binds a different set of symbols than svx.h declares"
    assert built.load().svx_version() == 200
    assert built.load().svx_build_flags() == 0        # production build: no environment switches, no CTA-pair instantiation


def test_config_struct_matches_header(built):
    # 3 + 1 + 24 + 4 + 4 + 1 + 4 + 4 + 4 + 4 + 4 + 2 int32 fields (the last two: att_pool, att_dim)
    assert ctypes.sizeof(built.ModelConfigStruct) == 4 * (3 + 1 + 24 + 4 + 4 + 1 + 4 + 4 + 4 + 4 + 4 + 2)


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from voxsrc2020_speaker_verification_b200.extractor import Extractor
    from voxsrc2020_speaker_verification_b200.scoring import Scorer
    with pytest.raises(built.SvxError, match="no CPU fallback"):
        Extractor("tdnn", 40)
    with pytest.raises(built.SvxError, match="no CPU fallback"):
        Scorer(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "voxsrc2020_speaker_verification_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, fn)).read(), fn
