"""Build libsvx.so (sm_100a only) in-tree with nvcc.  No JIT cache: the .so travels with the repo snapshot."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libsvx.so")
DBG_LIB_PATH = os.path.join(HERE, "libsvx_dbg.so")   # debug build: SVX_* environment switches, CTA-pair instantiation, 2 s watchdog
SOURCES = ["conv_flat.cu", "res2_chain.cu", "conv_pair.cu", "conv_umma.cu", "conv_simple.cu", "elementwise.cu", "frontend.cu", "scoring.cu", "asnorm_fused.cu", "metrics.cu", "model.cu", "api.cu"]
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libsvx cannot be built")


def _stale(lib_path: str = LIB_PATH) -> bool:
    if not os.path.exists(lib_path):
        return True
    t = os.path.getmtime(lib_path)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "svx.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False, debug: bool = False) -> str:
    """``debug``: the library with the SVX_* tuning / knock-out switches compiled in (libsvx_dbg.so; select it with SVX_LIB=…).
    The production libsvx.so reads no environment variable."""
    lib_path = DBG_LIB_PATH if debug else LIB_PATH
    if not force and not _stale(lib_path):
        return lib_path
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build_dbg" if debug else "build")
    extra = ["-DSVX_DEBUG_SWITCHES", "-DSVX_ENABLE_PAIR", "-DSVX_WATCHDOG_MS=2000"] if debug else []
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + ARCH_FLAGS + COMMON + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        with open(obj + ".ptxas.log", "w") as f:
            f.write(r.stderr)
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc] + ARCH_FLAGS + ["-shared", "-o", lib_path] + objs + ["-cudart", "static", "-Xlinker", "--no-undefined",
                                                                   "-ldl", "-lpthread", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return lib_path


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv, debug="--debug" in sys.argv))
