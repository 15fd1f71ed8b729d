"""Build the CUDA extension in-tree: pymgrid_b200/_lib/libpymgrid_b200.so (sm_100a, nvcc cross-compiles without a GPU)."""
import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
SRC = os.path.join(_PKG, "csrc", "mg_engine.cu")                    # the pymgrid25 / MicrogridGenerator module set
SRC_COMPOSE = os.path.join(_PKG, "csrc", "mg_compose.cu")           # any module list (include/pymgrid_b200_compose.h)
SOURCES = [SRC, SRC_COMPOSE]
HEADER = os.path.join(_ROOT, "include", "pymgrid_b200.h")
DEPENDS = SOURCES + [HEADER, os.path.join(_ROOT, "include", "pymgrid_b200_compose.h"),
                     os.path.join(_PKG, "csrc", "mg_compose_step.h")]
LIB_DIR = os.path.join(_PKG, "_lib")
LIB = os.path.join(LIB_DIR, "libpymgrid_b200.so")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-fmad=false",           # bit-exact parity with the reference's un-fused f64 arithmetic
              "-Xcompiler", "-fPIC", "-shared"]


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def is_stale():
    if not os.path.exists(LIB):
        return True
    return os.path.getmtime(LIB) < max(os.path.getmtime(f) for f in DEPENDS if os.path.exists(f))


def build(force=False, verbose=False, extra=()):
    if not force and not is_stale():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    extra = list(extra) + os.environ.get("PYMGRID_B200_NVCC_EXTRA", "").split()
    cmd = [nvcc_path(), *NVCC_FLAGS, *extra, "-I", os.path.join(_ROOT, "include"), "-o", LIB, *SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
