"""Host build of the composed-step source -- TEST INFRASTRUCTURE ONLY.

`pymgrid_b200/csrc/mg_compose.cu` (+ `mg_compose_step.h`, the per-env arithmetic) compiles for the host with
-DMGC_HOSTSIM: the kernel launch becomes a loop over tiles and threads and every pointer is a host pointer.  The CPU
suite runs that build under the package's own Python host layer (`select` below swaps the one function through which
`pymgrid_b200.compose` opens its library) to exercise the SAME C-ABI, layout validation and per-env code as the GPU,
against vectors recorded from the live reference -- without a GPU.  -ffp-contract=off is the
host-side twin of nvcc's -fmad=false.  Nothing in the package loads this library; the product path needs the CUDA build.
"""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
SRC = os.path.join(_ROOT, "pymgrid_b200", "csrc", "mg_compose.cu")
DEPENDS = [SRC, os.path.join(_ROOT, "pymgrid_b200", "csrc", "mg_compose_step.h"),
           os.path.join(_ROOT, "include", "pymgrid_b200_compose.h"), os.path.join(_ROOT, "include", "pymgrid_b200.h")]
LIB = os.path.join(_HERE, "_lib", "libmgc_hostsim.so")


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(f) for f in DEPENDS):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-Wall", "-x", "c++", "-DMGC_HOSTSIM", "-fPIC",
                           "-shared", "-I", os.path.join(_ROOT, "include"), "-I", os.path.dirname(SRC), "-o", LIB, SRC])
    return LIB


_PRODUCT_OPEN = None


def select(lib):
    """Point `pymgrid_b200.compose` at `lib` -- a ctypes handle of the host build, batches then live in host memory -- or,
    with None, back at the product's own CUDA library.  The swap lives here, in the test tree: the package has no
    parameter, environment variable or fallback that selects a CPU build."""
    import torch
    from pymgrid_b200 import compose
    global _PRODUCT_OPEN
    if _PRODUCT_OPEN is None:
        _PRODUCT_OPEN = compose._open_device_library
    if lib is None:
        compose._open_device_library = _PRODUCT_OPEN
    else:
        compose._open_device_library = lambda device=None: (compose.bind(lib), torch.device("cpu"))
