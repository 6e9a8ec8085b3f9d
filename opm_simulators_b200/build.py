"""In-tree build of libopmb200.so (hand-written CUDA for sm_100a + the C ABI) with nvcc.

The library is built next to its sources (opm_simulators_b200/libopmb200.so): it is git-ignored
but travels to the GPU box with the gpurun snapshot.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("OPMB200_LIB") or os.path.join(HERE, "libopmb200.so")
SOURCES = ["solver.cu", "analysis.cpp"]
DEPS = SOURCES + ["kernels.cuh", "tile_kernels.cuh", "extra_kernels.cuh", "layout.hpp", "../../include/opmb200.h", "../../include/opmb200/property_tree.hpp"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
           "-Xcompiler", "-fPIC,-O3,-Wall", "-shared", "-o", LIB]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    if os.environ.get("OPMB200_PROFILE"):  # in-kernel phase profile of the tile walkers (scripts/prof_tiles.py)
        cmd += ["-DOPMB200_PROFILE"]
    if os.environ.get("OPMB200_TWDBG"):  # b200.debug_timing switches parts of the tile walkers off (timing only)
        cmd += ["-DOPMB200_TWDBG"]
    cmd += os.environ.get("OPMB200_EXTRA_FLAGS", "").split()  # design experiments (-DCW_...=)
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    cmd += ["-lnccl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libopmb200.so")
    if verbose:
        print(r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
