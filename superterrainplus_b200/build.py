"""Builds the CUDA library in-tree: superterrainplus_b200/libshf_b200.so (sm_100a only, -lineinfo for ncu source pages).

nvcc cross-compiles without a GPU. The shared object is git-ignored but travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libshf_b200.so")
SOURCES = [os.path.join(_CSRC, "shf_capi.cu")]
DEPENDS = SOURCES + [os.path.join(_CSRC, "shf_kernels.cuh"), os.path.join(_CSRC, "shf_generic.cuh"), os.path.join(_CSRC, "shf_events.cuh"), os.path.join(_CSRC, "shf_heightfield.cuh"), os.path.join(_CSRC, "shf_biome.cuh"), os.path.join(_HERE, "..", "include", "shf_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-cudart", "static",
]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > built for p in DEPENDS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    out = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or out.returncode != 0:
        sys.stderr.write(out.stdout + out.stderr)
    if out.returncode != 0:
        raise RuntimeError("nvcc failed building libshf_b200.so")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
