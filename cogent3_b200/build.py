"""Build libc3cuda.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m cogent3_b200.build [--force]

The shared library lands next to this package (cogent3_b200/libc3cuda.so) so that
it travels with the repo snapshot to the GPU box.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libc3cuda.so")
SOURCES = [os.path.join(CSRC, "c3cuda.cu")]
HEADERS = [
    os.path.join(CSRC, f)
    for f in ("c3_types.cuh", "expm_kernels.cuh", "prune_kernels.cuh", "lab_kernels.cuh", "compress_kernels.cuh", "hmm_kernels.cuh")
] + [os.path.join(ROOT, "include", "c3cuda.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "--shared", "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden", "--fmad=true", *os.environ.get("C3_NVCC_EXTRA", "").split(),
]  # fmt: skip


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS + [os.path.abspath(__file__)])


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [nvcc_path(), *NVCC_FLAGS, "-I", os.path.join(ROOT, "include"), "-o", LIB, *SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
