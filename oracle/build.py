"""Build the C part of the oracle (gcc).  Output: oracle/_build/liboracle.so
(git-ignored; travels to the GPU box with the repo snapshot)."""

from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "pruning.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle.so")


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) > os.path.getmtime(SRC):
        return LIB
    cmd = ["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-fPIC", "-shared", "-fvisibility=hidden",
           "-o", LIB, SRC, "-lm"]  # fmt: skip
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        # -march may be unsupported on an odd host: retry portable
        cmd.remove("-march=x86-64-v3")
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("gcc failed:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
