"""Builds libzmconv_b200.so in-tree (sm_100a only; nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "zmconv_b200.cu")
import glob
DEPS = sorted(glob.glob(os.path.join(HERE, "csrc", "*.cu")) + glob.glob(os.path.join(HERE, "csrc", "*.cuh")) +
              glob.glob(os.path.join(HERE, "csrc", "*.h"))) + [os.path.join(os.path.dirname(HERE), "include", "zmconv_b200.h")]
LIB = os.path.join(HERE, "libzmconv_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",          # no implicit FMA contraction: bit-parity with the CPU restatement
    "-Xcompiler", "-fPIC", "-shared",
    "-diag-suppress", "128",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    extra = os.environ.get("ZM_NVCC_EXTRA", "").split()       # experiments only, e.g. -DPL_WARPS=8
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, SRC]
    env = dict(os.environ)
    env.pop("CXX", None)
    env.pop("CC", None)     # the image exports a gcc wrapper that nvcc should not pick up
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
