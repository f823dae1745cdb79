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


def source_hash() -> str:
    """sha256 over the contents of every source the library is built from and the compiler flags: compiled into the
    library (zm_build_info), so a binary that does not belong to the sources in the tree is recognised whatever the
    file times say."""
    import hashlib
    h = hashlib.sha256()
    for d in DEPS:
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()[:32]


def built_hash() -> str | None:
    """The source hash recorded inside the built library (read from the file, without loading it)."""
    if not os.path.exists(LIB):
        return None
    with open(LIB, "rb") as f:
        blob = f.read()
    i = blob.find(b"ZMSRCHASH:")
    return blob[i + 10:i + 42].decode(errors="replace") if i >= 0 else None


def needs_build() -> bool:
    return built_hash() != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    extra = os.environ.get("ZM_NVCC_EXTRA", "").split()       # experiments only, e.g. -DPL_WARPS=8
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
          [f"-DZM_SOURCE_HASH=\"{source_hash()}\"", "-o", LIB, SRC]
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
