"""In-tree build of libugaitnet_b200.so (nvcc, sm_100a only).

The shared library is built next to this file so that it travels with the repo snapshot
to the GPU box; nothing is installed into site-packages and there is no JIT cache.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libugaitnet_b200.so")
STAMP = os.path.join(HERE, ".build_stamp")

SOURCES = ["abi.cu", "simt.cu", "elementwise.cu", "triplet.cu", "knn.cu", "knn_tc.cu", "tc.cu", "gaitset.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
    "-diag-suppress", "550",
]


def _digest() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for fn in sorted(os.listdir(root)):
            if fn.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, fn), "rb") as f:
                    h.update(fn.encode())
                    h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def build(force: bool = False, verbose: bool = True) -> str:
    """Compile every CUDA source for sm_100a into libugaitnet_b200.so (no-op when fresh)."""
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as f:
            if f.read().strip() == dig:
                return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        print("[ugaitnet_b200] building:", " ".join(cmd), file=sys.stderr, flush=True)
    subprocess.run(cmd, check=True)
    with open(STAMP, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    print(LIB)
