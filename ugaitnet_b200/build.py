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


OBJ = os.path.join(HERE, "build")
COMPILE_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                 "-Xcompiler", "-fPIC", "-diag-suppress", "550"]


def _headers_digest() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for fn in sorted(os.listdir(root)):
            if fn.endswith((".cuh", ".h")):
                with open(os.path.join(root, fn), "rb") as f:
                    h.update(fn.encode())
                    h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _digest() -> str:
    h = hashlib.sha256(_headers_digest().encode())
    for fn in SOURCES:
        with open(os.path.join(CSRC, fn), "rb") as f:
            h.update(fn.encode())
            h.update(f.read())
    return h.hexdigest()


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def build(force: bool = False, verbose: bool = True) -> str:
    """Compile every CUDA source for sm_100a into libugaitnet_b200.so (no-op when fresh).  The translation units
    are compiled concurrently into build/*.o (each keyed by its own digest, so an edit recompiles one file) and
    linked into the shared library."""
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as f:
            if f.read().strip() == dig:
                return LIB
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(OBJ, exist_ok=True)
    hd = _headers_digest()
    nvcc = nvcc_path()

    def compile_one(src):
        path = os.path.join(CSRC, src)
        with open(path, "rb") as f:
            d = hashlib.sha256(hd.encode() + f.read()).hexdigest()
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        if not force and os.path.exists(obj) and os.path.exists(obj + ".stamp") and open(obj + ".stamp").read() == d:
            return obj
        cmd = [nvcc] + COMPILE_FLAGS + ["-c", "-o", obj, path]
        if verbose:
            print("[ugaitnet_b200] compiling:", " ".join(cmd), file=sys.stderr, flush=True)
        subprocess.run(cmd, check=True)
        with open(obj + ".stamp", "w") as f:
            f.write(d)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-o", LIB] + objs
    if verbose:
        print("[ugaitnet_b200] linking:", " ".join(cmd), file=sys.stderr, flush=True)
    subprocess.run(cmd, check=True)
    with open(STAMP, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    print(LIB)
