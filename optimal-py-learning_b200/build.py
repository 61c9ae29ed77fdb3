#!/usr/bin/env python
"""Build libopl_b200.so (the C-ABI library of include/opl_b200.h) in-tree with nvcc for sm_100a.

    python optimal-py-learning_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels with the
working tree to the GPU box.
"""
import argparse
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(OUT_DIR, "libopl_b200.so")
OBJ_DIR = os.path.join(HERE, "build")

SOURCES = ["common.cu", "gemm_f64.cu", "gemm_tf32.cu", "ozaki.cu", "mlp.cu", "bfgs.cu", "lbfgs.cu", "rbf.cu", "collective.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built")
    return exe


def _digest():
    h = hashlib.sha256()
    names = sorted(os.listdir(CSRC)) + ["../../include/opl_b200.h"]
    for name in names:
        path = os.path.join(CSRC, name)
        if os.path.isfile(path):
            with open(path, "rb") as fh:
                h.update(name.encode())
                h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS + SOURCES).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile if sources changed since the last build; return the library path."""
    stamp = os.path.join(OUT_DIR, "libopl_b200.sha256")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp):
        with open(stamp) as fh:
            if fh.read().strip() == digest:
                return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return src, obj, res

    objs = []
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        for src, obj, res in pool.map(compile_one, SOURCES):
            if verbose or res.returncode != 0:
                sys.stderr.write(res.stdout)
            if res.returncode != 0:
                raise RuntimeError("nvcc failed on %s" % src)
            objs.append(obj)
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("link failed")
    with open(stamp, "w") as fh:
        fh.write(digest)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    print(build(force=args.force, verbose=args.verbose))
