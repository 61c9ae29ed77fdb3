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
# tools only (e.g. OPL_EXTRA_NVCC_FLAGS=-DOPL_HEAD_TRACE_BUILD): part of the digest, so the library is rebuilt when it changes
NVCC_FLAGS += os.environ.get("OPL_EXTRA_NVCC_FLAGS", "").split()


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


def _up_to_date(stamp, digest):
    if os.path.exists(LIB) and os.path.exists(stamp):
        with open(stamp) as fh:
            return fh.read().strip() == digest
    return False


def build(force=False, verbose=False):
    """Compile if sources changed since the last build; return the library path.

    Safe under concurrent callers (one rank per GPU importing the package at the same time): the check-and-build runs
    under an exclusive file lock, objects go to a per-process directory and the library is moved into place atomically, so
    a process never loads a half-written file."""
    import fcntl
    stamp = os.path.join(OUT_DIR, "libopl_b200.sha256")
    digest = _digest()
    if not force and _up_to_date(stamp, digest):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    with open(os.path.join(OUT_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and _up_to_date(stamp, digest):      # another process built it while we waited
                return LIB
            return _build_locked(stamp, digest, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(stamp, digest, verbose):
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
    tmp_lib = LIB + ".tmp%d" % os.getpid()
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp_lib] + objs
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("link failed")
    if os.path.exists(stamp):
        os.remove(stamp)                 # never leave a stamp that describes another binary
    os.replace(tmp_lib, LIB)             # atomic: a concurrent dlopen sees the old or the new file, never a partial one
    with open(stamp + ".tmp", "w") as fh:
        fh.write(digest)
    os.replace(stamp + ".tmp", stamp)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    print(build(force=args.force, verbose=args.verbose))
