"""Build libaimnet2_b200.so (sm_100a) in-tree with nvcc.  `python -m aimnetcentral_b200.build [--force]`"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
LIB = os.path.join(OUT_DIR, "libaimnet2_b200.so")
SOURCES = ["engine.cu", "nblist.cu", "conv.cu", "conv_dense.cu", "gemm.cu", "gemm_tc.cu", "gemm_tc16.cu", "gemm_tc16p.cu", "gemm_tc16d.cu", "gemm_tc16c.cu", "pointwise.cu", "lr.cu", "ewald.cu", "seams.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _digest() -> str:
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [
        os.path.join(os.path.dirname(HERE), "include", "aimnet2_b200.h"), os.path.abspath(__file__)]
    for f in files:
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile if the sources changed.  Safe under torchrun (one process per GPU importing at once): an exclusive file
    lock serialises the ranks, objects and the library are written to temporary names and renamed into place."""
    import fcntl

    os.makedirs(OUT_DIR, exist_ok=True)
    stamp = os.path.join(OUT_DIR, "build.sha256")
    digest = _digest()

    def fresh():
        return os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == digest

    if not force and fresh():
        return LIB
    if not os.path.exists(NVCC):
        raise RuntimeError(f"nvcc not found at {NVCC}; cannot build {LIB}")
    with open(os.path.join(OUT_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and fresh():   # another rank built it while we waited
                return LIB
            return _build_locked(stamp, digest, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(stamp: str, digest: str, verbose: bool) -> str:

    def compile_one(src):
        obj = os.path.join(OUT_DIR, src.replace(".cu", ".o"))
        tmp = obj + f".tmp{os.getpid()}"
        cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", tmp]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        os.replace(tmp, obj)
        with open(os.path.join(OUT_DIR, src.replace(".cu", ".ptxas.txt")), "w") as fh:
            fh.write("".join(ln for ln in r.stderr.splitlines(True) if "Compile time" not in ln))   # keep the log diff-stable
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp_lib = LIB + f".tmp{os.getpid()}"
    cmd = [NVCC, "-shared", "-o", tmp_lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp_lib, LIB)   # atomic: a concurrent CDLL sees the old or the new library, never a partial one
    with open(stamp + ".tmp", "w") as fh:
        fh.write(digest)
    os.replace(stamp + ".tmp", stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
