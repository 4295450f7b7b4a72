"""Builds ``deepfit_b200/csrc/libdeepfit_b200.so`` in-tree with nvcc for sm_100a.

Plain ``nvcc`` on the ``.cu`` files -- no cmake, no torch extension machinery: the library exports
a C ABI (``include/deepfit_b200.h``) and is loaded with ctypes.  The ``.so`` is git-ignored but
travels to the GPU box with the tree snapshot.
"""
from __future__ import annotations

import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libdeepfit_b200.so")
SOURCES = ["api.cu", "knn.cu", "pca.cu", "wls.cu", "mlp.cu", "pipeline.cu", "textio.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found; deepfit_b200 needs the CUDA toolkit to build its kernels")


def _stamp() -> str:
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))]
    inc = os.path.join(HERE, "..", "include")
    files += [os.path.join(inc, f) for f in sorted(os.listdir(inc)) if f.endswith(".h")]
    for f in files:
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    stamp_file = LIB_PATH + ".stamp"
    if not (os.path.isfile(LIB_PATH) and os.path.isfile(stamp_file)):
        return False
    with open(stamp_file) as fh:
        return fh.read().strip() == _stamp()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a and link the shared library.  Returns its path.

    Safe under torchrun: concurrent callers serialise on a lock file, objects go to a per-process
    directory, and the library and its stamp are moved into place atomically (``os.replace``), so a
    rank never dlopens a half-written file."""
    import fcntl
    if not force and is_current():
        return LIB_PATH
    lock = open(LIB_PATH + ".lock", "w")
    try:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and is_current():          # another process built it while we waited
            return LIB_PATH
        return _build_locked(verbose)
    finally:
        fcntl.flock(lock, fcntl.LOCK_UN)
        lock.close()


def _build_locked(verbose: bool) -> str:
    sources = [s for s in SOURCES if os.path.isfile(os.path.join(CSRC, s))]
    nvcc = _nvcc()
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        objs = list(ex.map(compile_one, sources))
    tmp_lib = "%s.tmp.%d" % (LIB_PATH, os.getpid())
    cmd = [nvcc, "-shared", "-o", tmp_lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    os.replace(tmp_lib, LIB_PATH)
    tmp_stamp = "%s.stamp.tmp.%d" % (LIB_PATH, os.getpid())
    with open(tmp_stamp, "w") as fh:
        fh.write(_stamp())
    os.replace(tmp_stamp, LIB_PATH + ".stamp")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
