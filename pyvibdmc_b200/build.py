"""Builds libpvd_b200.so (sm_100a) in-tree with nvcc.  `python -m pyvibdmc_b200.build`"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "_lib")
LIB = os.path.join(LIBDIR, "libpvd_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "--use_fast_math=false"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    return "nvcc"


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) +
                  glob.glob(os.path.join(CSRC, "*.inl")) + glob.glob(os.path.join(CSRC, "*.h")) +
                  [os.path.join(HERE, "..", "include", "pvd_b200.h")])


def source_hash():
    """sha256 over the sources the library is built from (what decides staleness: file times do not survive a fresh checkout
    or a snapshot to another machine)."""
    import hashlib
    h = hashlib.sha256()
    for path in sources():
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_stale():
    if not os.path.exists(LIB):
        return True
    stamp = LIB + ".srchash"
    if os.path.exists(stamp):
        try:
            return open(stamp).read().strip() != source_hash()
        except OSError:
            pass
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build_library(force=False, verbose=False):
    """Compile csrc/pvd_b200.cu -> _lib/libpvd_b200.so.  nvcc cross-compiles without a GPU.  Safe when several processes
    (torchrun ranks) call it at once: one of them builds, under a file lock, into a temporary file that replaces the library
    atomically; the others wait and find it fresh."""
    if not force and not is_stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    import fcntl
    with open(os.path.join(LIBDIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():          # another process built it while this one waited
                return LIB
            flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
            tmp = LIB + f".tmp{os.getpid()}"
            cmd = [_nvcc()] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp, os.path.join(CSRC, "pvd_b200.cu")]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
            os.replace(tmp, LIB)
            with open(LIB + ".srchash", "w") as f:
                f.write(source_hash())
            if verbose:
                print(res.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
