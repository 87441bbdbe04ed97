"""In-tree build of libdavo_b200.so (sm_100a only) with nvcc.

The shared library is written next to its sources (``davo_b200/csrc``) so it
travels with a snapshot of the repo; it is git-ignored.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libdavo_b200.so")
SOURCES = ["davo_capi.cu", "host_convert.cpp"]


def _deps():
    import glob
    return (glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) +
            glob.glob(os.path.join(CSRC, "*.cpp")) + glob.glob(os.path.join(CSRC, "*.h")) +
            [os.path.join(HERE, "..", "include", "davo_b200.h")])

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-shared", "-ldl",
    "-Xptxas", "-v",
]


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; libdavo_b200.so cannot be built")
    return nvcc


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the library.  Safe under torchrun: ranks serialise on a file lock, the first one compiles into a
    temporary file and renames it into place (a rank can never dlopen a half-written .so), the others find it
    fresh when they get the lock."""
    import fcntl
    if not force and not is_stale():
        return LIB
    with open(os.path.join(CSRC, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():          # another rank built it while this one waited
                return LIB
            tmp = "%s.%d.tmp" % (LIB, os.getpid())
            cmd = [find_nvcc()] + NVCC_FLAGS + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", tmp]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed building libdavo_b200.so")
            os.replace(tmp, LIB)
            with open(os.path.join(CSRC, "build.log"), "w") as f:
                f.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
