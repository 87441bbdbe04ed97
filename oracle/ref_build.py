"""Builds what of the REFERENCE itself compiles here, from its sources where they lie under
/root/reference, into ``oracle/_ref/`` (git-ignored, not gpurun-ignored).  TEST INFRASTRUCTURE ONLY.

The pose forward path of the reference is TensorFlow graph code and cannot be built.  What can:
the KITTI odometry devkit ``kitti_benchmark/cpp/test_odometry_all.cpp`` + ``matrix.cpp`` (recipe
of the reference's own Makefile: ``g++ -O3 -DNDEBUG``).  It is the reference's consumer of the
file the path writes (``<seq>-pred_kitti_pose.txt``, test_kitti_pose.py:150-153), so running it
on our output pins the output format and the pose count against the reference's own reader.
No reference source is copied into this repository.
"""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DAVO_REFERENCE_DIR", "/root/reference")
OUT_DIR = os.path.join(HERE, "_ref")
DEVKIT = os.path.join(OUT_DIR, "test_odometry_all")


def reference_present() -> bool:
    return os.path.isfile(os.path.join(REF, "kitti_benchmark", "cpp", "test_odometry_all.cpp"))


def build(force: bool = False):
    """Returns the path of the devkit binary, or None when the reference tree is not here (the
    GPU box: a binary built earlier in ``oracle/_ref/`` is then used as it is)."""
    if os.path.isfile(DEVKIT) and not force:
        return DEVKIT
    if not reference_present():
        return DEVKIT if os.path.isfile(DEVKIT) else None
    os.makedirs(OUT_DIR, exist_ok=True)
    src = os.path.join(REF, "kitti_benchmark", "cpp")
    subprocess.run(["g++", "-O3", "-DNDEBUG", "-o", DEVKIT, os.path.join(src, "test_odometry_all.cpp"),
                    os.path.join(src, "matrix.cpp")], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return DEVKIT


if __name__ == "__main__":
    print(build(force=True))
