"""tests/golden/stream_poses.npz: fp32 reference poses of the seeded 4539-sample stream (4541 frames, the length of
KITTI sequence 00: BASELINE.json configs[2], north_star "composed trajectory ATE within 1e-3 m").

Samples s .. s+63 are ``synthetic.make_inputs(64, 128, 416, seed=1000 + s)`` for s = 0, 64, 128, ...; weights are
``synthetic.init_weights(HEADLINE)`` (TF-default initialisers, seed 8964).  The poses come from the oracle's float32
restatement (what TF's CPU kernels compute, up to summation order), which tests/test_oracle.py holds to the
reference's own graph code at 1e-9; where the reference checkout exists, chunks 0 and 2240 are ALSO run through the
reference's code itself (tests/golden/make_golden.run_reference, float32) and must agree to 2e-8.

    python tests/golden/make_stream_golden.py          # ~5-10 min of CPU
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from davo_b200 import synthetic as S            # noqa: E402
from oracle import davo_oracle as O             # noqa: E402
from tests.golden import make_golden as G       # noqa: E402

VERSION = G.CASES["headline"]
N, CHUNK, H, W = 4539, 64, 128, 416


def stream_chunk(s):
    return S.make_inputs(min(CHUNK, N - s), H, W, seed=1000 + s)


def main():
    w = S.init_weights(VERSION)
    out = np.zeros((N, 2, 6), np.float32)
    t0 = time.time()
    for s in range(0, N, CHUNK):
        inputs = stream_chunk(s)
        out[s:s + len(inputs[0])] = O.davo_forward(VERSION, *inputs, w, torch.float32)
        if s in (0, 2240) and os.path.isdir(G.REF):
            with G.reference_on_path():
                depth = S.make_depth(4, H, W)
                ref, _ = G.run_reference(VERSION, *(a[:4] for a in inputs), depth, w, "float32")
            err = float(np.abs(ref["pose"] - out[s:s + 4]).max())
            print("chunk %d: |oracle32 - reference code (float32)| = %.2e" % (s, err), flush=True)
            assert err < 2e-8
        if s % 640 == 0:
            print("%d / %d samples, %.0f s" % (s, N, time.time() - t0), flush=True)
    np.savez_compressed(os.path.join(HERE, "stream_poses.npz"), pose=out, version=np.array(VERSION),
                        seed0=np.int64(1000), chunk=np.int64(CHUNK))
    traj = O.compose_trajectory(out)
    print("wrote stream_poses.npz; path length %.3f m" % float(np.linalg.norm(np.diff(traj[:, :3, 3], axis=0), axis=1).sum()))


if __name__ == "__main__":
    main()
