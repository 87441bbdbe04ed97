"""One-off (not in the test suite: ~1-2 min of CPU oracle): ATE of the 4541-frame stream,
GPU poses vs fp32 oracle poses, random-init weights (BASELINE.json config 3 / north_star)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from davo_b200 import synthetic as S, geo_utils
from davo_b200.davo import DAVO
from oracle import davo_oracle as O
ver = "v1-decay100k-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_flow-abs_flow-fc_tanh"
n, chunk = 4539, 64
w = S.init_weights(ver)
modes = sys.argv[1:] or ["compensated"]      # DAVO_B200_WEIGHT_ROUNDING values to compare; "noresidual" = compensated without residual channels
systems = {}
for m in modes:
    os.environ["DAVO_B200_WEIGHT_ROUNDING"] = "compensated" if m == "noresidual" else m
    os.environ.pop("DAVO_B200_NO_RESIDUAL", None)
    if m == "noresidual":
        os.environ["DAVO_B200_NO_RESIDUAL"] = "1"
    systems[m] = DAVO(version=ver)
    systems[m].setup_inference(128, 416, "davo", 3, chunk, device=0)
    systems[m].load_weights(w)
gpu, ref = {m: [] for m in modes}, []
t0 = time.time()
for s in range(0, n, chunk):
    b = min(chunk, n - s)
    inputs = S.make_inputs(b, 128, 416, seed=1000 + s)
    for m in modes:
        gpu[m].append(systems[m].inference(None, "pose", inputs=inputs)["pose"])
    ref.append(O.davo_forward(ver, *inputs, w, torch.float32))
ref = np.concatenate(ref)
tr = O.compose_trajectory(ref)
path = float(np.linalg.norm(np.diff(tr[:, :3, 3], axis=0), axis=1).sum())
for m in modes:
    g = np.concatenate(gpu[m])
    tg = geo_utils.compose_trajectory(g)
    print("[%s] samples %d frames %d  max|dpose| %.3e  mean dpose %s  ATE %.3e m  path %.2f m  end-point error %.3e m  (%.0f s)" % (
        m, n, tg.shape[0], np.abs(g - ref).max(), np.array2string((g - ref).mean((0, 1)), precision=2), O.ate(tg, tr), path,
        np.linalg.norm(tg[-1, :3, 3] - tr[-1, :3, 3]), time.time() - t0), flush=True)
