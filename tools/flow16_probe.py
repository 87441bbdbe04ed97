"""Host entry point, 128 samples (256 pairs) per call from pinned numpy arrays: flow as float32 vs as binary16
(CPU conversion in the library), for several conversion thread counts and chunk sizes."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from davo_b200 import synthetic as S
from davo_b200.davo import DAVO
ver = "v1-decay100k-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_flow-abs_flow-fc_tanh"
B = 128
inputs = S.make_inputs(B, 128, 416, seed=3)
pinned = tuple(torch.as_tensor(x).pin_memory().numpy() for x in inputs)
w = S.init_weights(ver)
print("host threads:", os.cpu_count(), flush=True)


def run(env):
    for k, v in env.items():
        os.environ[k] = str(v)
    sysm = DAVO(version=ver)
    sysm.setup_inference(128, 416, "davo", 3, B, device=0)
    sysm.load_weights(w)
    for _ in range(5):
        sysm.inference(None, "pose", inputs=pinned)
    t0 = time.perf_counter()
    for _ in range(40):
        sysm.inference(None, "pose", inputs=pinned)
    dt = (time.perf_counter() - t0) / 40
    h2d = sysm.last_host_copy_bytes()[0]
    sysm.close()
    for k in env:
        os.environ.pop(k)
    print(json.dumps({"env": env, "ms": round(dt * 1e3, 3), "pairs_per_s": round(256 / dt), "h2d_MB": round(h2d / 1e6, 1),
                      "GBps": round(h2d / dt / 1e9, 1)}), flush=True)


run({"DAVO_B200_HOST_FLOW16": 0})
for rep in range(2):
    for fr in (1.0, 0.75):
        run({"DAVO_B200_HOST_FLOW16_FRAC": fr})
        run({"DAVO_B200_HOST_FLOW16_FRAC": fr, "DAVO_B200_HOST_SLICES": 1})
run({"DAVO_B200_HOST_FLOW16": 0})
