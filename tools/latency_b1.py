"""BASELINE.json configs[0]: batch-1 latency of the headline variant (one sample = 2 frame pairs; trajectory mode =
1 pair), device-resident (CUDA events, graph replay as DAVO.inference does from the third identical call) and
host-fed (numpy in, numpy out, wall clock).  DAVO_B200_SMALL_TILES=0 switches the 128-pixel latency plans off.
    python tools/latency_b1.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from davo_b200 import synthetic as S
from davo_b200.davo import DAVO
ver = "v1-decay100k-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_flow-abs_flow-fc_tanh"
w = S.init_weights(ver)
img, flow, seg = S.make_inputs(1, 128, 416, seed=7)
dev = [torch.as_tensor(x).cuda() for x in (img, flow, seg)]
pinned = [torch.as_tensor(x).pin_memory().numpy() for x in (img, flow, seg)]
out = {}
for small in ("1", "0"):
    os.environ["DAVO_B200_SMALL_TILES"] = small
    for graph in ("1", "0"):
        os.environ["DAVO_B200_GRAPH"] = graph
        sysm = DAVO(version=ver)
        sysm.setup_inference(128, 416, "davo", 3, 1, dev[0], input_flow=dev[1], input_seglabel=dev[2], device=0)
        sysm.load_weights(w)
        for pairs in ("all", "trajectory"):
            for _ in range(10):
                ref = sysm.inference(None, "pose", as_torch=True, pairs=pairs)["pose"]
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(500):
                sysm.inference(None, "pose", as_torch=True, pairs=pairs)
            e1.record()
            torch.cuda.synchronize()
            out["device small_tiles=%s graph=%s pairs=%s ms" % (small, graph, pairs)] = round(e0.elapsed_time(e1) / 500, 5)
        if graph == "1":
            sysm.inference(None, "pose", as_torch=True)
            torch.cuda.synchronize()
            lm, npairs = sysm.profile_layers(iters=50)
            out["layers small_tiles=%s (each kernel alone, %d pairs) us" % (small, npairs)] = {k: round(1e3 * v, 2) for k, v in lm.items()}
            for _ in range(5):
                sysm.inference(None, "pose", inputs=tuple(pinned))
            t0 = time.perf_counter()
            for _ in range(300):
                sysm.inference(None, "pose", inputs=tuple(pinned))
            out["host-fed small_tiles=%s pairs=all ms" % small] = round((time.perf_counter() - t0) / 300 * 1e3, 5)
            t0 = time.perf_counter()
            for _ in range(300):
                sysm.inference(None, "pose", inputs=tuple(pinned), pairs="trajectory")
            out["host-fed small_tiles=%s pairs=trajectory ms" % small] = round((time.perf_counter() - t0) / 300 * 1e3, 5)
        out.setdefault("poses", {})[small] = ref.cpu().numpy().tolist()
        del sysm
same = out["poses"]["0"] == out["poses"]["1"]
del out["poses"]
for k, v in out.items():
    print("%-60s %s" % (k, v))
print(json.dumps(out))
assert same, "the two tilings must give the same bits"
