"""BASELINE.json configs[4]: batch sweep 1..4096 frame pairs at 128x416 and 256x832 (inputs resident
in HBM, CUDA events, pass size 256 pairs), to map where the conv roofline is reached.
    python tools/sweep.py [max_pairs_128 [max_pairs_256]]  ->  one line per (size, pairs)"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from davo_b200 import synthetic as S
from davo_b200.davo import DAVO
ver = "v1-decay100k-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_flow-abs_flow-fc_tanh"
FLOP = {(128, 416): 7780171776, (256, 832): 4 * 7780171776}
caps = {(128, 416): int(sys.argv[1]) if len(sys.argv) > 1 else 4096,
        (256, 832): int(sys.argv[2]) if len(sys.argv) > 2 else 1024}
w = S.init_weights(ver)
for (H, W), cap in caps.items():
    base = [torch.as_tensor(x).cuda() for x in S.make_inputs(16, H, W, seed=77)]
    pairs = 1
    while pairs <= cap:
        B = max(1, pairs // 2)                       # 1 pair is not expressible: B=1 computes 2
        reps = (B + 15) // 16
        inputs = [torch.cat([t] * reps)[:B].contiguous() for t in base]
        sysm = DAVO(version=ver)
        sysm.setup_inference(H, W, "davo", 3, B, inputs[0], input_flow=inputs[1], input_seglabel=inputs[2], device=0)
        sysm.load_weights(w)
        for _ in range(3):
            sysm.inference(None, "pose", as_torch=True)
        torch.cuda.synchronize()
        iters = max(3, min(200, 4096 // max(B, 1)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            sysm.inference(None, "pose", as_torch=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        rate = 2 * B / (ms * 1e-3)
        print(json.dumps({"size": "%dx%d" % (H, W), "frame_pairs": 2 * B, "ms": round(ms, 4), "pairs_per_s": round(rate, 1),
                          "tflops": round(rate * FLOP[(H, W)] / 1e12, 1)}), flush=True)
        del sysm, inputs
        torch.cuda.empty_cache()
        pairs = max(2, pairs) * 2 if pairs > 1 else 2
