"""BASELINE.json configs[3]: the attention-ablation variants at 128 samples (256 frame pairs) per
step, inputs resident in HBM, CUDA events.  One line per variant."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from davo_b200 import synthetic as S
from davo_b200.davo import DAVO
from tests.golden import make_golden as G
H, W, B = 128, 416, 128
inputs = [torch.as_tensor(x).cuda() for x in S.make_inputs(B, H, W, seed=5)]
depth = torch.as_tensor(S.make_depth(B, H, W)).cuda()           # read by the se_depth sources only
for key, ver in G.CASES.items():
    sysm = DAVO(version=ver)
    sysm.setup_inference(H, W, "davo", 3, B, inputs[0], input_flow=inputs[1], input_seglabel=inputs[2], input_depth=depth, device=0)
    sysm.load_weights(S.init_weights(ver))
    for _ in range(3):
        sysm.inference(None, "pose", as_torch=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        sysm.inference(None, "pose", as_torch=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print(json.dumps({"variant": key, "version": ver, "ms_per_step": round(ms, 4), "pairs_per_s": round(2 * B / ms * 1e3, 1),
                      "launches": sysm.last_launch_count()}), flush=True)
    sysm.close()
