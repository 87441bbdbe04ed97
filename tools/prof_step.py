"""Tiny driver for ncu: a few forwards of one pass of B samples (default 128 = 256 frame pairs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from davo_b200 import synthetic as S
from davo_b200.davo import DAVO
ver = "v1-decay100k-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_flow-abs_flow-fc_tanh"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
inputs = [torch.as_tensor(x).cuda() for x in S.make_inputs(B, 128, 416)]
system = DAVO(version=ver)
system.setup_inference(128, 416, "davo", 3, B, inputs[0], input_flow=inputs[1], input_seglabel=inputs[2], device=0)
system.load_weights(S.init_weights(ver))
for _ in range(iters):
    out = system.inference(None, "pose", as_torch=True)["pose"]
torch.cuda.synchronize()
print("ok", out[0, 0].tolist(), "launches/forward", system.last_launch_count())
