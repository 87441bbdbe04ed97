"""Experiment: cnv7 (3x3 stride 2, 128 -> 256 per branch, spatial-sum epilogue) pixels-on-M (streamed 32-KB weight slabs per
CTA) against channels-on-M with 128-pixel tiles and 2-CTA clusters multicasting the 16-KB weight slabs
(DAVO_B200_CNV7_CM=1).  Prints each layer's time per 256-pair pass and the pose difference.  python tools/cnv7_probe.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from davo_b200 import synthetic as S
from davo_b200.davo import DAVO
ver = "v1-decay100k-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_flow-abs_flow-fc_tanh"
w = S.init_weights(ver)
inputs = [torch.as_tensor(x).cuda() for x in S.make_inputs(128, 128, 416, seed=3)]
res = {}
for knob in ("0", "1"):
    if knob == "1":
        os.environ["DAVO_B200_CNV7_CM"] = "1"
    else:
        os.environ.pop("DAVO_B200_CNV7_CM", None)
    sysm = DAVO(version=ver)
    sysm.setup_inference(128, 416, "davo", 3, 128, inputs[0], input_flow=inputs[1], input_seglabel=inputs[2], device=0)
    sysm.load_weights(w)
    for _ in range(5):
        pose = sysm.inference(None, "pose", as_torch=True)["pose"]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        sysm.inference(None, "pose", as_torch=True)
    e1.record()
    torch.cuda.synchronize()
    lm, n = sysm.profile_layers(iters=30)
    res[knob] = {"step_ms": e0.elapsed_time(e1) / 200, "layers_ms": {k: round(v, 4) for k, v in lm.items()}, "pose": pose.cpu().numpy()}
    print("DAVO_B200_CNV7_CM=%s: step %.4f ms, cnv7 %.4f ms, layers %s" % (knob, res[knob]["step_ms"], lm["cnv7"], res[knob]["layers_ms"]), flush=True)
print("max |pose difference| between the two cnv7 plans: %.3e" % np.abs(res["0"]["pose"] - res["1"]["pose"]).max())
