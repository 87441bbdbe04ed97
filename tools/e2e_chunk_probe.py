"""End-to-end rate (pinned host numpy in, numpy out, 128 samples per call) against the host entry point's chunk size
(DAVO_B200_HOST_CHUNK), for the reference's float32 inputs and for the compact inputs.  python tools/e2e_chunk_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from davo_b200 import synthetic as S
from davo_b200.davo import DAVO
ver = "v1-decay100k-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_flow-abs_flow-fc_tanh"
w = S.init_weights(ver)
img, flow, seg = S.make_inputs(128, 128, 416, seed=3)
c_flow, c_seg = S.compact_inputs(flow, seg)
pin = lambda x: torch.as_tensor(x).pin_memory().numpy()
h = [pin(x) for x in (img, flow, seg)]
c = [h[0], pin(c_flow), pin(c_seg)]
for chunk in (8, 16, 24, 32, 64):
    os.environ["DAVO_B200_HOST_CHUNK"] = str(chunk)
    sysm = DAVO(version=ver)
    sysm.setup_inference(128, 416, "davo", 3, 128, device=0)
    sysm.load_weights(w)
    out = []
    for name, inp in (("float32", h), ("compact", c)):
        for _ in range(3):
            sysm.inference(None, "pose", inputs=tuple(inp))
        t0 = time.perf_counter()
        for _ in range(40):
            sysm.inference(None, "pose", inputs=tuple(inp))
        out.append("%s %.1f k pairs/s" % (name, 256 * 40 / (time.perf_counter() - t0) / 1e3))
    print("chunk %3d samples: %s" % (chunk, ", ".join(out)), flush=True)
    del sysm
