"""max |pose - fp64 oracle| of the headline variant on the golden inputs (B=2) and on 8 more samples."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from davo_b200 import synthetic as S
from davo_b200.davo import DAVO
from oracle import davo_oracle as O
ver = "v1-decay100k-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_flow-abs_flow-fc_tanh"
w = S.init_weights(ver, seed=8964, random_bias=True)
for B, seed in ((2, 1234), (8, 77)):
    inputs = S.make_inputs(B, 128, 416, seed=seed, bad_label_frac=0.01)
    s = DAVO(version=ver)
    s.setup_inference(128, 416, "davo", 3, B, device=0)
    s.load_weights(w)
    dev = tuple(torch.as_tensor(x).cuda() for x in inputs)
    got = s.inference(None, "pose", inputs=dev)["pose"].astype(np.float64)
    host = s.inference(None, "pose", inputs=inputs)["pose"].astype(np.float64)
    ref = O.davo_forward(ver, *inputs, w, torch.float64)
    q = tuple([inputs[0], inputs[1].astype(np.float16).astype(np.float32), inputs[2]])
    refq = O.davo_forward(ver, *q, w, torch.float64)
    print("B=%d: max|gpu-oracle64| = %.3e (pose scale %.2e); host entry identical: %s; oracle(flow as binary16) vs oracle: %.3e"
          % (B, np.abs(got - ref).max(), np.abs(ref).max(), np.array_equal(got, host), np.abs(refq - ref).max()), flush=True)
