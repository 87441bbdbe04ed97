"""Diagnostic run on a GPU box: per-layer error of the direct and tcgen05 paths vs the oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from davo_b200 import synthetic as S
from davo_b200.davo import DAVO
from oracle import davo_oracle as O

ver = sys.argv[1] if len(sys.argv) > 1 else "v1-decay100k-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_flow-abs_flow-fc_tanh"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
H, W = 128, 416
w = S.init_weights(ver, random_bias=True)
img, flow, seg = S.make_inputs(B, H, W, bad_label_frac=0.01)
taps = {}
t = time.time()
ref = O.davo_forward(ver, img, flow, seg, w, torch.float64, tf32=True, taps=taps)
ref_exact = O.davo_forward(ver, img, flow, seg, w, torch.float64)
print("oracle s", time.time() - t, flush=True)

def rel(a, b):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))

sysm = DAVO(version=ver)
dimg, dflow, dseg = (torch.as_tensor(x).cuda() for x in (img, flow, seg))
sysm.setup_inference(H, W, "davo", 3, B, dimg, input_flow=dflow, input_seglabel=dseg)
sysm.load_weights(w)
for impl in (1, 0):
    sysm._debug_set_conv_impl(impl)
    try:
        out = sysm.inference(None, "pose")["pose"]
        torch.cuda.synchronize()
    except Exception as e:
        print("impl", impl, "FAILED:", e, flush=True)
        break
    print("== impl", "direct" if impl else "tcgen05", "launches", sysm.last_launch_count())
    print(" pose max abs err vs tf32-oracle %.3e  vs exact %.3e  (max |pose| %.3e)" % (
        np.abs(out - ref).max(), np.abs(out - ref_exact).max(), np.abs(ref).max()))
    for p in range(2 * B):
        b, k = p // 2, p % 2
        tp = taps["pair%d" % k]
        aw = sysm.get_intermediate("att_weights", p)
        if taps["attention_weights"] is not None:
            print("  pair", p, "att_w", rel(aw, taps["attention_weights"][1 + k][b]), end=" ")
        pk = sysm.get_intermediate("packed", p)
        if pk.size == H * W * 8:
            print("packed8 %.2e" % rel(pk.reshape(H, W, 8), tp["input"][b][..., [0, 1, 2, 5, 6, 7, 8, 9]]), end=" ")
        else:
            print("packed16 %.2e" % rel(pk.reshape(H, W, 16)[..., :10], tp["input"][b]), end=" ")
        for i, name in enumerate(["cnv1", "cnv2", "cnv3", "cnv4", "cnv5"]):
            g = sysm.get_intermediate(name, p)
            print(name, "%.2e" % rel(g, tp[name][b]), end=" ")
        g6 = sysm.get_intermediate("cnv6", p).reshape(32, 104, 256)
        print("cnv6r %.2e" % rel(g6[..., :128], tp["cnv6_rotation"][b]), "cnv6t %.2e" % rel(g6[..., 128:], tp["cnv6_translation"][b]), end=" ")
        s7 = sysm.get_intermediate("cnv7_sum", p).reshape(2, 256)
        print("sum7r %.2e" % rel(s7[0], tp["cnv7_rotation"][b].sum((0, 1))), "sum7t %.2e" % rel(s7[1], tp["cnv7_translation"][b].sum((0, 1))))
print("done")
