"""Per-role stall accounting of the conv kernels (debug build with -DDAVO_TIMING).
Builds a separate libdavo_b200_timing.so (never the product .so), runs one micro-batch,
and prints mean cycles per CTA each role spent waiting, per layer."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from davo_b200 import build as B
LIB = os.path.join(ROOT, "tools", "experiments", "libdavo_b200_timing.so")
if "--build" in sys.argv or not os.path.exists(LIB):
    cmd = [B.find_nvcc()] + [f for f in B.NVCC_FLAGS if f not in ("-Xptxas", "-v")] + ["-DDAVO_TIMING"] + \
          [os.path.join(B.CSRC, s) for s in B.SOURCES] + ["-o", LIB]
    subprocess.run(cmd, check=True)
    if "--build" in sys.argv:
        sys.exit(0)
import numpy as np, torch
B.LIB = LIB                      # load the timing build through the normal wrapper
B.is_stale = lambda: False
from davo_b200 import _capi, synthetic as S
_capi.SYMBOLS = _capi.SYMBOLS
from davo_b200.davo import DAVO
ver = "v1-decay100k-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_flow-abs_flow-fc_tanh"
Bn = int(sys.argv[-1]) if sys.argv[-1].isdigit() else 17
inputs = [torch.as_tensor(x).cuda() for x in S.make_inputs(Bn, 128, 416)]
system = DAVO(version=ver)
system.setup_inference(128, 416, "davo", 3, Bn, inputs[0], input_flow=inputs[1], input_seglabel=inputs[2], device=0)
system.load_weights(S.init_weights(ver))
system.inference(None, "pose")
lib = system._lib
lib.davo_debug_layer_timing.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
names = ["prod:p_empty", "prod:b_empty", "mma:p_full", "mma:b_full", "mma:acc_empty", "epi:acc_full", "epi:busy", "cta_total"]
print("%-6s" % "layer" + "".join("%14s" % n for n in names))
for layer in range(7):
    buf = np.zeros((148, 8), np.int64)
    rc = lib.davo_debug_layer_timing(system._h, layer, buf.ctypes.data, None)
    assert rc == 0, rc
    m = buf.mean(axis=0)
    print("cnv%d  " % (layer + 1) + "".join("%14.0f" % v for v in m))
