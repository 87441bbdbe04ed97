"""Experiment: how much of a step is launch gaps?  The same step as stream launches (DAVO_B200_GRAPH=0 keeps
the host wrapper from replaying a graph of its own) and captured by the caller into one CUDA graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DAVO_B200_GRAPH"] = "0"
import torch
from davo_b200 import synthetic as S
from davo_b200.davo import DAVO
ver = "v1-decay100k-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_flow-abs_flow-fc_tanh"
for B in (1, 8, 128):
    inputs = [torch.as_tensor(x).cuda() for x in S.make_inputs(B, 128, 416)]
    sysm = DAVO(version=ver)
    sysm.setup_inference(128, 416, "davo", 3, B, inputs[0], input_flow=inputs[1], input_seglabel=inputs[2], device=0)
    sysm.load_weights(S.init_weights(ver))
    def run():
        return sysm.inference(None, "pose", as_torch=True)["pose"]
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    def timeit(fn, n=100):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    t_stream = timeit(run)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        run()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = run()
    t_graph = timeit(g.replay)
    print("B=%d: stream launches %.4f ms, CUDA graph %.4f ms per step" % (B, t_stream, t_graph), flush=True)
    sysm.close()
