"""Experiment: do two handles on two streams fill each other's kernel tails?
One handle runs 256 pairs per step on one stream (the product configuration); the probe runs the same
256 pairs as two handles x 128 pairs on two streams (fork / join with events), and 2 x 256 vs 1 x 512.
Every kernel of a pass is persistent with one CTA per SM, so a lane's CTAs can only start where the
other lane's kernel has already drained: the question is whether that overlap beats the extra tails."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from davo_b200 import synthetic as S
from davo_b200.davo import DAVO

os.environ["DAVO_B200_GRAPH"] = "0"
ver = "v1-decay100k-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_flow-abs_flow-fc_tanh"
H, W = 128, 416
w = S.init_weights(ver)


def make(B, seed):
    t = tuple(torch.as_tensor(x).cuda() for x in S.make_inputs(B, H, W, seed=seed))
    s = DAVO(version=ver)
    s.setup_inference(H, W, "davo", 3, B, t[0], input_flow=t[1], input_seglabel=t[2], device=0)
    s.load_weights(w)
    return s


def timed(fn, steps=100, warm=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def lanes(systems):
    streams = [torch.cuda.Stream() for _ in systems]
    def step():
        cur = torch.cuda.current_stream()
        fork = torch.cuda.Event(); fork.record(cur)
        for s, st in zip(systems, streams):
            st.wait_event(fork)
            with torch.cuda.stream(st):
                s.inference(None, "pose", as_torch=True)
            j = torch.cuda.Event(); j.record(st); cur.wait_event(j)
    return step


out = {}
for total in (128, 256):                                   # samples per step
    one = make(total, 1)
    ms1 = timed(lambda: one.inference(None, "pose", as_torch=True))
    del one
    for n in (2, 4):
        many = [make(total // n, 10 + i) for i in range(n)]
        msn = timed(lanes(many))
        out["%d pairs: 1 lane / %d lanes" % (2 * total, n)] = [round(ms1, 4), round(msn, 4), round(ms1 / msn, 4)]
        del many
print(json.dumps(out))
