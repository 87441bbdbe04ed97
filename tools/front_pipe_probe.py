"""Front end of a 256-pair pass: se_pool_kernel + pack8_kernel (two launches) against front_pipeline_kernel (one launch,
pooling a few pairs ahead of packing so that the flow's second read is an L2 hit).  Same bits either way.
    python tools/front_pipe_probe.py"""
import os, subprocess, sys, json
if len(sys.argv) > 1:
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import numpy as np, torch
    from davo_b200 import synthetic as S
    from davo_b200.davo import DAVO
    ver = "v1-decay100k-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_flow-abs_flow-fc_tanh"
    w = S.init_weights(ver)
    out = {}
    for B in (128, 1):
        inputs = [torch.as_tensor(x).cuda() for x in S.make_inputs(B, 128, 416, seed=3, bad_label_frac=0.01)]
        sysm = DAVO(version=ver)
        sysm.setup_inference(128, 416, "davo", 3, B, inputs[0], input_flow=inputs[1], input_seglabel=inputs[2], device=0)
        sysm.load_weights(w)
        for _ in range(5):
            pose = sysm.inference(None, "pose", as_torch=True)["pose"]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 200 if B == 128 else 1000
        for _ in range(n):
            sysm.inference(None, "pose", as_torch=True)
        e1.record()
        torch.cuda.synchronize()
        lm, npairs = sysm.profile_layers(iters=50)
        out["B%d" % B] = {"step_ms": e0.elapsed_time(e1) / n, "front_ms": lm["front"], "launches": sysm.last_launch_count(),
                          "pose_crc": int(np.frombuffer(pose.cpu().numpy().tobytes(), np.uint32).sum())}
    print(json.dumps(out))
else:
    for knob in ("0", "1"):
        env = dict(os.environ, DAVO_B200_FRONT_PIPE=knob)
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        print("DAVO_B200_FRONT_PIPE=%s: %s" % (knob, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-800:]), flush=True)
