"""Front-end time per 256-pair pass (se_pool + pack8) for several pack8 grid sizes."""
import os, sys, subprocess, json
if len(sys.argv) > 1:
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    from davo_b200 import synthetic as S
    from davo_b200.davo import DAVO
    ver = "v1-decay100k-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_flow-abs_flow-fc_tanh"
    inputs = [torch.as_tensor(x).cuda() for x in S.make_inputs(128, 128, 416)]
    s = DAVO(version=ver)
    s.setup_inference(128, 416, "davo", 3, 128, inputs[0], input_flow=inputs[1], input_seglabel=inputs[2], device=0)
    s.load_weights(S.init_weights(ver))
    s.inference(None, "pose")
    best = min(s.profile_layers(50)[0]["front"] for _ in range(3))
    print(json.dumps({"pack8_blocks": os.environ.get("DAVO_B200_PACK8_BLOCKS", "52 (default)"), "front_ms": round(best, 4)}))
else:
    for n in ("52", "26", "13", "104", "208"):
        env = dict(os.environ, DAVO_B200_PACK8_BLOCKS=n)
        print(subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True).stdout.strip(), flush=True)
