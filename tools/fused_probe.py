"""cnv1 with the front end fused in (DAVO_B200_FUSED_FRONT=1) against pack8_kernel + cnv1: same bits, and what it costs.
    python tools/fused_probe.py"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANTS = ["headline", "static", "no_segmask", "se_seg", "v0_lrelu", "segmask_rgb", "se_rgb_to_seg", "gp2x2_flow", "batch_norm"]
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np, torch, zlib
    from davo_b200 import synthetic as S
    from davo_b200.davo import DAVO
    from tests.golden import make_golden as G
    out = {"fused": os.environ.get("DAVO_B200_FUSED_FRONT", "0")}
    for key in VARIANTS:
        ver = G.CASES[key]
        w = S.init_weights(ver, random_bias=True)
        for (B, H, W) in ((3, 128, 416), (2, 64, 208), (1, 136, 432)):
            inputs = S.make_inputs(B, H, W, seed=9, bad_label_frac=0.02)
            dev = [torch.as_tensor(x).cuda() for x in inputs]
            s = DAVO(version=ver)
            s.setup_inference(H, W, "davo", 3, B, dev[0], input_flow=dev[1], input_seglabel=dev[2], device=0)
            s.load_weights(w)
            for pairs in (("all",) if key == "batch_norm" else ("all", "trajectory_first")):
                p = s.inference(None, "pose", pairs=pairs)["pose"]
                out["%s %dx%d B%d %s" % (key, H, W, B, pairs)] = [int(zlib.crc32(p.tobytes())), int(s.last_launch_count())]
            if key != "batch_norm":
                ph = s.inference(None, "pose", inputs=inputs)["pose"]
                out["%s %dx%d B%d host" % (key, H, W, B)] = [int(zlib.crc32(ph.tobytes())), 0]
            if key == "headline" and H == 128:
                out["cnv1 crc"] = int(zlib.crc32(s.get_intermediate("cnv1", 1).tobytes()))
                out["packed crc"] = int(zlib.crc32(s.get_intermediate("packed", 1).tobytes()))
            del s
    ver = G.CASES["headline"]
    inputs = [torch.as_tensor(x).cuda() for x in S.make_inputs(128, 128, 416)]
    s = DAVO(version=ver)
    s.setup_inference(128, 416, "davo", 3, 128, inputs[0], input_flow=inputs[1], input_seglabel=inputs[2], device=0)
    s.load_weights(S.init_weights(ver))
    s.inference(None, "pose")
    best = None
    for _ in range(4):
        lm = s.profile_layers(30)[0]
        best = lm if best is None else {k: min(best[k], lm[k]) for k in lm}
    out["layers_ms"] = {k: round(v, 4) for k, v in best.items()}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(5):
        s.inference(None, "pose", as_torch=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(50):
        s.inference(None, "pose", as_torch=True)
    e1.record(); torch.cuda.synchronize()
    out["pairs_per_s_50_steps"] = round(256 * 50 / (e0.elapsed_time(e1) * 1e-3))
    print(json.dumps(out))
else:
    res = {}
    for mode in ("0", "1"):
        env = dict(os.environ, DAVO_B200_FUSED_FRONT=mode)
        r = subprocess.run(["timeout", "400", sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ""
        try:
            res[mode] = json.loads(line)
        except Exception:
            print("mode", mode, "FAILED rc", r.returncode, r.stderr[-1500:])
            res[mode] = None
    if res["0"] and res["1"]:
        bad = [k for k in res["0"] if isinstance(res["0"][k], list) and res["0"][k][0] != res["1"][k][0]]
        print("cases compared:", sum(isinstance(v, list) for v in res["0"].values()), "different bits:", bad)
        print("cnv1 crc equal:", res["0"]["cnv1 crc"] == res["1"]["cnv1 crc"], " packed crc equal:", res["0"]["packed crc"] == res["1"]["packed crc"])
        print("launches (headline all):", res["0"]["headline 128x416 B3 all"][1], "->", res["1"]["headline 128x416 B3 all"][1])
        for m in ("0", "1"):
            print("fused=%s layers_ms %s  %d pairs/s" % (m, res[m]["layers_ms"], res[m]["pairs_per_s_50_steps"]))
