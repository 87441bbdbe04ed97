"""Randomised search for divergences between the reference's graph code and this repo's restatements.

Draws version strings from the reference's grammar (net type, cnv6 width, input version, attention
source, masking, activation, normalisations, PoseNN-internal SE, -batch_norm, -seglabelid), and for each one runs

  * the REFERENCE (davo.py / nets/*.py unmodified over tests/tf_shim, float64), and
  * ``davo_b200.version.parse_version`` + ``synthetic.init_weights`` + ``oracle.davo_forward`` (float64)

on the same seeded 64x208 inputs.  Both must either raise the same exception type or agree on the poses to 1e-9
relative.  The hand-picked committed cases of make_golden.py were picked by hand; this looks where nobody picked.

    python tests/golden/fuzz_versions.py [n=200] [seed=0] [--sizes] [--feature]

Needs the reference checkout (here only).  Divergences are printed and the exit status is their count; a clean run of
the seeds recorded in DESIGN.md section 2 is part of the pinning evidence.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import make_golden as G                                    # noqa: E402

NETS = ["", "-sharedNN-dilatedPoseNN", "-sharedNN-dilatedCouplePoseNN", "-dilatedPoseNN", "-dilatedCouplePoseNN",
        "-couplePoseNN", "-sharedNN-couplePoseNN", "-sharedNN"]
NET_P = [0.12, 0.3, 0.15, 0.13, 0.1, 0.12, 0.04, 0.04]
SOURCES = ["", "-se_flow", "-se_gp2x2_flow_nobottle", "-se_gp2x2_flow", "-se_spp21_flow", "-se_spp2_flow", "-se_spp_flow",
           "-se_spp864_flow", "-se_depth_wo_tgt_to_seg", "-se_depth_to_seg", "-se_depth_wo_tgt", "-se_depth",
           "-se_disp_wo_tgt_to_seg", "-se_disp_to_seg", "-se_disp_wo_tgt", "-se_disp", "-se_rgb_wo_tgt_to_seg",
           "-se_rgb_to_seg", "-se_rgb_wo_tgt", "-se_rgb", "-se_seg_wo_tgt", "-se_seg", "-se_gp2x2_seg", "-se_spp21_seg",
           "-se_spp_seg_21", "-se_spp2_seg", "-se_spp_seg", "-se_spp864_seg", "-se_SegFlow_to_seg_8_wo_tgt",
           "-se_SegFlow_to_seg_8", "-se_SegFlow_to_seg_wo_tgt", "-se_SegFlow_to_seg", "-se_mixSegFlow",
           "-se_spp21_mixSegFlow", "-se_mixDepthFlow", "-se_mixDispFlow", "-se_flow_on_depthseg_seplayers_40",
           "-se_flow_on_depthseg_seplayers", "-se_flow_on_depthseg_sharedlayers", "-se_flow_on_depthseg"]
MASKS = ["", "-no_segmask", "-segmask_all", "-segmask_rgb", "-segmask_all-static", "-segmask_rgb-static", "-segmask", "-static"]


def draw(rng):
    v = rng.choice(["v0", "v1", "v1", "v1", "v1.555", "v2", ""])
    parts = [v, rng.choice(["", "-decay100k"]), rng.choice(NETS, p=NET_P),
             rng.choice(["", "-cnv6_64", "-cnv6_128", "-cnv6_256", "-cnv6_32"], p=[0.15, 0.2, 0.4, 0.2, 0.05]),
             rng.choice(MASKS)]
    if rng.random() < 0.8:
        parts.append(rng.choice(SOURCES))
    tail = []
    for tok, p in (("-abs_flow", 0.35), ("-abs_flow_h", 0.08), ("-abs_flow_v", 0.08), ("-norm_flow", 0.3), ("-norm_depth", 0.3),
                   ("-fc_tanh", 0.4), ("-fc_lrelu", 0.2), ("-seglabelid", 0.06), ("-se_insert", 0.08), ("-se_skipadd", 0.08),
                   ("-se_replace", 0.08), ("-batch_norm", 0.1)):
        if rng.random() < p:
            tail.append(tok)
    rng.shuffle(tail)
    s = "".join(parts + tail)
    return s[1:] if s.startswith("-") else s


def ours(ver, img, flow, seg, depth):
    import torch
    from davo_b200 import synthetic as S
    from oracle import davo_oracle as O
    w = S.init_weights(ver, seed=G.GOLDEN["weight_seed"], random_bias=True)
    return O.davo_forward(ver, img, flow, seg, w, torch.float64, depth=depth), w


def run(n, seed, verbose=True, sizes=False, feature=False):
    """-> (list of (version, what differs), summary line).  sizes: a random frame size (multiples of 8, H <= W) and
    batch per string instead of 64x208 x 2.  feature: inference(mode='feature') -- attention maps, masked frames, the
    upsampled cnv6 -- instead of the poses alone."""
    from davo_b200 import synthetic as S
    rng = np.random.default_rng(seed)
    H, W, B = 64, 208, 2
    img, flow, seg = S.make_inputs(B, H, W, seed=5, bad_label_frac=0.01)
    depth = S.make_depth(B, H, W)
    seen, bad, built, raised = set(), [], 0, {}
    t0 = time.time()
    with G.reference_on_path():
        while len(seen) < n:
            ver = draw(rng)
            if ver in seen:
                continue
            seen.add(ver)
            if sizes:
                H = int(rng.integers(4, 26)) * 8
                W = max(H, int(rng.integers(8, 60)) * 8)
                B = int(rng.integers(1, 4))
                img, flow, seg = S.make_inputs(B, H, W, seed=int(rng.integers(1 << 30)), bad_label_frac=0.01)
                depth = S.make_depth(B, H, W)
            try:
                mine, w = ours(ver, img, flow, seg, depth)
                mine_exc = None
            except Exception as e:  # noqa: BLE001
                mine, w, mine_exc = None, {}, type(e).__name__
            try:
                out, _ = G.run_reference(ver, img, flow, seg, depth, w, "float64", "feature" if feature else "pose")
                ref, ref_exc = np.asarray(out["pose"], np.float64), None
                if feature and mine_exc is None:
                    import torch
                    from oracle import davo_oracle as O
                    want, got = G.feature_digest(out), G.feature_digest(O.davo_features(ver, img, flow, seg, w, torch.float64, depth=depth))
                    for name in want:
                        a, b = np.asarray(got[name], np.float64), np.asarray(want[name], np.float64)
                        if a.shape != b.shape or not np.allclose(a, b, rtol=1e-9, atol=1e-12):
                            bad.append((ver, "%dx%d x%d: feature %s differs" % (H, W, B, name)))
                            if verbose:
                                print("DIVERGENCE", bad[-1], flush=True)
                            break
            except AssertionError as e:
                # the reference BUILT its graph; with no weight set to feed (ours raised) only the variable check fails
                ref, ref_exc = None, ("builds" if "variable surface mismatch" in str(e) else "AssertionError")
            except Exception as e:  # noqa: BLE001
                ref, ref_exc = None, type(e).__name__
            if mine_exc or ref_exc:
                same = mine_exc == ref_exc
                raised[ref_exc] = raised.get(ref_exc, 0) + 1
                if not same:
                    bad.append((ver, "%dx%d x%d: ours raises %s, the reference %s" % (H, W, B, mine_exc, ref_exc)))
                    if verbose:
                        print("DIVERGENCE", bad[-1], flush=True)
                continue
            built += 1
            err = np.abs(mine - ref).max() / max(np.abs(ref).max(), 1e-30)
            if not (mine.shape == ref.shape and err < 1e-9):
                bad.append((ver, "%dx%d x%d: poses differ: max rel %.3e" % (H, W, B, err)))
                if verbose:
                    print("DIVERGENCE", bad[-1], flush=True)
    summary = ("seed %d: %d version strings in %.0f s: %d built and agree, both raise %s, %d divergences"
               % (seed, len(seen), time.time() - t0, built - sum(1 for b in bad if "poses" in b[1]),
                  {k: v for k, v in raised.items()}, len(bad)))
    return bad, summary


if __name__ == "__main__":
    sizes_, feature_ = "--sizes" in sys.argv, "--feature" in sys.argv
    sys.argv = [a for a in sys.argv if a not in ("--sizes", "--feature")]
    bad_, summary_ = run(int(sys.argv[1]) if len(sys.argv) > 1 else 200, int(sys.argv[2]) if len(sys.argv) > 2 else 0, sizes=sizes_,
                         feature=feature_)
    print(summary_)
    sys.exit(min(len(bad_), 100))
