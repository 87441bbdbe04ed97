"""Regenerates tests/golden/*.npz from the oracle (fp64) on seeded synthetic inputs.

PARITY UNPINNED: the reference holds no golden vectors for this path and TF 1.13
is not installable, so these fixtures pin the ORACLE (oracle/davo_oracle.py,
cross-checked by oracle/posenn_ref.c), not the reference run itself.  Inputs and
weights are regenerated from seeds (davo_b200/synthetic.py), so only outputs are
stored.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from davo_b200 import synthetic as S          # noqa: E402
from oracle import davo_oracle as O           # noqa: E402

CASES = {
    "headline": "v1-decay100k-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_flow-abs_flow-fc_tanh",
    "no_segmask": "v1-decay100k-sharedNN-dilatedPoseNN-cnv6_128-no_segmask",
    "static": "v1-decay100k-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-static",
    "segmask_rgb": "v1-decay100k-sharedNN-dilatedPoseNN-cnv6_128-segmask_rgb-se_flow-abs_flow-fc_tanh",
    "v0_lrelu": "v0-sharedNN-dilatedPoseNN-cnv6_64-segmask_rgb-se_flow-abs_flow_h-norm_flow-fc_lrelu",
    "se_seg_wo_tgt": "v1-decay100k-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_seg_wo_tgt-fc_tanh",
    "se_seg": "v1-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_seg-fc_tanh",
    "se_rgb_to_seg": "v1-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_rgb_wo_tgt_to_seg-fc_tanh",
    "se_insert": "v1-decay100k-sharedNN-dilatedPoseNN-cnv6_128-no_segmask-se_insert",
    "couple_shared": "v1-sharedNN-dilatedCouplePoseNN-cnv6_128-segmask_all-se_flow-abs_flow-fc_tanh",
    "couple_shared_se_insert": "v1-sharedNN-dilatedCouplePoseNN-cnv6_64-no_segmask-se_insert",
    "se_depth": "v1-decay100k-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_depth_wo_tgt_to_seg-fc_tanh",
    "se_depth_norm_tgt": "v1-sharedNN-dilatedPoseNN-cnv6_128-segmask_rgb-se_depth_to_seg-norm_depth-fc_lrelu",
    "gp2x2_flow": "v1-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_gp2x2_flow-abs_flow-fc_tanh",
    "gp2x2_flow_nobottle": "v1-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_gp2x2_flow_nobottle-norm_flow-fc_tanh",
    "decouple_net": "v1-dilatedPoseNN-cnv6_128-segmask_all-se_flow-abs_flow-fc_tanh",
    "couple_net_v0": "v0-dilatedCouplePoseNN-cnv6_128-segmask_rgb-se_seg-fc_tanh",
    "plain_decouple_net": "v1-cnv6_128-segmask_all-se_flow-abs_flow-fc_tanh",
    "plain_couple_net": "v1-couplePoseNN-cnv6_64-segmask_all-static",
    "segflow_to_seg": "v1-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_SegFlow_to_seg-norm_flow-abs_flow-fc_tanh",
    "se_replace": "v1-decay100k-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_flow-abs_flow-fc_tanh-se_replace",
    "couple_net_se_replace": "v1-dilatedCouplePoseNN-cnv6_64-no_segmask-se_replace",
    "spp864_flow": "v1-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_spp_flow-abs_flow-fc_tanh",
    "spp21_flow_net": "v1-dilatedPoseNN-cnv6_128-segmask_rgb-se_spp21_flow-norm_flow-fc_lrelu",
    "gp2x2_seg": "v1-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_gp2x2_seg-fc_tanh",
    "spp864_seg": "v1-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_spp_seg-fc_tanh",
    "spp21_seg_couple": "v0-sharedNN-dilatedCouplePoseNN-cnv6_64-segmask_rgb-se_spp_seg_21-fc_lrelu",
    "se_disp": "v1-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_disp_to_seg-norm_depth-fc_tanh",
    "pix_rgb": "v1-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_rgb-fc_tanh",
    "pix_depth_wo_tgt": "v1-sharedNN-dilatedPoseNN-cnv6_128-segmask_rgb-se_depth_wo_tgt-norm_depth-fc_lrelu",
    "pix_disp": "v1-sharedNN-dilatedCouplePoseNN-cnv6_64-segmask_all-se_disp-fc_tanh",
    "pix_mix_segflow": "v0-sharedNN-dilatedPoseNN-cnv6_128-segmask_rgb-se_mixSegFlow-abs_flow-norm_flow-fc_tanh",
    "pix_mix_depthflow": "v1-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_mixDepthFlow-norm_depth-abs_flow-fc_tanh",
    "pix_mix_dispflow": "v0-sharedNN-dilatedCouplePoseNN-cnv6_128-segmask_rgb-se_mixDispFlow-norm_flow-fc_lrelu",
    "depthseg_seplayers": "v1-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_flow_on_depthseg_seplayers_40-abs_flow-fc_tanh",
    "spp21_mix_segflow": "v1-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_spp21_mixSegFlow-norm_flow-fc_tanh",
    "segflow_8_wo_tgt": "v0-sharedNN-dilatedCouplePoseNN-cnv6_128-segmask_rgb-se_SegFlow_to_seg_8_wo_tgt-fc_lrelu",
}
GOLDEN = dict(batch=2, height=128, width=416, input_seed=1234, weight_seed=8964, bad_label_frac=0.01)


def main():
    out = {}
    for key, ver in CASES.items():
        w = S.init_weights(ver, seed=GOLDEN["weight_seed"], random_bias=True)
        img, flow, seg = S.make_inputs(GOLDEN["batch"], GOLDEN["height"], GOLDEN["width"],
                                       seed=GOLDEN["input_seed"], bad_label_frac=GOLDEN["bad_label_frac"])
        taps = {}
        depth = S.make_depth(GOLDEN["batch"], GOLDEN["height"], GOLDEN["width"])
        pose = O.davo_forward(ver, img, flow, seg, w, torch.float64, taps=taps, depth=depth)
        out[key + "/pose"] = pose
        if taps["attention_weights"] is not None:
            out[key + "/att_w"] = np.stack(taps["attention_weights"][1:], 1)     # [B,2,19] src0, src1
        # per-layer checksums (mean and mean-abs) of pair 0 / sample 0
        for name in ("input", "cnv1", "cnv2", "cnv3", "cnv4", "cnv5", "cnv6_rotation", "cnv7_translation"):
            if name not in taps["pair0"]:
                continue                         # couple nets have a single branch
            a = taps["pair0"][name][0]
            out[key + "/stat/" + name] = np.array([a.mean(), np.abs(a).mean(), a.max()])
        print(key, pose[0, 0])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "poses.npz"), **out)


if __name__ == "__main__":
    main()
