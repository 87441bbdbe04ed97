"""Version-string variant selection for the pose forward path.

Mirrors the ordered substring / regex tests of the reference's inference graph
builder (reference ``davo.py:1010-1102`` for the option groups, ``:1117-1400``
for the attention-source chain, ``:1404-1450`` for masking).  First match wins
inside a group; groups are independent.  The result is the small config struct
handed to the C ABI (``include/davo_b200.h``: ``davo_config``).

Bad strings raise the same exception types and messages as the reference
(``NameError`` for PoseNN selection, ``davo.py:1035-1037``).
"""
from __future__ import annotations

import re
from dataclasses import dataclass, asdict

# posenn kinds (reference nets/posenn.py)
POSENN_DECOUPLE_SHARED_DIL = 0   # decouple_sharednet_v0_dilation  :189-254  (headline)
POSENN_COUPLE_SHARED_DIL = 1     # couple_sharednet_v0_dilation    :133-187
POSENN_DECOUPLE_DIL = 2          # decouple_net_v0_dilation        :69-131
POSENN_COUPLE_DIL = 3            # couple_net_v0_dilation          :12-66
POSENN_COUPLE = 4                # couple_net_v0                   :257-311
POSENN_DECOUPLE = 5              # decouple_net_v0                 :314-378

ATT_NONE, ATT_SE_FLOW, ATT_STATIC, ATT_SE_SEG, ATT_SE_RGB_SEG, ATT_SE_DEPTH_SEG, ATT_SE_SEGFLOW_SEG = 0, 1, 2, 3, 4, 5, 6
MASK_OFF, MASK_RGB, MASK_ALL, MASK_ALL_555 = 0, 1, 2, 3
ACT_RELU, ACT_TANH, ACT_LRELU = 0, 1, 2
ABS_NONE, ABS_BOTH, ABS_H, ABS_V = 0, 1, 2, 3
PSE_NONE, PSE_INSERT, PSE_SKIPADD, PSE_REPLACE = 0, 1, 2, 3
# SE pooling of the flow sources (nets/attention_module.py:64-86): global mean, 2x2 quadrants, or the
# spatial pyramid with out_pool_size [2,1] / [2] / [8,6,4]
SE_POOL_GP, SE_POOL_GP2X2, SE_POOL_SPP21, SE_POOL_SPP2, SE_POOL_SPP864 = 0, 1, 2, 3, 4
SPP_SIZES = {SE_POOL_SPP21: (2, 1), SE_POOL_SPP2: (2,), SE_POOL_SPP864: (8, 6, 4)}

# Attention sources of the reference chain that this build does not implement
# (davo.py:1117-1383), in the reference's evaluation order.
_UNBUILT_SOURCES = ()
_UNBUILT_AFTER_SE_FLOW = (
    "-se_depth_wo_tgt_to_seg", "-se_depth_to_seg",
    "-se_depth_wo_tgt", "-se_depth", "-se_disp_wo_tgt_to_seg", "-se_disp_to_seg",
    "-se_disp_wo_tgt", "-se_disp", "-se_rgb_wo_tgt_to_seg", "-se_rgb_to_seg",
    "-se_rgb_wo_tgt", "-se_rgb", "-se_seg_wo_tgt", "-se_seg", "-se_gp2x2_seg",
    "-se_spp21_seg", "-se_spp_seg_21", "-se_spp2_seg", "-se_spp_seg", "-se_spp864_seg",
    "-se_SegFlow_to_seg_8_wo_tgt", "-se_SegFlow_to_seg_8", "-se_SegFlow_to_seg_wo_tgt",
    "-se_SegFlow_to_seg", "-se_mixSegFlow", "-se_spp21_mixSegFlow",
)


# Sources of that chain that ARE built: token -> (att_src, att_tgt_ones).  `_wo_tgt` forces the
# target map to ones (davo.py:1283, 1310); without it the target frame gets its own SE map, and
# since no variable lives under 'pose_exp_net/se_flow' the G11 override does not fire
# (davo.py:1404-1414).
_BUILT_AFTER_SE_FLOW = {
    "-se_depth_wo_tgt_to_seg": (ATT_SE_DEPTH_SEG, 1), # davo.py:1211-1219
    "-se_depth_to_seg": (ATT_SE_DEPTH_SEG, 0),        # davo.py:1220-1227
    "-se_disp_wo_tgt_to_seg": (ATT_SE_DEPTH_SEG, 1, 0, 0, 2),     # davo.py:1253-1262: se(1. / depth, "se_disp", [8,19])
    "-se_disp_to_seg": (ATT_SE_DEPTH_SEG, 0, 0, 0, 2),            # davo.py:1263-1270
    "-se_rgb_wo_tgt_to_seg": (ATT_SE_RGB_SEG, 1),     # davo.py:1274-1283
    "-se_rgb_to_seg": (ATT_SE_RGB_SEG, 0),            # davo.py:1284-1292
    "-se_seg_wo_tgt": (ATT_SE_SEG, 1),                # davo.py:1304-1310
    "-se_seg": (ATT_SE_SEG, 0),                       # davo.py:1311-1316
    # se(concat(seg_19, SE flow), "se_segflow", [8|19, 19]) -> weights gathered by label
    "-se_SegFlow_to_seg_8_wo_tgt": (ATT_SE_SEGFLOW_SEG, 1, 8),    # davo.py:1341-1349
    "-se_SegFlow_to_seg_8": (ATT_SE_SEGFLOW_SEG, 0, 8),           # davo.py:1350-1357
    "-se_SegFlow_to_seg_wo_tgt": (ATT_SE_SEGFLOW_SEG, 1, 19),     # davo.py:1358-1366
    "-se_SegFlow_to_seg": (ATT_SE_SEGFLOW_SEG, 0, 19),            # davo.py:1367-1374
    # se_block / se_spp_block on the one-hot label map with cell-wise pooling (hidden width 19 = ratio 1)
    "-se_gp2x2_seg": (ATT_SE_SEG, 0, 0, SE_POOL_GP2X2),           # davo.py:1317-1322
    "-se_spp21_seg": (ATT_SE_SEG, 0, 0, SE_POOL_SPP21),           # davo.py:1323-1328
    "-se_spp_seg_21": (ATT_SE_SEG, 0, 0, SE_POOL_SPP21),
    "-se_spp2_seg": (ATT_SE_SEG, 0, 0, SE_POOL_SPP2),             # davo.py:1329-1334
    "-se_spp_seg": (ATT_SE_SEG, 0, 0, SE_POOL_SPP864),            # davo.py:1335-1340
    "-se_spp864_seg": (ATT_SE_SEG, 0, 0, SE_POOL_SPP864),
    # se_block sources whose map is reduce_sum(input * excitation) per pixel (no label gather), ratio=1
    "-se_depth_wo_tgt": dict(att_src=ATT_SE_DEPTH_SEG, att_tgt_ones=1, pixel_map=1),    # davo.py:1228-1237
    "-se_depth": dict(att_src=ATT_SE_DEPTH_SEG, att_tgt_ones=0, pixel_map=1),           # davo.py:1238-1245
    "-se_disp_wo_tgt": dict(att_src=ATT_SE_DEPTH_SEG, att_tgt_ones=1, pixel_map=1, depth_norm=2),   # davo.py:1271-1281
    "-se_disp": dict(att_src=ATT_SE_DEPTH_SEG, att_tgt_ones=0, pixel_map=1, depth_norm=2),          # davo.py:1282-1292
    "-se_rgb_wo_tgt": dict(att_src=ATT_SE_RGB_SEG, att_tgt_ones=1, pixel_map=1),        # davo.py:1293-1298
    "-se_rgb": dict(att_src=ATT_SE_RGB_SEG, att_tgt_ones=0, pixel_map=1),               # davo.py:1299-1303
    "-se_mixSegFlow": dict(att_src=ATT_SE_SEGFLOW_SEG, att_tgt_ones=0, pixel_map=1),    # davo.py:1375-1379
    "-se_spp21_mixSegFlow": dict(att_src=ATT_SE_SEGFLOW_SEG, att_tgt_ones=0, pixel_map=1, se_pool=SE_POOL_SPP21),   # davo.py:1380-1383
}


@dataclass
class DavoConfig:
    """Field-for-field the C struct ``davo_config`` (minus H, W, max_batch)."""
    posenn: int = POSENN_DECOUPLE_SHARED_DIL
    cnv6_out: int = 128
    in_mode: int = 1            # 0 = v0 (RGB only), 1 = v1 (RGB + flow)
    att_src: int = ATT_NONE
    att_tgt_ones: int = 1       # target-frame attention map forced to ones
    mask_mode: int = MASK_OFF
    se_act: int = ACT_RELU
    flow_abs: int = ABS_NONE
    flow_norm: int = 0
    posenn_se: int = PSE_NONE
    depth_norm: int = 0         # "-norm_depth" (davo.py:1108-1111); only read by the se_depth sources
    se_pool: int = 0            # se_flow: SE_POOL_* (davo.py:1175-1210)
    depth_split: int = 0        # 1: -se_flow_on_depthseg_seplayers: near / far class-weight tables split by a learned depth threshold
    pixel_map: int = 0          # 1: map = reduce_sum(SE input * excitation) per pixel (-se_rgb, -se_depth, -se_disp, -se_mixSegFlow)
    se_hidden: int = 0          # SE bottleneck width, 0 = default (8; se_seg 19); gp2x2_flow_nobottle: 19
    needs_depth: int = 0        # "depth"/"disp" in the version: the graph reads input_depth (davo.py:960)
    batch_norm: int = 0         # "-batch_norm": slim.batch_norm on every conv but pred, BATCH statistics at test time
    version_tag: str = "v0"

    def as_dict(self):
        return asdict(self)


def parse_version(version: str) -> DavoConfig:
    """Resolve a version string exactly as ``build_pose_test_graph_davo`` does."""
    assert version is not None                                  # davo.py:959
    cfg = DavoConfig()
    cfg.needs_depth = 1 if ("depth" in version or "disp" in version) else 0     # davo.py:960
    cfg.depth_norm = 1 if "-norm_depth" in version else 0                       # davo.py:1108-1111
    # G1 PoseNN-internal SE (davo.py:1010-1017)
    if "-se_insert" in version:
        cfg.posenn_se = PSE_INSERT
    elif "-se_skipadd" in version:
        cfg.posenn_se = PSE_SKIPADD                                 # checked against the cnv6 width below (G3)
    elif "-se_replace" in version:
        cfg.posenn_se = PSE_REPLACE
    # G2 PoseNN type (davo.py:1027-1049)
    if "-sharedNN" in version:
        if "-dilatedPoseNN" in version:
            cfg.posenn = POSENN_DECOUPLE_SHARED_DIL
        elif "-dilatedCouplePoseNN" in version:
            cfg.posenn = POSENN_COUPLE_SHARED_DIL
        elif "-couplePoseNN" in version:
            raise NameError("not support `-sharedNN-couplePoseNN' mode.")
        else:
            raise NameError("unknown PoseNN type.")
    elif "-dilatedPoseNN" in version:
        cfg.posenn = POSENN_DECOUPLE_DIL
    elif "-dilatedCouplePoseNN" in version:
        cfg.posenn = POSENN_COUPLE_DIL
    elif "-couplePoseNN" in version:
        cfg.posenn = POSENN_COUPLE
    else:
        cfg.posenn = POSENN_DECOUPLE
    # G3 cnv6 width (davo.py:1052-1053)
    m = re.search("-cnv6_([0-9]+)", version)
    cfg.cnv6_out = 128 if m is None else int(m.group(1))
    # G4 input version (davo.py:1057-1065)
    m = re.search("^(v[0-9.]+)", version)
    tag = "v0" if m is None else m.group(1)
    cfg.version_tag = tag
    if "v0" in tag:
        cfg.in_mode = 0
    elif "v1" in tag:
        cfg.in_mode = 1
    else:
        cfg.in_mode = 0     # pred_info stays None (davo.py:1060)
    # G5 (davo.py:1066-1073)
    # (-seglabelid: the token is harmless here; the graph fails where the masking reads pred_info, see G12 below)
    # G6 SE activation (davo.py:1077-1085)
    if "-fc_tanh" in version:
        cfg.se_act = ACT_TANH
    elif "-fc_lrelu" in version:
        cfg.se_act = ACT_LRELU
    else:
        cfg.se_act = ACT_RELU
    # G7 (davo.py:1088-1091)
    cfg.flow_norm = 1 if "-norm_flow" in version else 0
    # G8 (davo.py:1094-1102) -- order matters: _h and _v before the bare token
    if "-abs_flow_h" in version:
        cfg.flow_abs = ABS_H
    elif "-abs_flow_v" in version:
        cfg.flow_abs = ABS_V
    elif "-abs_flow" in version:
        cfg.flow_abs = ABS_BOTH
    # G10 attention source (davo.py:1117-1400), reference order
    if "-se_flow_on_depthseg_sharedlayers" in version:
        # davo.py:1118-1119 reads `depth_thres` before anything assigns it: the reference stops here
        raise UnboundLocalError("local variable 'depth_thres' referenced before assignment (reference davo.py:1119)")
    if "-se_flow_on_depthseg" in version and "-se_flow_on_depthseg_seplayers" not in version:
        raise NameError("please select `-se_flow_on_depthseg_seplayers' or `-se_flow_on_depthseg_sharedlayers'.")   # davo.py:1155-1156
    for tok in _UNBUILT_SOURCES:
        if tok in version:
            raise NotImplementedError("davo_b200: attention source %s is not built" % tok)
    if "-se_flow_on_depthseg_seplayers" in version:                             # davo.py:1136-1154
        # se(flow, "se_flow_near" | "se_flow_far", [8,19]) applied to the labels of the pixels nearer / farther than
        # the variable se_flow/depth_threshold; everything lives under pose_exp_net/se_flow*, so the target map is ones
        cfg.att_src, cfg.att_tgt_ones, cfg.depth_split = ATT_SE_FLOW, 1, 1
    elif "-se_mixDepthFlow" in version or "-se_mixDispFlow" in version:         # davo.py:1157-1174
        # se_block(concat(depth term, SE flow), "se_depthflow" | "se_dispflow", ratio=1): a per-pixel map of 3 channels
        if not cfg.needs_depth:
            # "Depth" / "Disp" in these two tokens are capitalised, so davo.py:960's `"depth" in version or "disp" in
            # version` is False unless another token (-norm_depth, ...) says so; the branch then reads
            # se_input_depths (:1161) / pred_depths (:1167), which were never assigned: the reference stops here.
            # Found by running the reference's graph code itself (tests/golden/make_golden.py).
            raise UnboundLocalError("local variable '%s' referenced before assignment (reference davo.py:%d: no "
                                    "lower-case 'depth' or 'disp' in the version, so input_depth is not read)"
                                    % (("se_input_depths", 1161) if "-se_mixDepthFlow" in version else ("pred_depths", 1167)))
        cfg.att_src, cfg.att_tgt_ones, cfg.pixel_map = ATT_SE_DEPTH_SEG, 0, 2
        if "-se_mixDepthFlow" not in version:
            cfg.depth_norm = 2                                   # 1. / depth (davo.py:1167), whatever -norm_depth says
    elif "-se_flow" in version:                                   # davo.py:1175
        cfg.att_src = ATT_SE_FLOW
        cfg.att_tgt_ones = 1                                    # davo.py:1404-1412
    elif "-se_gp2x2_flow_nobottle" in version or "-se_gp2x2_flow" in version:     # davo.py:1181-1192
        # se(flow, "se_flow", [19,19] | [8,19], mode='gp2x2'): the variables live under
        # pose_exp_net/se_flow, so the G11 override (target map := 1) fires as for -se_flow
        cfg.att_src = ATT_SE_FLOW
        cfg.att_tgt_ones = 1
        cfg.se_pool = 1
        cfg.se_hidden = 19 if "-se_gp2x2_flow_nobottle" in version else 8
    elif re.search("-se_spp(21|2|864|)_flow", version):                          # davo.py:1193-1210
        # se(flow, "se_flow", [8,19], mode='spp', spp_size): spatial_pyramid_pool of the SE flow; the variables
        # live under pose_exp_net/se_flow, so the target map is forced to ones as for -se_flow
        cfg.att_src = ATT_SE_FLOW
        cfg.att_tgt_ones = 1
        cfg.se_pool = (SE_POOL_SPP21 if "-se_spp21_flow" in version else
                       SE_POOL_SPP2 if "-se_spp2_flow" in version else SE_POOL_SPP864)
    else:
        chain_hit = None
        for tok in _UNBUILT_AFTER_SE_FLOW:                      # reference order: first match wins
            if tok in version:
                if tok not in _BUILT_AFTER_SE_FLOW:
                    raise NotImplementedError("davo_b200: attention source %s is not built" % tok)
                chain_hit = tok
                break
        if chain_hit is not None:
            hit = _BUILT_AFTER_SE_FLOW[chain_hit]
            if isinstance(hit, dict):
                for key, val in hit.items():
                    setattr(cfg, key, val)
                hit = (cfg.att_src, cfg.att_tgt_ones)
            cfg.att_src, cfg.att_tgt_ones = hit[0], hit[1]
            if len(hit) > 2:
                cfg.se_hidden = hit[2]
            if len(hit) > 3:
                cfg.se_pool = hit[3]
            if len(hit) > 4:
                cfg.depth_norm = hit[4]                              # the disparity sources ignore -norm_depth
        elif "-no_segmask" in version:                          # davo.py:1385
            cfg.att_src = ATT_NONE
            cfg.att_tgt_ones = 1
        elif "-segmask_" in version and "-static" in version:   # davo.py:1390
            cfg.att_src = ATT_STATIC
            cfg.att_tgt_ones = 1
        else:                                                   # davo.py:1395
            cfg.att_src = ATT_STATIC
            cfg.att_tgt_ones = 0
    # G12 masking (davo.py:1415-1450)
    if "-seglabelid" in version:
        # The reference's inference graph cannot be built with this token: davo.py:1069-1073 zips the four
        # pred_info entries with the THREE label maps (or replaces them by the three label maps), and
        # davo.py:1442 (and :1428/:1434 before it) then reads pred_info[3].  Same exception, same reason -- and raised
        # here, after the attention-source chain, because a version with two faults stops at the reference's first one
        # (tests/golden/fuzz_versions.py).
        raise IndexError("list index out of range (reference davo.py:1442: -seglabelid leaves pred_info with 3 entries)")
    if cfg.in_mode == 1:
        if "-segmask_" in version:
            if "-segmask_all" in version and ".555" in tag:
                cfg.mask_mode = MASK_ALL_555
            elif "-segmask_all" in version:
                cfg.mask_mode = MASK_ALL
            else:
                cfg.mask_mode = MASK_RGB
        else:
            cfg.mask_mode = MASK_OFF
    else:
        cfg.mask_mode = MASK_RGB if "-segmask" in version else MASK_OFF
    # (a version that says "depth" / "disp" but selects a source that never touches the depth -- e.g. -se_seg-norm_depth --
    # makes the reference slice input_depth and then ignore it (davo.py:991-996, 1108-1111): needs_depth stays set, the
    # kernels simply do not read it)
    if "-batch_norm" in version:                                # davo.py:1453
        # slim.batch_norm with its default is_training=True (no normalizer_params, posenn.py:206): batch statistics
        # at test time, so the poses depend on which samples share a call
        cfg.batch_norm = 1
    if cfg.posenn_se == PSE_SKIPADD and cfg.cnv6_out != 256:
        # cnv6 = relu(cnv5 + se_block(cnv6)) (posenn.py:229-233): TensorFlow refuses to add 256 and cnv6_out channels.
        # The PoseNN is built after everything above (davo.py:1456), so this is the last fault the reference reports.
        raise ValueError("Dimensions must be equal, but are 256 and %d (reference posenn.py:233: cnv5 + se_cnv6 "
                         "needs -cnv6_256)" % cfg.cnv6_out)
    return cfg
