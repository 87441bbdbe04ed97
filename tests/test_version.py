"""Table test of the version-string parser against doc/arch-variants.md and davo.py's chain."""
import pytest

from davo_b200 import version as V
from tests.golden import make_golden as G

BASE = "v1-decay100k-sharedNN-dilatedPoseNN-cnv6_128"


def test_headline_variant():
    c = V.parse_version(BASE + "-segmask_all-se_flow-abs_flow-fc_tanh")
    assert (c.posenn, c.cnv6_out, c.in_mode) == (V.POSENN_DECOUPLE_SHARED_DIL, 128, 1)
    assert (c.att_src, c.att_tgt_ones, c.mask_mode) == (V.ATT_SE_FLOW, 1, V.MASK_ALL)
    assert (c.se_act, c.flow_abs, c.flow_norm, c.posenn_se) == (V.ACT_TANH, V.ABS_BOTH, 0, V.PSE_NONE)


@pytest.mark.parametrize("suffix,att,tgt1,mask", [
    ("-no_segmask", V.ATT_NONE, 1, V.MASK_OFF),
    ("-segmask_all-static", V.ATT_STATIC, 1, V.MASK_ALL),
    ("-segmask_rgb-static", V.ATT_STATIC, 1, V.MASK_RGB),
    ("-segmask_all", V.ATT_STATIC, 0, V.MASK_ALL),           # final else of the chain (davo.py:1395)
    ("-segmask_rgb-se_flow", V.ATT_SE_FLOW, 1, V.MASK_RGB),
    ("-segmask_all-se_seg_wo_tgt-fc_tanh", V.ATT_SE_SEG, 1, V.MASK_ALL),       # davo.py:1304-1310
    ("-segmask_all-se_seg-fc_tanh", V.ATT_SE_SEG, 0, V.MASK_ALL),              # davo.py:1311-1316
    ("-segmask_all-se_rgb_wo_tgt_to_seg", V.ATT_SE_RGB_SEG, 1, V.MASK_ALL),    # davo.py:1274-1283
    ("-segmask_rgb-se_rgb_to_seg", V.ATT_SE_RGB_SEG, 0, V.MASK_RGB),           # davo.py:1284-1292
    ("-segmask_all-se_SegFlow_to_seg_8_wo_tgt", V.ATT_SE_SEGFLOW_SEG, 1, V.MASK_ALL),   # davo.py:1341-1349
    ("-segmask_all-se_SegFlow_to_seg_8", V.ATT_SE_SEGFLOW_SEG, 0, V.MASK_ALL),          # davo.py:1350-1357
    ("-segmask_rgb-se_SegFlow_to_seg_wo_tgt", V.ATT_SE_SEGFLOW_SEG, 1, V.MASK_RGB),     # davo.py:1358-1366
    ("-segmask_all-se_SegFlow_to_seg", V.ATT_SE_SEGFLOW_SEG, 0, V.MASK_ALL),            # davo.py:1367-1374
])
def test_attention_and_mask_modes(suffix, att, tgt1, mask):
    c = V.parse_version(BASE + suffix)
    assert (c.att_src, c.att_tgt_ones, c.mask_mode) == (att, tgt1, mask)


def test_segflow_bottleneck_width_follows_the_token():
    """se(..., layer_channels=[8,19]) for the "_8" tokens, [19,19] otherwise (davo.py:1345, 1354, 1362, 1371)."""
    widths = {"-se_SegFlow_to_seg_8_wo_tgt": 8, "-se_SegFlow_to_seg_8": 8, "-se_SegFlow_to_seg_wo_tgt": 19, "-se_SegFlow_to_seg": 19}
    for tok, hid in widths.items():
        assert V.parse_version(BASE + "-segmask_all" + tok).se_hidden == hid


def test_pyramid_pooling_tokens():
    """se(flow, "se_flow", [8,19], mode='spp', spp_size=...) (davo.py:1193-1210): first match wins; the variables
    live under pose_exp_net/se_flow, so the target map is ones as for -se_flow (davo.py:1404-1412)."""
    for tok, pool in (("-se_spp21_flow", V.SE_POOL_SPP21), ("-se_spp2_flow", V.SE_POOL_SPP2),
                      ("-se_spp_flow", V.SE_POOL_SPP864), ("-se_spp864_flow", V.SE_POOL_SPP864)):
        c = V.parse_version(BASE + "-segmask_all" + tok)
        assert (c.att_src, c.att_tgt_ones, c.se_pool) == (V.ATT_SE_FLOW, 1, pool), tok


def test_label_map_pooling_tokens():
    """se_block(seg_19, "se_seg", mode='gp2x2') and se_spp_block(seg_19, "se_spp_seg", spp_size) (davo.py:1317-1340):
    the target frame gets its own map (no _wo_tgt form exists, nothing lives under se_flow)."""
    for tok, pool in (("-se_gp2x2_seg", V.SE_POOL_GP2X2), ("-se_spp21_seg", V.SE_POOL_SPP21), ("-se_spp_seg_21", V.SE_POOL_SPP21),
                      ("-se_spp2_seg", V.SE_POOL_SPP2), ("-se_spp_seg", V.SE_POOL_SPP864), ("-se_spp864_seg", V.SE_POOL_SPP864)):
        c = V.parse_version(BASE + "-segmask_all" + tok)
        assert (c.att_src, c.att_tgt_ones, c.se_pool) == (V.ATT_SE_SEG, 0, pool), tok


def test_disparity_sources():
    """se(1. / depth, "se_disp", [8,19]) (davo.py:1246-1270): reads input_depth, ignores -norm_depth."""
    for tok, tgt1 in (("-se_disp_wo_tgt_to_seg", 1), ("-se_disp_to_seg", 0)):
        c = V.parse_version(BASE + "-segmask_all" + tok + "-norm_depth")
        assert (c.att_src, c.att_tgt_ones, c.depth_norm, c.needs_depth) == (V.ATT_SE_DEPTH_SEG, tgt1, 2, 1)


def test_per_pixel_sources():
    """se_block sources whose map is reduce_sum(input * excitation) (davo.py:1228-1245, 1271-1303, 1375-1379)."""
    table = {"-se_rgb": (V.ATT_SE_RGB_SEG, 0, 0), "-se_rgb_wo_tgt": (V.ATT_SE_RGB_SEG, 1, 0),
             "-se_depth": (V.ATT_SE_DEPTH_SEG, 0, 0), "-se_depth_wo_tgt": (V.ATT_SE_DEPTH_SEG, 1, 0),
             "-se_disp": (V.ATT_SE_DEPTH_SEG, 0, 2), "-se_disp_wo_tgt": (V.ATT_SE_DEPTH_SEG, 1, 2),
             "-se_mixSegFlow": (V.ATT_SE_SEGFLOW_SEG, 0, 0)}
    for tok, (att, tgt1, dn) in table.items():
        c = V.parse_version(BASE + "-segmask_all" + tok)
        assert (c.att_src, c.att_tgt_ones, c.pixel_map, c.depth_norm) == (att, tgt1, 1, dn), tok
    assert V.parse_version(BASE + "-segmask_all-se_rgb_to_seg").pixel_map == 0
    c = V.parse_version(BASE + "-segmask_all-se_mixDepthFlow-norm_depth")            # davo.py:1157-1165
    assert (c.att_src, c.att_tgt_ones, c.pixel_map, c.depth_norm, c.needs_depth) == (V.ATT_SE_DEPTH_SEG, 0, 2, 1, 1)
    c = V.parse_version(BASE + "-segmask_all-se_mixDispFlow-norm_depth")             # davo.py:1166-1174
    assert (c.att_src, c.att_tgt_ones, c.pixel_map, c.depth_norm, c.needs_depth) == (V.ATT_SE_DEPTH_SEG, 0, 2, 2, 1)
    c = V.parse_version("v1-dilatedPoseNN-cnv6_128-segmask_all-se_rgb")              # the non-shared nets take them too
    assert (c.posenn, c.att_src, c.pixel_map) == (V.POSENN_DECOUPLE_DIL, V.ATT_SE_RGB_SEG, 1)
    with pytest.raises(UnboundLocalError):                                           # capital "Disp": input_depth is never read (davo.py:960, 1167)
        V.parse_version(BASE + "-segmask_all-se_mixDispFlow")
    c = V.parse_version("v1-dilatedPoseNN-cnv6_128-segmask_all-se_flow_on_depthseg_seplayers")   # the depth split in a non-shared net
    assert (c.posenn, c.depth_split, c.att_tgt_ones) == (V.POSENN_DECOUPLE_DIL, 1, 1)


def test_order_sensitive_tokens():
    assert V.parse_version(BASE + "-se_flow-abs_flow_h").flow_abs == V.ABS_H      # _h before bare token
    assert V.parse_version(BASE + "-se_flow-abs_flow_v").flow_abs == V.ABS_V
    assert V.parse_version(BASE + "-se_flow-abs_flow").flow_abs == V.ABS_BOTH
    assert V.parse_version(BASE + "-se_flow-fc_lrelu").se_act == V.ACT_LRELU
    assert V.parse_version(BASE + "-se_flow-fc_relu").se_act == V.ACT_RELU        # no test for it: default
    assert V.parse_version(BASE + "-se_flow-norm_flow").flow_norm == 1


def test_version_tag_and_cnv6():
    assert V.parse_version("sharedNN-x-sharedNN-dilatedPoseNN").in_mode == 0      # no ^v tag -> v0
    assert V.parse_version("v0-sharedNN-dilatedPoseNN").in_mode == 0
    assert V.parse_version("v1.555-sharedNN-dilatedPoseNN-segmask_all-se_flow").mask_mode == V.MASK_ALL_555
    assert V.parse_version("v1-sharedNN-dilatedPoseNN-cnv6_64").cnv6_out == 64
    assert V.parse_version("v1-sharedNN-dilatedPoseNN").cnv6_out == 128


def test_posenn_selection_and_errors():
    assert V.parse_version("v1-dilatedPoseNN").posenn == V.POSENN_DECOUPLE_DIL
    assert V.parse_version("v1-dilatedCouplePoseNN").posenn == V.POSENN_COUPLE_DIL
    assert V.parse_version("v1-couplePoseNN").posenn == V.POSENN_COUPLE
    assert V.parse_version("v1").posenn == V.POSENN_DECOUPLE
    assert V.parse_version("v1-sharedNN-dilatedCouplePoseNN").posenn == V.POSENN_COUPLE_SHARED_DIL
    with pytest.raises(NameError, match="not support `-sharedNN-couplePoseNN' mode."):
        V.parse_version("v1-sharedNN-couplePoseNN")
    with pytest.raises(NameError, match="unknown PoseNN type."):
        V.parse_version("v1-sharedNN")
    with pytest.raises(AssertionError):
        V.parse_version(None)


def test_all_six_posenn_kinds_resolve_as_in_the_reference():
    """davo.py:1027-1049: -sharedNN {-dilatedPoseNN | -dilatedCouplePoseNN}, else -dilatedPoseNN,
    -dilatedCouplePoseNN, -couplePoseNN, default decouple_net_v0."""
    assert V.parse_version("v1-sharedNN-dilatedPoseNN").posenn == V.POSENN_DECOUPLE_SHARED_DIL
    assert V.parse_version("v1-sharedNN-dilatedCouplePoseNN").posenn == V.POSENN_COUPLE_SHARED_DIL
    assert V.parse_version("v1-dilatedPoseNN").posenn == V.POSENN_DECOUPLE_DIL
    assert V.parse_version("v1-dilatedCouplePoseNN").posenn == V.POSENN_COUPLE_DIL
    assert V.parse_version("v1-couplePoseNN").posenn == V.POSENN_COUPLE
    assert V.parse_version("v1-cnv6_128").posenn == V.POSENN_DECOUPLE


def test_posenn_internal_se_tokens():
    assert V.parse_version(BASE + "-no_segmask-se_insert").posenn_se == V.PSE_INSERT
    assert V.parse_version(BASE + "-se_replace").posenn_se == V.PSE_REPLACE


def test_skipadd_and_batch_norm_tokens():
    """-se_skipadd (posenn.py:229-233) type-checks in the reference only with -cnv6_256 (cnv5 + se_block(cnv6)): the
    same ValueError otherwise.  -batch_norm (posenn.py:206) is a flag of the config; not built next to an SE block."""
    c = V.parse_version("v1-sharedNN-dilatedPoseNN-cnv6_256-segmask_all-se_flow-se_skipadd")
    assert (c.posenn_se, c.cnv6_out) == (V.PSE_SKIPADD, 256)
    with pytest.raises(ValueError, match="Dimensions must be equal"):
        V.parse_version(BASE + "-segmask_all-se_skipadd-fc_tanh")               # BASE has -cnv6_128
    assert V.parse_version("v1-couplePoseNN-cnv6_256-no_segmask-se_skipadd").posenn_se == V.PSE_SKIPADD   # cnv6 at stride 1 there
    assert V.parse_version(BASE + "-segmask_all-se_flow-batch_norm").batch_norm == 1
    assert V.parse_version(BASE + "-segmask_all-se_flow").batch_norm == 0
    c = V.parse_version(BASE + "-no_segmask-se_insert-batch_norm")
    assert (c.batch_norm, c.posenn_se) == (1, V.PSE_INSERT)


def test_depth_variants():
    c = V.parse_version("v1-dilatedPoseNN-cnv6_128-segmask_all-se_depth_wo_tgt-fc_tanh")     # per-pixel source in a non-shared net
    assert (c.posenn, c.att_src, c.pixel_map, c.att_tgt_ones) == (V.POSENN_DECOUPLE_DIL, V.ATT_SE_DEPTH_SEG, 1, 1)
    c = V.parse_version(BASE + "-segmask_all-se_depth_wo_tgt_to_seg-norm_depth-fc_tanh")     # davo.py:1211-1219
    assert (c.att_src, c.att_tgt_ones, c.depth_norm, c.needs_depth) == (V.ATT_SE_DEPTH_SEG, 1, 1, 1)
    assert V.parse_version(BASE + "-segmask_all-se_depth_to_seg").att_tgt_ones == 0


def test_seglabelid_fails_the_way_the_reference_graph_does():
    """davo.py:1069-1073 leaves pred_info with three entries (zip with the three label maps) and
    davo.py:1442 reads pred_info[3]: the reference cannot build its inference graph with -seglabelid."""
    for ver in (BASE + "-seglabelid-segmask_all-se_flow", "v0-sharedNN-dilatedPoseNN-seglabelid", BASE + "-seglabelid-no_segmask"):
        with pytest.raises(IndexError, match="list index out of range"):
            V.parse_version(ver)


def test_a_version_with_two_faults_stops_at_the_reference_s_first_one():
    """Evaluation order of the reference: PoseNN type (davo.py:1027-1049), attention-source chain (:1117-1400), masking
    (:1415-1450, where -seglabelid fails), PoseNN build (posenn.py:233).  Each string is also run through the reference's
    own code where the checkout exists (make_golden.REFERENCE_RAISES, tests/test_tf_shim.py)."""
    for ver, exc in G.REFERENCE_RAISES.items():
        with pytest.raises(Exception) as e:
            V.parse_version(ver)
        assert type(e.value).__name__ == exc, (ver, e.value)


def test_depthseg_tokens_fail_the_way_the_reference_does():
    """davo.py:1117-1156: `_sharedlayers` reads depth_thres before assignment (:1119); the bare token raises the
    reference's NameError (:1155-1156); `_seplayers` is built."""
    with pytest.raises(UnboundLocalError, match="depth_thres"):
        V.parse_version(BASE + "-segmask_all-se_flow_on_depthseg_sharedlayers_15")
    with pytest.raises(NameError, match="please select"):
        V.parse_version(BASE + "-segmask_all-se_flow_on_depthseg")
    c = V.parse_version(BASE + "-segmask_all-se_flow_on_depthseg_seplayers_15")       # davo.py:1136-1154
    assert (c.att_src, c.att_tgt_ones, c.depth_split, c.needs_depth) == (V.ATT_SE_FLOW, 1, 1, 1)



def test_depth_token_next_to_a_source_that_ignores_depth():
    """davo.py:960, 991-996, 1108-1111: "depth" in the version makes the graph slice input_depth; a source that never
    uses it (-se_seg, -se_flow, static ...) is unaffected -- the same config as without the token."""
    a = V.parse_version(BASE + "-segmask_all-se_seg-norm_depth-fc_tanh")
    b = V.parse_version(BASE + "-segmask_all-se_seg-fc_tanh")
    assert a.needs_depth == 1 and b.needs_depth == 0
    da, db = a.as_dict(), b.as_dict()
    for k in ("needs_depth", "depth_norm"):
        da.pop(k), db.pop(k)
    assert da == db
