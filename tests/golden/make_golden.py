"""Regenerates tests/golden/poses.npz by EXECUTING THE REFERENCE'S OWN GRAPH CODE on seeded synthetic inputs.

The reference's ``davo.py`` (``DAVO.setup_inference`` -> ``build_pose_test_graph_davo``, :955-1494 ->
``inference``, :1553-1569), ``nets/posenn.py``, ``nets/attention_module.py``, ``data_loader.py``
(``batch_unpack_image_sequence``), ``utils/flow_utils.py`` and ``utils/seg_utils/get_dataset_colormap.py``
are imported UNMODIFIED from the reference checkout with ``tests/tf_shim`` first on ``sys.path``: a test-only
``tensorflow`` package that executes the ~60 TensorFlow 1.13 ops they call eagerly on torch-CPU tensors
(float64 = truth; float32 for the uint8 colourings, one rounding per op).  TensorFlow 1.13 itself cannot be
installed here (no wheel for this interpreter, no network).  So the WIRING of every fixture -- frame and channel
order, version-token dispatch, variable names and shapes, scope reuse, masking, the PoseNN topologies, output
assembly -- is the reference's; what stays restated is the semantics of the individual ops (listed in
tests/tf_shim/tensorflow/__init__.py and unit-tested in tests/test_tf_shim.py).

Weights are assigned BY VARIABLE NAME (the part ``Saver.restore`` plays, reference test_kitti_pose.py:129-131):
the generator fails if the reference graph creates a trainable variable that ``synthetic.init_weights`` did
not supply, or if a supplied one is never created -- which pins the checkpoint surface (names + shapes).

Per case the fixture holds: ``pose`` [B,2,6] (float64 run) and ``pose_f32`` (float32 run); ``att_w`` [B,2,n]
(the SE excitation of the two source frames, where the attention is one class-weight vector); ``stat/<scope>``
(mean, mean |x|, max of every slim.conv2d output of the first PoseNN call, sample 0) and ``stat/input``;
``amap`` (the three attention maps, every 8th pixel) and their means; ``feat`` digests of
``inference(mode='feature')``: resized cnv6 (strided samples), masked images, and CRC32s of the uint8 / int
colourings (``flows``, ``segs``) and of ``seg_19``.

Run from the repo root, where the reference checkout exists:   python tests/golden/make_golden.py
(only CASES / GOLDEN are imported by the tests on the GPU box, where the reference is absent).
"""
import contextlib
import io
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("DAVO_REFERENCE_DIR", "/root/reference")
SHIM = os.path.join(ROOT, "tests", "tf_shim")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

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
    # "Disp" is capitalised: without another token holding lower-case "depth" / "disp" the reference stops with an
    # UnboundLocalError (davo.py:960, 1167) -- found by this generator; -norm_depth makes the graph buildable
    "pix_mix_dispflow": "v0-sharedNN-dilatedCouplePoseNN-cnv6_128-segmask_rgb-se_mixDispFlow-norm_depth-norm_flow-fc_lrelu",
    "depthseg_seplayers": "v1-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_flow_on_depthseg_seplayers_40-abs_flow-fc_tanh",
    "spp21_mix_segflow": "v1-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_spp21_mixSegFlow-norm_flow-fc_tanh",
    "segflow_8_wo_tgt": "v0-sharedNN-dilatedCouplePoseNN-cnv6_128-segmask_rgb-se_SegFlow_to_seg_8_wo_tgt-fc_lrelu",
    # the per-pixel (se_block) sources inside the non-shared nets: one evaluation per sample on (tgt, src0, src1)
    "pix_rgb_net": "v1-dilatedPoseNN-cnv6_128-segmask_all-se_rgb-fc_tanh",
    "pix_mix_segflow_net": "v1-dilatedCouplePoseNN-cnv6_128-segmask_rgb-se_mixSegFlow-abs_flow-norm_flow-fc_tanh",
    "pix_mix_depthflow_net": "v0-cnv6_128-segmask_rgb-se_mixDepthFlow-norm_depth-norm_flow-fc_lrelu",
    "pix_disp_wo_tgt_net": "v1-couplePoseNN-cnv6_64-segmask_all-se_disp_wo_tgt-fc_tanh",
    "spp21_mix_segflow_net": "v1-dilatedPoseNN-cnv6_128-segmask_all-se_spp21_mixSegFlow-norm_flow-fc_tanh",
    # -se_skipadd: cnv6 = relu(cnv5 + se_block(cnv6)) (posenn.py:229-233); only type-checks with -cnv6_256
    "se_skipadd": "v1-sharedNN-dilatedPoseNN-cnv6_256-segmask_all-se_flow-abs_flow-fc_tanh-se_skipadd",
    "couple_net_se_skipadd": "v1-dilatedCouplePoseNN-cnv6_256-no_segmask-se_skipadd",
    # -batch_norm: slim.batch_norm with batch statistics at test time (posenn.py:206), BatchNorm/beta instead of biases
    "batch_norm": "v1-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_flow-abs_flow-fc_tanh-batch_norm",
    "batch_norm_couple_net": "v1-dilatedCouplePoseNN-cnv6_64-no_segmask-batch_norm",
    "batch_norm_plain_net": "v0-cnv6_128-segmask_rgb-static-batch_norm",
    "batch_norm_se_insert": "v1-sharedNN-dilatedPoseNN-cnv6_128-no_segmask-se_insert-batch_norm",
    "batch_norm_se_skipadd": "v1-dilatedPoseNN-cnv6_256-segmask_all-se_flow-abs_flow-fc_tanh-se_skipadd-batch_norm",
    "batch_norm_se_replace": "v1-sharedNN-dilatedCouplePoseNN-cnv6_128-segmask_rgb-se_seg-fc_tanh-se_replace-batch_norm",
    # the depth-split source in a non-shared net; PoseNN-internal SE in the two original stride-2 nets (4x13 / 2x7 maps)
    "depthseg_seplayers_net": "v1-dilatedPoseNN-cnv6_128-segmask_all-se_flow_on_depthseg_seplayers_40-abs_flow-fc_tanh",
    "plain_couple_se_insert": "v1-couplePoseNN-cnv6_64-no_segmask-se_insert",
    "plain_decouple_se_replace": "v0-cnv6_128-segmask_rgb-static-se_replace",
    "plain_decouple_se_insert": "v1-cnv6_128-segmask_all-se_flow-abs_flow-fc_tanh-se_insert",
    "plain_decouple_se_skipadd": "v1-cnv6_256-segmask_all-se_flow-abs_flow-fc_tanh-se_skipadd",     # cnv6 at stride 1 (posenn.py:355)
    "plain_couple_se_skipadd": "v0-couplePoseNN-cnv6_256-segmask_rgb-static-se_skipadd",
    # -cnv6_256 on its own (the regex at davo.py:1052 takes any width): a 512-wide fused cnv6 in the dilated shared net,
    # two groups on one input in the stride-2 net, one 256-wide branch in a couple net (found missing by tools/fuzz_gpu.py)
    "cnv6_256": "v1-sharedNN-dilatedPoseNN-cnv6_256-segmask_all-se_flow-abs_flow-fc_tanh",
    "cnv6_256_plain_net": "v0-cnv6_256-segmask_rgb-static",
    "cnv6_256_couple_se_insert": "v1-dilatedCouplePoseNN-cnv6_256-no_segmask-se_insert",
}
# version strings the reference itself cannot build, with the exception its graph code raises (checked by the generator)
REFERENCE_RAISES = {
    "v1-sharedNN-couplePoseNN-cnv6_128-no_segmask": "NameError",
    "v1-sharedNN-cnv6_128-no_segmask": "NameError",
    "v1-sharedNN-dilatedPoseNN-segmask_all-se_flow_on_depthseg_sharedlayers-fc_tanh": "UnboundLocalError",
    "v1-sharedNN-dilatedPoseNN-segmask_all-se_flow_on_depthseg-fc_tanh": "NameError",
    "v1-sharedNN-dilatedPoseNN-segmask_all-se_mixDispFlow-fc_tanh": "UnboundLocalError",
    "v1-sharedNN-dilatedPoseNN-segmask_all-se_mixDepthFlow-fc_tanh": "UnboundLocalError",
    "v1-sharedNN-dilatedPoseNN-segmask_all-se_flow-seglabelid": "IndexError",
    "v1-sharedNN-dilatedPoseNN-cnv6_128-no_segmask-se_skipadd": "ValueError",      # cnv5 (256) + se_block(cnv6) (128)
    # two faults in one string: the reference stops at the first one in ITS evaluation order -- PoseNN type, then the
    # attention-source chain, then the masking (where -seglabelid fails), then the PoseNN build (found by fuzz_versions.py)
    "v1-cnv6_128-segmask-se_depth_wo_tgt_to_seg-seglabelid-se_skipadd-fc_tanh": "IndexError",
    "v1-cnv6_128-segmask-se_mixDispFlow-seglabelid-norm_flow": "UnboundLocalError",
    "v1-dilatedPoseNN-cnv6_64-segmask-se_mixDepthFlow-se_skipadd-norm_flow-fc_tanh": "UnboundLocalError",
    "v1-sharedNN-couplePoseNN-cnv6_128-seglabelid-se_skipadd": "NameError",
}
GOLDEN = dict(batch=2, height=128, width=416, input_seed=1234, weight_seed=8964, bad_label_frac=0.01)
# other frame sizes of the headline variant (BASELINE configs[4]: 256x832; a partial widened run; a width that is not
# a multiple of 16; a small map): (H, W, batch), inputs make_inputs(batch, H, W, seed=11, bad_label_frac=0.01),
# weights init_weights(headline, random_bias=True) -> "size/<H>x<W>/pose"
SIZE_CASES = [(256, 832, 2), (64, 208, 3), (128, 400, 2), (136, 424, 1)]
AMAP_STRIDE, FEAT_STRIDE = 8, (16, 16, 8)


# ------------------------------------------------------------------------------------------------------
# running the reference
# ------------------------------------------------------------------------------------------------------
_REF_MODULES = ("tensorflow", "davo", "data_loader", "nets", "utils", "matplotlib", "png", "cv2")


@contextlib.contextmanager
def reference_on_path():
    """sys.path / sys.modules arranged so that `import davo` is the reference's and `tensorflow` the shim;
    everything is put back afterwards (the repo's own packages are never shadowed outside this block)."""
    if not os.path.isdir(REF):
        raise FileNotFoundError("reference checkout not found at %s" % REF)
    saved_path = list(sys.path)
    saved_mods = {k: v for k, v in sys.modules.items() if k.split(".")[0] in _REF_MODULES}
    for k in saved_mods:
        del sys.modules[k]
    sys.path[:0] = [SHIM, REF]
    try:
        yield
    finally:
        for k in [k for k in sys.modules if k.split(".")[0] in _REF_MODULES]:
            del sys.modules[k]
        sys.modules.update(saved_mods)
        sys.path[:] = saved_path


def run_reference(version, img, flow, seg, depth, weights, float_dtype="float64", mode="pose", quiet=True):
    """DAVO(version).setup_inference(...).inference(sess, mode) of the REFERENCE over the shim.

    Returns (results dict of numpy arrays, info) with info = {"records": {...}, "variables": {name: shape}}.
    Must be called inside ``reference_on_path()``.
    """
    import tensorflow as tf
    from davo import DAVO                                              # the reference's class
    assert "shim" in tf.__version__ and os.path.abspath(sys.modules["davo"].__file__).startswith(os.path.abspath(REF))
    tf.shim_configure(float_dtype)
    tf.shim_reset(feed=weights)
    B, H, W3, _ = img.shape
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink if quiet else sys.stdout):
        system = DAVO(version=version)                                 # test_kitti_pose.py:126
        system.setup_inference(H, W3 // 3, "davo", 3, B,               # test_kitti_pose.py:127-128
                               input_img_uint8=tf.constant(img, dtype=tf.uint8), input_flow=tf.constant(flow),
                               input_depth=tf.constant(depth), input_seglabel=tf.constant(seg))
        out = system.inference(tf.Session(), mode)                     # test_kitti_pose.py:135
    missing, unused = tf.shim_initialised(), tf.shim_unused_feed()
    if missing or unused:
        raise AssertionError("variable surface mismatch for %r: the reference creates %s that the weight set lacks; "
                             "the weight set holds %s that the reference never creates" % (version, missing, unused))
    info = {"records": {k: [o.numpy() for o in v] for k, v in tf.shim_records().items()},
            "variables": {k: tuple(v.t.shape) for k, v in tf.shim_variables().items() if v.trainable}}
    return out, info


# ------------------------------------------------------------------------------------------------------
# digests (shared with the tests that compare the oracle / the CUDA path against the fixture)
# ------------------------------------------------------------------------------------------------------
def stat3(a):
    a = np.asarray(a, np.float64)
    return np.array([a.mean(), np.abs(a).mean(), a.max()])


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def att_w_from_records(records):
    """[B,2,n] excitation of (src0, src1) when the attention is one class-weight vector per frame, else None."""
    scopes = [k for k in records if k.startswith("dense:pose_exp_net/") and k.endswith("/recover_fc")
              and not k.startswith("dense:pose_exp_net/pose/")]
    if len(scopes) != 1 or len(records[scopes[0]]) != 3 or records[scopes[0]][0].shape[-1] != 19:
        return None
    r = records[scopes[0]]                                             # call order: tgt, src0, src1 (davo.py `for i in range(seq_length)`)
    return np.stack([r[1].reshape(r[1].shape[0], -1), r[2].reshape(r[2].shape[0], -1)], 1)


def conv_stats_from_records(records):
    out = {}
    for k, v in records.items():
        if k.startswith("conv2d:pose_exp_net/"):
            out[k[len("conv2d:pose_exp_net/"):].replace("/", ".")] = stat3(v[0][0])
        if k == "conv2d_in:pose_exp_net/cnv1":
            out["input"] = stat3(v[0][0])
    return out


def feature_digest(f):
    """Small, order-sensitive summary of DAVO.inference(mode='feature') (davo.py:1557-1564)."""
    sh, sw, sc = FEAT_STRIDE
    d = {}
    for i, a in enumerate(f["masks"]["attention"]):
        d["amap%d" % i] = np.asarray(a, np.float64)[:, ::AMAP_STRIDE, ::AMAP_STRIDE, 0]
        d["amap%d_mean" % i] = np.asarray(a, np.float64).mean()
    for i, a in enumerate(f["masks"]["image"]):
        d["mimg%d" % i] = np.asarray(a, np.float64)[:, ::sh, ::sw]
    for name in ("rot", "trans"):
        a = np.asarray(f["features"][name], np.float64)
        d["feat_" + name] = a[:, ::sh, ::sw, ::sc]
        d["feat_%s_stat" % name] = stat3(a)
    return d


def colour_digest(f):
    d = {}
    for i, a in enumerate(f["flows"]):
        d["flow_crc%d" % i] = crc(np.asarray(a, np.uint8))
        d["flow_px%d" % i] = np.asarray(a, np.uint8)[:, ::16, ::16]
    for i, a in enumerate(f["segs"]):
        d["seg_crc%d" % i] = crc(np.asarray(a).astype(np.int32))
    for i, a in enumerate(f["seg_19"]):
        d["seg19_crc%d" % i] = crc(np.asarray(a).astype(np.uint8))
    return d


def golden_inputs():
    from davo_b200 import synthetic as S
    g = GOLDEN
    img, flow, seg = S.make_inputs(g["batch"], g["height"], g["width"], seed=g["input_seed"],
                                   bad_label_frac=g["bad_label_frac"])
    return img, flow, seg, S.make_depth(g["batch"], g["height"], g["width"])


def reference_case(key):
    """All fixture entries of one case, computed by the reference's code (inside reference_on_path())."""
    from davo_b200 import synthetic as S
    ver = CASES[key]
    w = S.init_weights(ver, seed=GOLDEN["weight_seed"], random_bias=True)
    img, flow, seg, depth = golden_inputs()
    out = {}
    f64, info = run_reference(ver, img, flow, seg, depth, w, "float64", "feature")
    out["pose"] = np.asarray(f64["pose"], np.float64)
    aw = att_w_from_records(info["records"])
    if aw is not None:
        out["att_w"] = aw
    for name, s in conv_stats_from_records(info["records"]).items():
        out["stat/" + name] = s
    for name, a in feature_digest(f64).items():
        out["feat/" + name] = a
    f32, _ = run_reference(ver, img, flow, seg, depth, w, "float32", "feature")
    out["pose_f32"] = np.asarray(f32["pose"], np.float32)
    for name, a in colour_digest(f32).items():
        out["feat/" + name] = a
    out["nvars"] = np.int64(len(info["variables"]))
    return out


def reference_cli_loop(pred_poses, batch_size):
    """The batch loop of the reference's CLI, EXECUTED FROM ITS SOURCE (test_kitti_pose.py, the lines from
    ``round_num = ...`` to ``recover_pose.append(prev_pose)``, :132-149) on given network outputs.

    pred_poses [N,2,6] are the poses of the N valid samples; the sample list is padded with the reference's
    ``complete_batch_size`` first, as :96-101 does.  ``sess.run(pose_mat_tensor, ...)`` is served by the reference's
    ``utils/geo_utils.pose_vec2mat`` over the shim in float32.  Returns recover_pose [*,4,4] float64.
    Must be called inside ``reference_on_path()``.
    """
    import textwrap
    import types
    import tensorflow as tf
    from utils import geo_utils as ref_geo
    from utils.common_utils import complete_batch_size
    src = open(os.path.join(REF, "test_kitti_pose.py")).read().split("\n")
    a = next(i for i, l in enumerate(src) if l.strip().startswith("round_num = len(image_sequence_names)"))
    b = next(i for i, l in enumerate(src) if l.strip() == "recover_pose.append(prev_pose)")
    code = textwrap.dedent("\n".join(src[a:b + 1]))
    order = complete_batch_size(list(range(len(pred_poses))), batch_size)
    tf.shim_configure("float32")
    tf.shim_reset()

    class System:
        calls = 0

        def inference(self, sess, mode):
            assert mode == "pose"
            sel = order[self.calls * batch_size:(self.calls + 1) * batch_size]
            self.calls += 1
            return {"pose": np.asarray(pred_poses, np.float32)[sel]}

    class Sess:
        def run(self, fetch, feed_dict):
            (vec,) = feed_dict.values()
            return ref_geo.pose_vec2mat(tf.constant(np.asarray(vec, np.float32))).numpy()

    prev_pose = np.eye(4).astype(float)
    ns = dict(np=np, FLAGS=types.SimpleNamespace(batch_size=batch_size), image_sequence_names=order, tgt_inds=order,
              system=System(), sess=Sess(), max_src_offset=1, pose_mat_tensor="pose_mat", pose_vec_ph="pose_vec_ph",
              pred_pose_list=[], prev_pose=prev_pose, recover_pose=[prev_pose])
    exec(compile(code, "test_kitti_pose.py[%d:%d]" % (a + 1, b + 1), "exec"), ns)
    return np.stack(ns["recover_pose"])


def cli_loop_inputs(n=10):
    rng = np.random.default_rng(77)
    p = np.zeros((n, 2, 6), np.float32)
    p[..., :3] = rng.normal(0, 0.02, size=(n, 2, 3))
    p[:, 0, 3:] = rng.normal(0, 0.05, size=(n, 3)) + [0, 0, 0.8]
    p[:, 1, 3:] = rng.normal(0, 0.05, size=(n, 3)) - [0, 0, 0.8]
    return p


def reference_size_case(h, w_, b):
    from davo_b200 import synthetic as S
    ver = CASES["headline"]
    w = S.init_weights(ver, random_bias=True)
    img, flow, seg = S.make_inputs(b, h, w_, seed=11, bad_label_frac=0.01)
    out, _ = run_reference(ver, img, flow, seg, S.make_depth(b, h, w_), w, "float64", "pose")
    return np.asarray(out["pose"], np.float64)


def main(keys=None):
    out = {}
    with reference_on_path():
        for key in (keys or CASES):
            for name, a in reference_case(key).items():
                out[key + "/" + name] = a
            print("%-26s pose[0,0] = %s" % (key, np.array2string(out[key + "/pose"][0, 0], precision=8)), flush=True)
        from davo_b200 import synthetic as S
        img, flow, seg, depth = golden_inputs()
        for ver, exc in REFERENCE_RAISES.items():                      # strings the reference's own code refuses
            try:
                try:
                    w = S.init_weights(ver, seed=GOLDEN["weight_seed"])
                except Exception:
                    w = {}
                run_reference(ver, img[:1], flow[:1], seg[:1], depth[:1], w)
            except AssertionError:
                raise
            except Exception as e:  # noqa: BLE001
                assert type(e).__name__ == exc, (ver, type(e).__name__, exc)
                print("reference raises %-18s for %s" % (exc, ver))
            else:
                raise AssertionError("the reference built %r, expected %s" % (ver, exc))
        for (h, w_, b) in SIZE_CASES:
            out["size/%dx%d/pose" % (h, w_)] = reference_size_case(h, w_, b)
            print("headline at %dx%d, batch %d: from the reference's code" % (h, w_, b), flush=True)
        for B in (1, 4, 5):                                            # 10 samples: 4 pads to 12, 5 divides
            out["cli_loop/B%d" % B] = reference_cli_loop(cli_loop_inputs(), B)
            print("reference CLI loop, batch %d: %d poses for 10 samples" % (B, len(out["cli_loop/B%d" % B])))
    if keys is None:
        np.savez_compressed(os.path.join(HERE, "poses.npz"), **out)
        print("wrote", os.path.join(HERE, "poses.npz"), len(out), "arrays")
    return out


if __name__ == "__main__":
    main(sys.argv[1:] or None)
