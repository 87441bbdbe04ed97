"""CPU restatement of DAVO's pose-estimation forward graph (torch-CPU, fp64 or fp32).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  PINNED on the reference's own graph code: tests/test_oracle.py
holds every function here to tests/golden/poses.npz, which tests/golden/make_golden.py produces by running the
reference's unmodified davo.py / nets/*.py over tests/tf_shim (TF 1.13 itself is not installable here).

Every function cites the reference file:line it follows (paths relative to the
reference checkout).  Activations are NHWC at the interface (as in TF) and NCHW
inside the conv helper only because ``torch.nn.functional.conv2d`` wants it.

TF 1.13 op semantics encoded here (they live in TensorFlow, not in the tree):
  * ``slim.conv2d``: padding='SAME' (asymmetric: pad_before = total // 2), HWIO
    weights, bias add, then ReLU; ``rate=r`` is a dilated convolution.
  * ``tf.image.convert_image_dtype(u8 -> f32)`` multiplies by 1/255.
  * ``tf.cast(f32 -> i32)`` truncates toward zero; ``tf.one_hot`` of an index
    outside [0, depth) is an all-zero row.
  * ``tf.layers.dense`` on [B,1,1,C] contracts the last axis with kernel [C,units].
  * ``tf.nn.leaky_relu`` default alpha = 0.2.
"""
from __future__ import annotations

import math
import re
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

NUM_CLASSES = 19  # Cityscapes trainIds 0..18 (davo.py:1115, one_hot depth=19)


# --------------------------------------------------------------------------- #
# TF 'SAME' padding (tensorflow/core/framework/common_shape_fns.cc semantics)
# --------------------------------------------------------------------------- #
def tf_same_pad(in_size: int, k: int, stride: int, dil: int) -> Tuple[int, int, int]:
    """Returns (out_size, pad_before, pad_after) of TF padding='SAME'."""
    out = -(-in_size // stride)
    eff = (k - 1) * dil + 1
    total = max((out - 1) * stride + eff - in_size, 0)
    before = total // 2
    return out, before, total - before


def _round_tf32(x: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest (ties away) to TF32's 10-bit mantissa, as cvt.rna.tf32.f32.

    Used only to emulate the tensor-core operand format when a test wants to
    separate indexing errors from TF32 rounding.
    """
    xf = x.to(torch.float32).contiguous()
    bits = xf.view(torch.int32)
    bits = (bits + 0x1000) & ~0x1FFF
    return bits.view(torch.float32).to(x.dtype)


def conv2d_same(x_nhwc: torch.Tensor, w_hwio: torch.Tensor, b: Optional[torch.Tensor],
                stride: int = 1, rate: int = 1, relu: bool = True,
                tf32: bool = False) -> torch.Tensor:
    """``slim.conv2d(x, Cout, [k,k], stride=, rate=)`` (nets/posenn.py:211-215, 238-240).

    tf32=True rounds both operands to TF32 first (products then accumulate in
    the working dtype) -- the arithmetic contract of ``tcgen05.mma.kind::tf32``.
    """
    kh, kw, cin, cout = w_hwio.shape
    _, H, W, C = x_nhwc.shape
    assert C == cin, (C, cin)
    _, pt, pb = tf_same_pad(H, kh, stride, rate)
    _, pl, pr = tf_same_pad(W, kw, stride, rate)
    x = x_nhwc.permute(0, 3, 1, 2)
    w = w_hwio.permute(3, 2, 0, 1)
    if tf32:
        x = _round_tf32(x)
        w = _round_tf32(w)
    x = F.pad(x, (pl, pr, pt, pb))
    y = F.conv2d(x, w.contiguous(), b, stride=stride, padding=0, dilation=rate)
    if relu:
        y = torch.relu(y)
    return y.permute(0, 2, 3, 1).contiguous()


def conv_layer(x_nhwc: torch.Tensor, wts: Dict[str, torch.Tensor], scope: str, stride: int = 1, rate: int = 1,
               relu: bool = True, tf32: bool = False) -> torch.Tensor:
    """One ``slim.conv2d`` of the PoseNN arg_scope (nets/posenn.py:205-209).

    With ``-batch_norm`` the arg_scope sets ``normalizer_fn=slim.batch_norm`` (:206) for every conv except ``pred``
    (:240 passes normalizer_fn=None): slim then creates NO biases but ``<scope>/BatchNorm/beta`` and, because the
    reference passes no normalizer_params, runs batch_norm with its default ``is_training=True`` -- at test time too:
    y = (conv - mean_batch) / sqrt(var_batch + 0.001) + beta with the biased variance over (N, H, W) of THIS call's
    batch, then ReLU.  Which mode a layer is in is read off the variables: beta present <=> batch norm.
    """
    beta = wts.get(scope + "/BatchNorm/beta")
    if beta is None:
        return conv2d_same(x_nhwc, wts[scope + "/weights"], wts[scope + "/biases"], stride=stride, rate=rate, relu=relu, tf32=tf32)
    y = conv2d_same(x_nhwc, wts[scope + "/weights"], None, stride=stride, rate=rate, relu=False, tf32=tf32)
    mean = y.mean(dim=(0, 1, 2))
    var = ((y - mean) ** 2).mean(dim=(0, 1, 2))
    y = (y - mean) * torch.rsqrt(var + 1e-3) + beta
    return torch.relu(y) if relu else y


# --------------------------------------------------------------------------- #
# Attention module
# --------------------------------------------------------------------------- #
def _act(name: str):
    if name == "tanh":
        return torch.tanh
    if name == "lrelu":
        return lambda t: F.leaky_relu(t, 0.2)
    return torch.relu


def spatial_pyramid_pool(x_nhwc: torch.Tensor, out_pool_size) -> torch.Tensor:
    """nets/attention_module.py:137-167 as TensorFlow evaluates it -> [B, sum(n*n) * C].

    Per level n: h_size = ceil(H/n), w_size = ceil(W/n); the map is zero-padded at the bottom / right to
    n*h_size x n*w_size (tf.pad: the zeros count as data); then ``avg_pool`` with
    ``ksize=[1, h_size, h_size, 1]`` -- the reference passes h_size for BOTH window sides (:158) -- and
    strides (h_size, w_size), padding 'SAME'.  For H <= W the window is narrower than the stride, so
    'SAME' adds no padding of its own and output cell (i, j) is the plain mean over rows
    [i*h_size, (i+1)*h_size) and columns [j*w_size, j*w_size + h_size): only the left part of every
    cell is looked at.  Cells are flattened row-major with the channel last (:163-165), levels concatenated.
    """
    B, H, W, C = x_nhwc.shape
    if H > W:
        _unsupported("spatial_pyramid_pool on a portrait map (SAME padding of the square window)")
    out = []
    for n in out_pool_size:
        hs, ws = math.ceil(H / n), math.ceil(W / n)
        xp = torch.nn.functional.pad(x_nhwc, (0, 0, 0, n * ws - W, 0, n * hs - H))
        cells = [xp[:, i * hs:(i + 1) * hs, j * ws:j * ws + hs].mean(dim=(1, 2)) for i in range(n) for j in range(n)]
        out.append(torch.stack(cells, dim=1).reshape(B, n * n * C))
    return torch.cat(out, dim=1)


def se_weights(x_nhwc: torch.Tensor, wts: Dict[str, torch.Tensor], scope: str,
               activation: str, mode: str = "gp", spp_size=(8, 6, 4)) -> torch.Tensor:
    """``se(input, name, layer_channels, mode='gp', activation)`` -> [B, units2].

    nets/attention_module.py:54-103: global average pool (:66), dense +
    activation (:89-94), dense + sigmoid (:96-101); returns the excitation
    vector only.
    """
    if mode == "gp2x2":                                                  # :68-78
        h, w = x_nhwc.shape[1] // 2, x_nhwc.shape[2] // 2
        pool = torch.cat([x_nhwc[:, :h, :w].mean(dim=(1, 2)), x_nhwc[:, :h, w:].mean(dim=(1, 2)),
                          x_nhwc[:, h:, :w].mean(dim=(1, 2)), x_nhwc[:, h:, w:].mean(dim=(1, 2))], dim=-1)
    elif mode == "spp":                                                  # :79-86
        pool = spatial_pyramid_pool(x_nhwc, spp_size)
    else:
        pool = x_nhwc.mean(dim=(1, 2))                                   # :66
    fc1 = _act(activation)(pool @ wts[scope + "/bottleneck_fc/kernel"]
                           + wts[scope + "/bottleneck_fc/bias"])        # :89-94
    return torch.sigmoid(fc1 @ wts[scope + "/recover_fc/kernel"]
                         + wts[scope + "/recover_fc/bias"])             # :96-101


def se_block(x_nhwc: torch.Tensor, wts: Dict[str, torch.Tensor], scope: str,
             activation: str = "relu") -> torch.Tensor:
    """``se_block(input_feature, name, ratio, mode='gp', activation)``.

    nets/attention_module.py:9-52: same two dense layers, returns
    ``input_feature * excitation`` (:51).
    """
    exc = se_weights(x_nhwc, wts, scope, activation)
    return x_nhwc * exc[:, None, None, :]


def class_gather(seg_f32: torch.Tensor, w19: torch.Tensor) -> torch.Tensor:
    """``reduce_sum(one_hot(int32(seg), 19) * w, -1)`` -> [B,H,W,1].

    davo.py:1115 (cast + one_hot + squeeze) and davo.py:1178 (multiply +
    reduce_sum + expand_dims): the per-pixel class weight, 0 for a label
    outside 0..18.  ``w19`` is [B,19] (SE output) or [19] (static weights).
    """
    lab = torch.trunc(seg_f32[..., 0]).to(torch.int64)                   # tf.cast truncates
    valid = (lab >= 0) & (lab < NUM_CLASSES)
    idx = lab.clamp(0, NUM_CLASSES - 1)
    if w19.dim() == 1:
        a = w19[idx]
    else:
        B, H, W = idx.shape
        a = torch.gather(w19, 1, idx.reshape(B, H * W)).reshape(B, H, W)
    a = torch.where(valid, a, torch.zeros_like(a))
    return a[..., None]


# --------------------------------------------------------------------------- #
# PoseNN
# --------------------------------------------------------------------------- #
def decouple_sharednet_v0_dilation(tgt: torch.Tensor, src: torch.Tensor,
                                   wts: Dict[str, torch.Tensor], se_attention=False,
                                   tf32: bool = False,
                                   taps: Optional[Dict[str, torch.Tensor]] = None
                                   ) -> Tuple[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]]:
    """nets/posenn.py:189-254 (dropout=False, batch_norm=False).

    Returns (pose [B,1,6] = [rz,ry,rx,tx,ty,tz], (cnv6_rot, cnv6_trans)).
    ``taps`` (optional dict) receives every intermediate activation.
    """
    P = "pose_exp_net/"

    def cv(x, name, stride=1, rate=1, relu=True):
        return conv_layer(x, wts, P + name, stride=stride, rate=rate, relu=relu, tf32=tf32)

    x = torch.cat([tgt, src], dim=3)                                     # :198
    c1 = cv(x, "cnv1", stride=2)                                         # :211
    c2 = cv(c1, "cnv2", stride=2)                                        # :212
    c3 = cv(c2, "cnv3", rate=2)                                          # :213
    c4 = cv(c3, "cnv4", rate=4)                                          # :214
    c5 = cv(c4, "cnv5", rate=8)                                          # :215
    if taps is not None:
        taps.update(input=x, cnv1=c1, cnv2=c2, cnv3=c3, cnv4=c4, cnv5=c5)
    avgs, c6s = {}, {}
    for name in ("rotation", "translation"):                            # :222
        br = "pose/" + name + "/"
        if se_attention is True:                                         # :225-228
            # NB: cnv5 is re-assigned, so translation's SE sits on top of rotation's.
            c5 = se_block(c5, wts, P + br + "cnv5_se_attention", "relu")
            c6 = cv(c5, br + "cnv6", rate=2)
        elif se_attention == "se_skipadd":                               # :229-233
            c6 = cv(c5, br + "cnv6", rate=2)
            c6 = torch.relu(c5 + se_block(c6, wts, P + br + "cnv6_se_attention", "relu"))
        elif se_attention == "se_replace":                               # :234-236
            c6 = se_block(c5, wts, P + br + "cnv6_se_attention", "relu")
        else:
            c6 = cv(c5, br + "cnv6", rate=2)                             # :238
        c7 = cv(c6, br + "cnv7", stride=2)                               # :239
        pred = cv(c7, br + "pred", relu=False)                           # :240
        avgs[name] = pred.mean(dim=(1, 2))                               # :241
        c6s[name] = c6
        if taps is not None:
            taps["cnv6_" + name] = c6
            taps["cnv7_" + name] = c7
            taps["pred_" + name] = pred
    pose = 0.01 * torch.cat([avgs["rotation"].reshape(-1, 1, 3),
                             avgs["translation"].reshape(-1, 1, 3)], dim=-1)   # :248-250
    return pose, (c6s["rotation"], c6s["translation"])


def couple_sharednet_v0_dilation(tgt: torch.Tensor, src: torch.Tensor,
                                 wts: Dict[str, torch.Tensor], se_attention=False,
                                 tf32: bool = False,
                                 taps: Optional[Dict[str, torch.Tensor]] = None
                                 ) -> Tuple[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]]:
    """nets/posenn.py:133-187 (dropout=False, batch_norm=False): the same shared trunk, ONE pose
    branch under scope ``pose/``: cnv6, cnv7, pred 256 -> 6; pose = 0.01 * mean (:181-184).

    Returns (pose [B,1,6], (cnv6, cnv6)).
    """
    P = "pose_exp_net/"

    def cv(x, name, stride=1, rate=1, relu=True):
        return conv_layer(x, wts, P + name, stride=stride, rate=rate, relu=relu, tf32=tf32)

    x = torch.cat([tgt, src], dim=3)                                     # :142
    c1 = cv(x, "cnv1", stride=2)                                         # :153
    c2 = cv(c1, "cnv2", stride=2)
    c3 = cv(c2, "cnv3", rate=2)
    c4 = cv(c3, "cnv4", rate=4)
    c5 = cv(c4, "cnv5", rate=8)                                          # :157
    if taps is not None:
        taps.update(input=x, cnv1=c1, cnv2=c2, cnv3=c3, cnv4=c4, cnv5=c5)
    if se_attention is True:                                             # :163-166
        c5 = se_block(c5, wts, P + "pose/cnv5_se_attention", "relu")
        c6 = cv(c5, "pose/cnv6", rate=2)
    elif se_attention == "se_skipadd":                                   # :167-171
        c6 = cv(c5, "pose/cnv6", rate=2)
        c6 = torch.relu(c5 + se_block(c6, wts, P + "pose/cnv6_se_attention", "relu"))
    elif se_attention == "se_replace":                                   # :172-174
        c6 = se_block(c5, wts, P + "pose/cnv6_se_attention", "relu")
    else:
        c6 = cv(c5, "pose/cnv6", rate=2)                                 # :176
    c7 = cv(c6, "pose/cnv7", stride=2)                                   # :177
    pred = cv(c7, "pose/pred", relu=False)                               # :178
    if taps is not None:
        taps.update(cnv6_rotation=c6, cnv7_rotation=c7, pred_rotation=pred)
    pose = 0.01 * pred.mean(dim=(1, 2)).reshape(-1, 1, 6)                # :179-181
    return pose, (c6, c6)


def net_v0_dilation(tgt: torch.Tensor, src0: torch.Tensor, src1: torch.Tensor,
                    wts: Dict[str, torch.Tensor], decouple: bool, se_attention=False, tf32: bool = False,
                    taps: Optional[Dict[str, torch.Tensor]] = None, dilated: bool = True):
    """``decouple_net_v0_dilation`` (nets/posenn.py:69-131) / ``couple_net_v0_dilation`` (:12-66),
    dropout=False, batch_norm=False: ONE evaluation per sample on concat(tgt, src0, src1) (:21, :78),
    num_source = 2.  Decouple: rotation / translation branches, pred 256 -> 3*2 each, reshaped
    [-1, 2, 3] and concatenated (:117-123).  Couple: one branch, pred 256 -> 12 -> [-1, 2, 6] (:58-62).

    ``dilated=False``: ``decouple_net_v0`` (:314-378) / ``couple_net_v0`` (:257-311), the same
    graphs with cnv3..cnv6 at stride 2 instead of dilation (:280-282, :300).

    Returns (pose [B,2,6], (cnv6_rot, cnv6_trans)).
    """
    P = "pose_exp_net/"

    def cv(x, name, stride=1, rate=1, relu=True):
        return conv_layer(x, wts, P + name, stride=stride, rate=rate, relu=relu, tf32=tf32)

    def mid(x, name, rate):          # cnv3..cnv6: dilated, or stride 2 in the original nets
        return cv(x, name, rate=rate) if dilated else cv(x, name, stride=2)

    x = torch.cat([tgt, src0, src1], dim=3)
    c1 = cv(x, "cnv1", stride=2)
    c2 = cv(c1, "cnv2", stride=2)
    c3 = mid(c2, "cnv3", 2)
    c4 = mid(c3, "cnv4", 4)
    c5 = mid(c4, "cnv5", 8)
    if taps is not None:
        taps.update(input=x, cnv1=c1, cnv2=c2, cnv3=c3, cnv4=c4, cnv5=c5)
    avgs, c6s = [], []
    for br in (("pose/rotation/", "pose/translation/") if decouple else ("pose/",)):
        if se_attention is True:             # cnv5 is re-assigned: the second branch sees the first's output
            c5 = se_block(c5, wts, P + br + "cnv5_se_attention", "relu")
            c6 = mid(c5, br + "cnv6", 2)     # posenn.py:44, 103 (rate 2) / :289, 351 (stride 2)
        elif se_attention == "se_skipadd":
            c6 = cv(c5, br + "cnv6", rate=2) if dilated else cv(c5, br + "cnv6", stride=1)   # posenn.py:292, 355: stride 1 there
            c6 = torch.relu(c5 + se_block(c6, wts, P + br + "cnv6_se_attention", "relu"))
        elif se_attention == "se_replace":
            c6 = se_block(c5, wts, P + br + "cnv6_se_attention", "relu")
        else:
            c6 = mid(c5, br + "cnv6", 2)
        c7 = cv(c6, br + "cnv7", stride=2)
        pred = cv(c7, br + "pred", relu=False)
        avgs.append(pred.mean(dim=(1, 2)))
        c6s.append(c6)
    if taps is not None:
        taps.update(cnv6_rotation=c6s[0], cnv6_translation=c6s[-1])
    if decouple:
        pose = 0.01 * torch.cat([avgs[0].reshape(-1, 2, 3), avgs[1].reshape(-1, 2, 3)], dim=-1)
    else:
        pose = 0.01 * avgs[0].reshape(-1, 2, 6)
    return pose, (c6s[0], c6s[-1])


# --------------------------------------------------------------------------- #
# Whole inference graph
# --------------------------------------------------------------------------- #
def _unsupported(what):
    raise NotImplementedError("oracle: variant component not restated: " + what)


def davo_forward(version: str, img_u8: np.ndarray, flow: np.ndarray, seg: np.ndarray,
                 weights: Dict[str, np.ndarray], dtype=torch.float64, tf32: bool = False,
                 taps: Optional[dict] = None, depth: Optional[np.ndarray] = None) -> np.ndarray:
    """``DAVO.build_pose_test_graph_davo`` + one ``sess.run`` (davo.py:955-1494, 1553-1569).

    img_u8 [B,H,3W,3] uint8; flow [B,4,H,W,2] f32; seg [B,3,H,W,1] f32.
    Returns pred_poses [B,2,6] as float64 numpy.  Restates the
    ``-sharedNN-dilatedPoseNN`` family with attention sources se_flow /
    static / none; other sources raise NotImplementedError.
    """
    assert version is not None                                           # davo.py:959
    is_read_depth = "depth" in version or "disp" in version            # davo.py:960
    wts = {k: torch.as_tensor(np.asarray(v), dtype=dtype) for k, v in weights.items()}
    B, H, W3, _ = img_u8.shape
    W = W3 // 3
    scale = torch.tensor(1.0 / 255.0, dtype=dtype)                        # convert_image_dtype
    x = torch.as_tensor(img_u8).to(dtype) * scale * 2.0 - 1.0             # davo.py:1519-1522
    # data_loader.py:537-557 -- centre frame is the target
    tgt = x[:, :, W:2 * W, :]
    src0 = x[:, :, 0:W, :]
    src1 = x[:, :, 2 * W:3 * W, :]
    fl = torch.as_tensor(flow).to(dtype)
    sg = torch.as_tensor(seg).to(dtype)
    pred_flows = [torch.zeros_like(fl[:, 0]), fl[:, 0], fl[:, 1]]        # davo.py:978-982
    pred_segs = [sg[:, 1], sg[:, 0], sg[:, 2]]                           # davo.py:1000-1004

    # 0. PoseNN-internal SE (davo.py:1010-1017)
    if "-se_insert" in version:
        se_attention = True
    elif "-se_skipadd" in version:
        se_attention = "se_skipadd"
    elif "-se_replace" in version:
        se_attention = "se_replace"
    else:
        se_attention = False
    input_images = [tgt, src0, src1, tgt]                                # davo.py:1019-1024
    # 1. PoseNN type (davo.py:1027-1049)
    pose_net = decouple_sharednet_v0_dilation
    if "-sharedNN" in version:
        if "-dilatedPoseNN" in version:
            pass
        elif "-dilatedCouplePoseNN" in version:
            pose_net = couple_sharednet_v0_dilation
        elif "-couplePoseNN" in version:
            raise NameError("not support `-sharedNN-couplePoseNN' mode.")
        else:
            raise NameError("unknown PoseNN type.")
    elif "-dilatedPoseNN" in version:                                    # davo.py:1040-1041
        pose_net = "decouple_net_v0_dilation"
    elif "-dilatedCouplePoseNN" in version:                              # davo.py:1042-1043
        pose_net = "couple_net_v0_dilation"
    elif "-couplePoseNN" in version:                                     # davo.py:1044-1045
        pose_net = "couple_net_v0"
    else:                                                                # davo.py:1046-1047
        pose_net = "decouple_net_v0"
    if re.search("-cnv6_([0-9]+)", version) is not None:                 # davo.py:1052-1053
        pass  # width is carried by the weight shapes
    # 2. inputs (davo.py:1057-1073)
    m = re.search("^(v[0-9.]+)", version)
    Version = "v0" if m is None else m.group(1)
    pred_info = None
    if "v0" in Version:
        pass
    elif "v1" in Version:
        pred_info = pred_flows + [pred_flows[0]]
    if "-seglabelid" in version:                                         # davo.py:1066-1073
        # zip(pred_info (4), pred_seglabels (3)) -> 3 entries; davo.py:1442 then reads pred_info[3]
        raise IndexError("list index out of range")
    # 3.1 SE activation (davo.py:1077-1085)
    if "-fc_tanh" in version:
        act = "tanh"
    elif "-fc_lrelu" in version:
        act = "lrelu"
    else:
        act = "relu"
    # 3.2 / 3.3 SE inputs (davo.py:1087-1102)
    se_in = list(pred_flows) + [pred_flows[0]]
    if "-norm_flow" in version:
        se_in = [(f - 0.32140523) / 15.384229 for f in se_in]
    if "-abs_flow_h" in version:
        se_in = [torch.stack([f[..., 0].abs(), f[..., 1]], -1) for f in se_in]
    elif "-abs_flow_v" in version:
        se_in = [torch.stack([f[..., 0], f[..., 1].abs()], -1) for f in se_in]
    elif "-abs_flow" in version:
        se_in = [f.abs() for f in se_in]
    # 3.5 attention maps (davo.py:1114-1400)
    use_se_flow = False
    att_w = None
    if "-se_flow_on_depthseg_seplayers" in version:                      # davo.py:1136-1154
        dp = torch.as_tensor(depth).to(dtype)
        pred_depths = [dp[:, 1], dp[:, 0], dp[:, 2]]
        thres = wts["pose_exp_net/se_flow/depth_threshold"]
        att = []
        for i in range(3):
            near = (pred_depths[i] < thres).to(dtype)                    # :1143
            w_near = se_weights(se_in[i], wts, "pose_exp_net/se_flow_near", act)
            w_far = se_weights(se_in[i], wts, "pose_exp_net/se_flow_far", act)
            att.append(class_gather(pred_segs[i], w_near) * near + class_gather(pred_segs[i], w_far) * (near - 1.0) * (-1.0))
        use_se_flow = True                               # the variables live under pose_exp_net/se_flow* (davo.py:1404)
    elif "-se_flow_on_depthseg" in version:
        _unsupported("depth-split attention")
    elif "-se_mixDepthFlow" in version or "-se_mixDispFlow" in version:  # davo.py:1157-1174
        if not is_read_depth:        # capital D: davo.py:960 is False unless another token holds "depth" / "disp"
            raise UnboundLocalError("local variable 'pred_depths' referenced before assignment")   # davo.py:1161 / 1167
        dp = torch.as_tensor(depth).to(dtype)
        pred_depths = [dp[:, 1], dp[:, 0], dp[:, 2]]
        if "-se_mixDepthFlow" in version:
            terms = [d + pred_depths[0] for d in pred_depths]            # davo.py:1109
            if "-norm_depth" in version:
                terms = [d / 80.0 for d in terms]
            scope = "pose_exp_net/se_depthflow"
        else:
            terms = [1.0 / d for d in pred_depths]                       # davo.py:1167
            scope = "pose_exp_net/se_dispflow"
        att = [se_block(torch.cat([terms[i], se_in[i]], dim=-1), wts, scope, act).sum(-1, keepdim=True) for i in range(3)]
    elif "-se_flow" in version:                                          # davo.py:1175-1180
        att, att_w = [], []
        for i in range(3):
            w19 = se_weights(se_in[i], wts, "pose_exp_net/se_flow", act)
            att_w.append(w19)
            att.append(class_gather(pred_segs[i], w19))
        use_se_flow = True                                               # davo.py:1404
    elif "-se_gp2x2_flow_nobottle" in version or "-se_gp2x2_flow" in version:   # davo.py:1181-1192
        att, att_w = [], []
        for i in range(3):
            w19 = se_weights(se_in[i], wts, "pose_exp_net/se_flow", act, mode="gp2x2")
            att_w.append(w19)
            att.append(class_gather(pred_segs[i], w19))
        use_se_flow = True                               # variables under pose_exp_net/se_flow (davo.py:1404)
    elif re.search("-se_spp(21|2|864|)_flow", version):                  # davo.py:1193-1210, first match wins
        sizes = (2, 1) if "-se_spp21_flow" in version else (2,) if "-se_spp2_flow" in version else (8, 6, 4)
        att, att_w = [], []
        for i in range(3):
            w19 = se_weights(se_in[i], wts, "pose_exp_net/se_flow", act, mode="spp", spp_size=sizes)
            att_w.append(w19)
            att.append(class_gather(pred_segs[i], w19))
        use_se_flow = True                               # variables under pose_exp_net/se_flow (davo.py:1404)

    elif "-se_depth_wo_tgt_to_seg" in version or "-se_depth_to_seg" in version:     # davo.py:1211-1227
        dp = torch.as_tensor(depth).to(dtype)
        pred_depths = [dp[:, 1], dp[:, 0], dp[:, 2]]                     # davo.py:991-996: tgt, src0, src1
        # davo.py:1109: `[d for d in pred_depths] + pred_depths[0]` is list + Tensor: TensorFlow
        # converts the list to a [3,B,H,W,1] tensor and BROADCASTS the add, so every SE input is
        # depth_i + depth_tgt.
        se_in_d = [d + pred_depths[0] for d in pred_depths]
        if "-norm_depth" in version:                                     # davo.py:1110-1111
            se_in_d = [d / 80.0 for d in se_in_d]
        att, att_w = [], []
        for i in range(3):
            w19 = se_weights(se_in_d[i], wts, "pose_exp_net/se_depth", act)
            att_w.append(w19)
            att.append(class_gather(pred_segs[i], w19))
        if "-se_depth_wo_tgt_to_seg" in version:
            att[0] = torch.ones_like(att[0])                             # davo.py:1218
    elif "-se_depth_wo_tgt" in version or "-se_depth" in version:        # davo.py:1228-1245
        dp = torch.as_tensor(depth).to(dtype)
        pred_depths = [dp[:, 1], dp[:, 0], dp[:, 2]]
        se_in_d = [d + pred_depths[0] for d in pred_depths]              # davo.py:1109 (list + Tensor broadcast)
        if "-norm_depth" in version:
            se_in_d = [d / 80.0 for d in se_in_d]
        att = [se_block(x, wts, "pose_exp_net/se_depth", act).sum(-1, keepdim=True) for x in se_in_d]
        if "-se_depth_wo_tgt" in version:
            att[0] = torch.ones_like(att[0])                             # davo.py:1236
    elif "-se_disp_wo_tgt_to_seg" in version or "-se_disp_to_seg" in version:   # davo.py:1246-1270
        dp = torch.as_tensor(depth).to(dtype)
        pred_depths = [dp[:, 1], dp[:, 0], dp[:, 2]]                     # davo.py:991-996: tgt, src0, src1
        att, att_w = [], []
        for i in range(3):
            w19 = se_weights(1.0 / pred_depths[i], wts, "pose_exp_net/se_disp", act)
            att_w.append(w19)
            att.append(class_gather(pred_segs[i], w19))
        if "-se_disp_wo_tgt_to_seg" in version:
            att[0] = torch.ones_like(att[0])                             # davo.py:1261
    elif "-se_disp_wo_tgt" in version or "-se_disp" in version:          # davo.py:1271-1292
        dp = torch.as_tensor(depth).to(dtype)
        pred_depths = [dp[:, 1], dp[:, 0], dp[:, 2]]
        att = [se_block(1.0 / d, wts, "pose_exp_net/se_disp", act).sum(-1, keepdim=True) for d in pred_depths]
        if "-se_disp_wo_tgt" in version:
            att[0] = torch.ones_like(att[0])                             # davo.py:1280
    elif "-se_rgb_wo_tgt_to_seg" in version or "-se_rgb_to_seg" in version:   # davo.py:1274-1292
        att, att_w = [], []
        for i in range(3):
            w19 = se_weights(input_images[i], wts, "pose_exp_net/se_rgb", act)
            att_w.append(w19)
            att.append(class_gather(pred_segs[i], w19))
        if "-se_rgb_wo_tgt_to_seg" in version:
            att[0] = torch.ones_like(att[0])                             # davo.py:1283
    elif "-se_rgb_wo_tgt" in version or "-se_rgb" in version:            # davo.py:1293-1303
        att = [se_block(input_images[i], wts, "pose_exp_net/se_rgb", act).sum(-1, keepdim=True) for i in range(3)]
        if "-se_rgb_wo_tgt" in version:
            att[0] = torch.ones_like(att[0])                             # davo.py:1297
    elif ("-se_seg_wo_tgt" in version or "-se_seg" in version or "-se_gp2x2_seg" in version
          or re.search("-se_spp(21|2|864|)_seg", version) or "-se_spp_seg_21" in version):   # davo.py:1304-1340
        # se_block / se_spp_block on the one-hot label map, ratio=1: pooled class frequencies -> 19 -> 19.
        # Reference order: -se_seg_wo_tgt, -se_seg, -se_gp2x2_seg, -se_spp21_seg | -se_spp_seg_21, -se_spp2_seg,
        # -se_spp_seg | -se_spp864_seg
        if "-se_seg_wo_tgt" in version or "-se_seg" in version:
            seg_scope, seg_mode, seg_sizes = "pose_exp_net/se_seg", "gp", None
        elif "-se_gp2x2_seg" in version:
            seg_scope, seg_mode, seg_sizes = "pose_exp_net/se_seg", "gp2x2", None
        else:
            seg_scope, seg_mode = "pose_exp_net/se_spp_seg", "spp"
            seg_sizes = ((2, 1) if ("-se_spp21_seg" in version or "-se_spp_seg_21" in version)
                         else (2,) if "-se_spp2_seg" in version else (8, 6, 4))
        att, att_w = [], []
        for i in range(3):
            lab = torch.trunc(pred_segs[i][..., 0]).to(torch.int64)      # davo.py:1115
            onehot = torch.nn.functional.one_hot(lab.clamp(0, NUM_CLASSES - 1), NUM_CLASSES).to(dtype)
            onehot = onehot * ((lab >= 0) & (lab < NUM_CLASSES)).to(dtype)[..., None]
            exc = se_weights(onehot, wts, seg_scope, act, mode=seg_mode, spp_size=seg_sizes)   # 19 per cell -> 19 -> 19
            att_w.append(exc)
            att.append((onehot * exc[:, None, None, :]).sum(-1, keepdim=True))   # attention_module.py:51, davo.py:1306
        if "-se_seg_wo_tgt" in version:
            att[0] = torch.ones_like(att[0])                             # davo.py:1310
    elif "-se_SegFlow_to_seg" in version:                                # davo.py:1341-1374 (four tokens)
        att, att_w = [], []
        for i in range(3):
            lab = torch.trunc(pred_segs[i][..., 0]).to(torch.int64)      # davo.py:1115
            onehot = torch.nn.functional.one_hot(lab.clamp(0, NUM_CLASSES - 1), NUM_CLASSES).to(dtype)
            onehot = onehot * ((lab >= 0) & (lab < NUM_CLASSES)).to(dtype)[..., None]
            se_input = torch.cat([onehot, se_in[i]], dim=-1)             # 19 + 2 channels
            w19 = se_weights(se_input, wts, "pose_exp_net/se_segflow", act)   # [8|19, 19] by the weight shapes
            att_w.append(w19)
            att.append((onehot * w19[:, None, None, :]).sum(-1, keepdim=True))
        if "-se_SegFlow_to_seg_8_wo_tgt" in version or (                 # first match wins: davo.py:1341, 1350, 1358
                "-se_SegFlow_to_seg_8" not in version and "-se_SegFlow_to_seg_wo_tgt" in version):
            att[0] = torch.ones_like(att[0])
    elif "-se_mixSegFlow" in version or "-se_spp21_mixSegFlow" in version:   # davo.py:1375-1383
        spp21 = "-se_mixSegFlow" not in version          # first match wins (the two tokens do not contain each other)
        att = []
        for i in range(3):
            lab = torch.trunc(pred_segs[i][..., 0]).to(torch.int64)
            onehot = torch.nn.functional.one_hot(lab.clamp(0, NUM_CLASSES - 1), NUM_CLASSES).to(dtype)
            onehot = onehot * ((lab >= 0) & (lab < NUM_CLASSES)).to(dtype)[..., None]
            x = torch.cat([onehot, se_in[i]], dim=-1)                    # 19 + 2 channels
            if spp21:                                                    # se_spp_block(..., "se_spp_segflow", ratio=1, spp_size=[2,1])
                exc = se_weights(x, wts, "pose_exp_net/se_spp_segflow", act, mode="spp", spp_size=(2, 1))
                att.append((x * exc[:, None, None, :]).sum(-1, keepdim=True))
            else:
                att.append(se_block(x, wts, "pose_exp_net/se_segflow", act).sum(-1, keepdim=True))
    elif "-no_segmask" in version:                                       # davo.py:1385-1389
        att = [torch.ones_like(s) for s in pred_segs]
    elif "-segmask_" in version and "-static" in version:                # davo.py:1390-1394
        w19 = torch.sigmoid(wts["pose_exp_net/pose_exp_net/seg_channel_weight/weight"])
        att = [class_gather(s, w19) for s in pred_segs]                  # posenn.py:380-394
        att[0] = torch.ones_like(att[0])
    else:                                                                # davo.py:1395-1399
        w19 = torch.sigmoid(wts["pose_exp_net/pose_exp_net/seg_channel_weight/weight"])
        att = [class_gather(s, w19) for s in pred_segs]
    # 4.1 masking (davo.py:1404-1450)
    a_tgt, a_s0, a_s1 = att
    if use_se_flow:
        a_tgt = torch.ones_like(a_tgt)
        a_ts1 = torch.ones_like(a_tgt)
    else:
        a_ts1 = a_tgt
    amaps = [a_tgt, a_s0, a_s1, a_ts1]
    if pred_info is not None:
        pred_info = list(pred_info)
        if "-segmask_" in version:
            input_images = [im * a for im, a in zip(input_images, amaps)]
            if "-segmask_all" in version and ".555" in Version:
                pred_info[0] = pred_info[0] * a_tgt
                pred_info[3] = pred_info[3] * a_ts1
            elif "-segmask_all" in version:
                pred_info = [pi * a for pi, a in zip(pred_info, amaps)]
            elif "-segmask_rgb" in version:
                pass
        input_images = [torch.cat([im, pi], 3) for im, pi in zip(input_images, pred_info)]
    else:
        if "-segmask" in version:
            input_images = [im * a for im, a in zip(input_images, amaps)]
    # 4.2 PoseNN x2 with shared weights (davo.py:1453-1458)
    t0 = {} if taps is not None else None
    t1 = {} if taps is not None else None
    if isinstance(pose_net, str):                                        # davo.py:1459-1460: one evaluation per sample
        pred_poses, _ = net_v0_dilation(input_images[0], input_images[1], input_images[2], wts,
                                        pose_net.startswith("decouple"), se_attention, tf32, t0,
                                        dilated=pose_net.endswith("_dilation"))
        t1 = t0
    else:
        pose0, _ = pose_net(input_images[0], input_images[1], wts, se_attention, tf32, t0)
        pose1, _ = pose_net(input_images[3], input_images[2], wts, se_attention, tf32, t1)
        pred_poses = torch.cat([pose0, pose1], dim=-2)                   # davo.py:1458
    if taps is not None:
        taps["images"] = [t.numpy() for t in (tgt, src0, src1)]                          # davo.py:967-971
        taps["masked_images"] = [im[..., :3].numpy() for im in input_images[:3]]        # davo.py:1470-1474
        taps["attention_maps"] = [a.numpy() for a in (a_tgt, a_s0, a_s1)]
        taps["attention_weights"] = None if att_w is None else [w.numpy() for w in att_w]
        taps["pair0"] = {k: v.numpy() for k, v in t0.items()}
        taps["pair1"] = {k: v.numpy() for k, v in t1.items()}
    return pred_poses.to(torch.float64).numpy()


# --------------------------------------------------------------------------- #
# mode='feature' (davo.py:1463-1494, 1553-1564)
# --------------------------------------------------------------------------- #
def resize_bilinear(x_nhwc: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """``tf.image.resize_bilinear(x, [out_h, out_w])`` of TF 1.13 (align_corners=False, no half-pixel
    centres): source coordinate = index * (in / out) in float32, lower = trunc, upper = min(lower + 1,
    in - 1), columns blended first, then rows (davo.py:1463-1464)."""
    x = np.asarray(x_nhwc)
    B, h, w, C = x.shape
    fy = np.arange(out_h, dtype=np.float32) * np.float32(h / np.float32(out_h))
    fx = np.arange(out_w, dtype=np.float32) * np.float32(w / np.float32(out_w))
    y0, x0 = fy.astype(np.int64), fx.astype(np.int64)
    y1, x1 = np.minimum(y0 + 1, h - 1), np.minimum(x0 + 1, w - 1)
    ly = (fy - y0).astype(x.dtype)[None, :, None, None]
    lx = (fx - x0).astype(x.dtype)[None, None, :, None]
    top = x[:, y0][:, :, x0] + (x[:, y0][:, :, x1] - x[:, y0][:, :, x0]) * lx
    bot = x[:, y1][:, :, x0] + (x[:, y1][:, :, x1] - x[:, y1][:, :, x0]) * lx
    return top + (bot - top) * ly


def cityscapes_colormap() -> np.ndarray:
    """utils/seg_utils/get_dataset_colormap.py:208-234: 256 rows, the 19 train ids coloured, the rest black."""
    cm = np.zeros((256, 3), np.uint8)
    cm[:19] = [(128, 64, 128), (244, 35, 232), (70, 70, 70), (102, 102, 156), (190, 153, 153), (153, 153, 153),
               (250, 170, 30), (220, 220, 0), (107, 142, 35), (152, 251, 152), (70, 130, 180), (220, 20, 60),
               (255, 0, 0), (0, 0, 142), (0, 0, 70), (0, 60, 100), (0, 80, 100), (0, 0, 230), (119, 11, 32)]
    return cm


def label_to_color_image(seg_f32: np.ndarray) -> np.ndarray:
    """get_dataset_colormap.py:383-411: gather_nd(colormap, int32(label)); an index outside the table
    gives zeros (the GPU gather_nd rule; the table has 256 rows)."""
    lab = np.trunc(np.asarray(seg_f32)[..., 0]).astype(np.int64)
    ok = (lab >= 0) & (lab < 256)
    return cityscapes_colormap()[np.where(ok, lab, 255)] * ok[..., None].astype(np.uint8)


def middlebury_wheel() -> np.ndarray:
    """utils/flow_utils.py:546-593 (float64 [55,3])."""
    wheel, col = np.zeros((55, 3)), 0
    for n, fixed, ramp, rising in ((15, 0, 1, True), (6, 1, 0, False), (4, 1, 2, True),
                                   (11, 2, 1, False), (13, 2, 0, True), (6, 0, 2, False)):
        steps = np.floor(255 * np.arange(n) / n)
        wheel[col:col + n, fixed] = 255
        wheel[col:col + n, ramp] = steps if rising else 255 - steps
        col += n
    return wheel


def flow_to_uint8_image(flow_f32: np.ndarray) -> np.ndarray:
    """``convert_to_tf_image(flow_to_image(flow))`` (davo.py:988-989, 1530-1531; utils/flow_utils.py:240-272,
    461-500), float32 step by step as the TF ops are.  Kept quirks: the max radius is taken over the whole
    [B,h,w] tensor; ``col1`` is assigned from ``col0`` (:491), so the wheel is not interpolated; ``k1`` is unused."""
    f32 = np.float32
    fl = np.asarray(flow_f32, f32)
    u, v = fl[..., 0].copy(), fl[..., 1].copy()
    u[np.abs(u) > 1e7] = 0
    v[np.abs(v) > 1e7] = 0
    rad = np.sqrt(u * u + v * v)
    maxrad = max(f32(-1), rad.max())
    u = u / (maxrad + f32(1e-5))
    v = v / (maxrad + f32(1e-5))
    rad = np.sqrt(u * u + v * v)
    a = np.arctan2(-v, -u) / f32(np.pi)
    fk = (a + f32(1)) / f32(2) * f32(54) + f32(1)
    k0 = np.floor(fk)
    f = fk - k0
    wheel = middlebury_wheel()
    idx = np.clip(k0.astype(np.int64) - 1, 0, 54)
    out = np.empty(u.shape + (3,), np.uint8)
    for ch in range(3):
        col0 = (wheel[:, ch][idx] / 255.).astype(f32)
        col = (f32(1) - f) * col0 + f * col0
        col = np.where(rad <= 1, f32(1) - rad * (f32(1) - col), col * f32(0.75))
        img = np.floor(f32(255.) * col) / f32(255.)
        out[..., ch] = np.clip(img * f32(255.5), 0, 255).astype(np.uint8)     # convert_image_dtype(uint8)
    return out


def davo_features(version: str, img_u8: np.ndarray, flow: np.ndarray, seg: np.ndarray,
                  weights: Dict[str, np.ndarray], dtype=torch.float64,
                  depth: Optional[np.ndarray] = None) -> dict:
    """``DAVO.inference(sess, mode='feature')`` (davo.py:1553-1564): the fetched dict."""
    taps: dict = {}
    pose = davo_forward(version, img_u8, flow, seg, weights, dtype, taps=taps, depth=depth)
    H, W = img_u8.shape[1], img_u8.shape[2] // 3
    last = taps["pair1"]                                  # the call whose cnv6 is kept (davo.py:1456-1460)
    rot = last["cnv6_rotation"]
    trans = last.get("cnv6_translation", rot)            # couple nets return (cnv6, cnv6)
    sg = np.asarray(seg, np.float32)
    segs = [sg[:, 1], sg[:, 0], sg[:, 2]]                 # davo.py:1000-1004
    seg_19 = []
    for s_ in segs:                                       # davo.py:1115
        lab = np.trunc(s_[..., 0]).astype(np.int64)
        oh = (lab[..., None] == np.arange(NUM_CLASSES)).astype(np.float32)
        seg_19.append(np.squeeze(oh))
    return {
        "pose": pose,
        "masks": {"image": taps["masked_images"], "attention": taps["attention_maps"]},
        "features": {"rot": resize_bilinear(rot, H, W), "trans": resize_bilinear(trans, H, W)},
        "images": taps["images"],
        "flows": [flow_to_uint8_image(np.asarray(flow)[:, k]) for k in range(2)],      # davo.py:989: [1:]
        "segs": [label_to_color_image(s_) for s_ in segs],
        "seg_19": seg_19,
    }


# --------------------------------------------------------------------------- #
# Host trajectory composition
# --------------------------------------------------------------------------- #
def euler2mat(z, y, x, dtype=np.float32):
    """utils/geo_utils.py:12-63: clip to [-pi, pi] (:29-31), R = Rx @ Ry @ Rz (:62)."""
    z = np.clip(np.asarray(z, dtype), -np.pi, np.pi).astype(dtype)
    y = np.clip(np.asarray(y, dtype), -np.pi, np.pi).astype(dtype)
    x = np.clip(np.asarray(x, dtype), -np.pi, np.pi).astype(dtype)
    n = z.shape[0]
    zmat = np.zeros((n, 3, 3), dtype)
    ymat = np.zeros((n, 3, 3), dtype)
    xmat = np.zeros((n, 3, 3), dtype)
    cz, sz = np.cos(z), np.sin(z)
    zmat[:, 0, 0], zmat[:, 0, 1] = cz, -sz                               # :41-46
    zmat[:, 1, 0], zmat[:, 1, 1] = sz, cz
    zmat[:, 2, 2] = 1
    cy, sy = np.cos(y), np.sin(y)
    ymat[:, 0, 0], ymat[:, 0, 2] = cy, sy                                # :48-53
    ymat[:, 1, 1] = 1
    ymat[:, 2, 0], ymat[:, 2, 2] = -sy, cy
    cx, sx = np.cos(x), np.sin(x)
    xmat[:, 0, 0] = 1                                                    # :55-60
    xmat[:, 1, 1], xmat[:, 1, 2] = cx, -sx
    xmat[:, 2, 1], xmat[:, 2, 2] = sx, cx
    return (xmat @ ymat @ zmat).astype(dtype)                            # :62


def pose_vec2mat(vec, dtype=np.float32):
    """utils/geo_utils.py:93-119: [rz,ry,rx,tx,ty,tz] -> 4x4 (fp32 in the TF graph)."""
    vec = np.asarray(vec, dtype)
    n = vec.shape[0]
    out = np.zeros((n, 4, 4), dtype)
    out[:, :3, :3] = euler2mat(vec[:, 0], vec[:, 1], vec[:, 2], dtype)
    out[:, :3, 3] = vec[:, 3:6]
    out[:, 3, 3] = 1
    return out


def compose_trajectory(pred_poses: np.ndarray, batch_size: int = 1,
                       reference_batch_semantics: bool = False) -> np.ndarray:
    """test_kitti_pose.py:133-149 for seq_length 3.

    Default: pred_poses [N,2,6] in sample order -> absolute poses [N+2,4,4] (fp64): identity, then
    T(tgt->src0) of the first sample, then inv(T(tgt->src1)) of every sample, chained by
    right-multiplication -- the reference's output at --batch_size 1.

    ``reference_batch_semantics``: the loop as written, for any batch size: ``i`` is the batch index, so every
    sample ``j`` of batch 0 appends its tgt->src0 pose (:143-144); pred_poses must already hold the padding
    duplicates of complete_batch_size (:96-101), which the loop composes like any other sample.
    """
    pred_poses = np.asarray(pred_poses, np.float32)
    B = batch_size if reference_batch_semantics else 1
    rel = []
    for s in range(pred_poses.shape[0]):
        i = s // B                                                       # :133 `for i in range(round_num)`, :136 `for j`
        v = np.insert(pred_poses[s], 1, np.zeros((1, 6), np.float32), axis=0)   # :141
        m = pose_vec2mat(v)                                              # :142
        if (i == 0) if reference_batch_semantics else (s == 0):
            rel.append(m[0])                                             # :143-144
        rel.append(np.linalg.inv(m[2]))                                  # :145
    prev = np.eye(4).astype(float)
    out = [prev]
    for p in rel:                                                        # :147-149
        prev = np.dot(prev, p)
        out.append(prev)
    return np.stack(out)


def kitti_lines(traj: np.ndarray) -> List[str]:
    """test_kitti_pose.py:150-153: 12 ``str(float)`` per line."""
    return [" ".join(str(float(v)) for v in p[:3, :].reshape(12)) for p in traj]


def snippet_ate(gtruth_xyz: np.ndarray, pred_xyz: np.ndarray) -> float:
    """``compute_ate`` of data/kitti/pose_evaluation_utils.py:7-27 on already associated positions:
    align the first frames (:20-21), fit one scale factor (:24), RMSE divided by the number of
    matches -- the reference divides sqrt(sum) by N, not by sqrt(N) (:26).  Pinned against the
    reference's own function in tests/golden/reference_pins.json."""
    g = np.asarray(gtruth_xyz, np.float64)
    p = np.asarray(pred_xyz, np.float64).copy()
    p += (g[0] - p[0])[None, :]
    scale = np.sum(g * p) / np.sum(p ** 2)
    err = p * scale - g
    return float(np.sqrt(np.sum(err ** 2)) / g.shape[0])


def ate(traj_a: np.ndarray, traj_b: np.ndarray) -> float:
    """RMSE of translation differences, same origin, no alignment."""
    d = traj_a[:, :3, 3] - traj_b[:, :3, 3]
    return float(math.sqrt(np.mean(np.sum(d * d, axis=1))))
