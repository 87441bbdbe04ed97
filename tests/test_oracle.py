"""CPU tests of the oracle itself (it is the checker, so it gets checked first)."""
import os

import numpy as np
import pytest
import torch

from davo_b200 import synthetic as S
from davo_b200 import version as V
from oracle import c_ref
from oracle import davo_oracle as O
from tests.golden import make_golden as G

HEADLINE = G.CASES["headline"]
GOLD = np.load(os.path.join(os.path.dirname(G.__file__), "poses.npz"))


def test_tf_same_padding_cases():
    # the three strided layers at 128x416 and the dilated ones
    assert O.tf_same_pad(128, 7, 2, 1) == (64, 2, 3)
    assert O.tf_same_pad(416, 7, 2, 1) == (208, 2, 3)
    assert O.tf_same_pad(64, 5, 2, 1) == (32, 1, 2)
    assert O.tf_same_pad(32, 3, 2, 1) == (16, 0, 1)
    assert O.tf_same_pad(104, 3, 2, 1) == (52, 0, 1)
    assert O.tf_same_pad(32, 3, 1, 8) == (32, 8, 8)
    assert O.tf_same_pad(33, 3, 2, 1) == (17, 1, 1)     # odd input: symmetric again


def test_conv_same_matches_manual_loop():
    rng = np.random.default_rng(0)
    x = torch.tensor(rng.normal(size=(1, 7, 9, 3)))
    w = torch.tensor(rng.normal(size=(3, 3, 3, 4)))
    b = torch.tensor(rng.normal(size=(4,)))
    for stride, rate in ((1, 1), (2, 1), (1, 2)):
        y = O.conv2d_same(x, w, b, stride=stride, rate=rate, relu=False).numpy()
        Ho, pt, _ = O.tf_same_pad(7, 3, stride, rate)
        Wo, pl, _ = O.tf_same_pad(9, 3, stride, rate)
        ref = np.zeros((Ho, Wo, 4))
        for oh in range(Ho):
            for ow in range(Wo):
                acc = b.numpy().copy()
                for ty in range(3):
                    for tx in range(3):
                        ih, iw = oh * stride + ty * rate - pt, ow * stride + tx * rate - pl
                        if 0 <= ih < 7 and 0 <= iw < 9:
                            acc += x[0, ih, iw].numpy() @ w[ty, tx].numpy()
                ref[oh, ow] = acc
        np.testing.assert_allclose(y[0], ref, rtol=1e-12, atol=1e-12)


def test_class_gather_out_of_range_and_truncation():
    seg = torch.tensor([[[[0.0], [18.0], [19.0], [255.0]], [[-1.0], [3.9], [-0.5], [7.0]]]])
    w = torch.arange(1, 20, dtype=torch.float64)[None]
    a = O.class_gather(seg, w)[0, ..., 0].numpy()
    # 19, 255, -1 -> 0; 3.9 truncates to 3; -0.5 truncates to 0 (class 0)
    np.testing.assert_array_equal(a, [[1, 19, 0, 0], [0, 4, 1, 8]])


def test_round_tf32_emulation():
    x = torch.tensor([1.0, 1.0 + 2 ** -11, 1.0 + 2 ** -10, -1.0 - 2 ** -11, 3.14159265], dtype=torch.float32)
    r = O._round_tf32(x).numpy()
    assert r[0] == 1.0 and r[1] == np.float32(1.0 + 2 ** -10) and r[2] == np.float32(1.0 + 2 ** -10)
    assert r[3] == np.float32(-1.0 - 2 ** -10)            # ties away from zero
    assert abs(r[4] - 3.14159265) < 2 ** -10


@pytest.mark.parametrize("key", ["headline", "static", "v0_lrelu", "no_segmask"])
def test_c_restatement_agrees_with_torch_restatement(key):
    ver = G.CASES[key]
    w = S.init_weights(ver, random_bias=True)
    img, flow, seg = S.make_inputs(1, 40, 56, seed=7, seg_block=8, bad_label_frac=0.02)
    pc, aw = c_ref.forward(V.parse_version(ver).as_dict(), img, flow, seg, w)
    taps = {}
    pt = O.davo_forward(ver, img, flow, seg, w, torch.float64, taps=taps)
    np.testing.assert_allclose(pc, pt, rtol=1e-10, atol=1e-14)
    if taps["attention_weights"] is not None:
        np.testing.assert_allclose(aw[0, 0], taps["attention_weights"][1][0], rtol=1e-12)
        np.testing.assert_allclose(aw[0, 1], taps["attention_weights"][2][0], rtol=1e-12)


@pytest.mark.parametrize("key", list(G.CASES))
def test_oracle_matches_the_reference_fixture(key):
    """The oracle against tests/golden/poses.npz, which holds what the REFERENCE'S OWN graph code computed for
    the same seeded inputs and weights (tests/golden/make_golden.py runs davo.py / nets/*.py unmodified over
    tests/tf_shim): poses, SE class weights, per-layer statistics of the first PoseNN call, attention maps,
    masked frames, the upsampled cnv6 of mode='feature', and the uint8 colourings byte for byte (CRC)."""
    ver, g = G.CASES[key], G.GOLDEN
    w = S.init_weights(ver, seed=g["weight_seed"], random_bias=True)
    img, flow, seg, depth = G.golden_inputs()
    taps = {}
    pose = O.davo_forward(ver, img, flow, seg, w, torch.float64, taps=taps, depth=depth)
    np.testing.assert_allclose(pose, GOLD[key + "/pose"], rtol=1e-9, atol=1e-13)
    if key + "/att_w" in GOLD.files:
        got = np.stack([a.reshape(a.shape[0], -1) for a in taps["attention_weights"][1:]], 1)
        np.testing.assert_allclose(got, GOLD[key + "/att_w"], rtol=1e-10, atol=1e-14)
    else:
        assert taps["attention_weights"] is None
    # every slim.conv2d output of the first PoseNN call (sample 0): mean, mean |x|, max
    names = {n[len(key) + 6:] for n in GOLD.files if n.startswith(key + "/stat/")}
    t0 = dict(taps["pair0"])
    alias = {"pose.rotation.cnv6": "cnv6_rotation", "pose.translation.cnv6": "cnv6_translation",
             "pose.rotation.cnv7": "cnv7_rotation", "pose.translation.cnv7": "cnv7_translation",
             "pose.cnv6": "cnv6_rotation", "pose.cnv7": "cnv7_rotation"}
    if "skipadd" in key:                   # the oracle taps cnv6 AFTER relu(cnv5 + se_block(cnv6)); the fixture holds the conv's own output
        alias = {k: v for k, v in alias.items() if "cnv6" not in k}
    checked = 0
    for n in sorted(names):
        tap = alias.get(n, n)
        if tap in t0:
            np.testing.assert_allclose(G.stat3(t0[tap][0]), GOLD[key + "/stat/" + n], rtol=1e-9, atol=1e-13, err_msg=n)
            checked += 1
    assert checked >= 6, (checked, sorted(names), sorted(t0))       # input + cnv1..5 at least (se_replace has no cnv6 conv)
    for i in range(3):                                   # attention maps after the target override, every 8th pixel
        a = np.asarray(taps["attention_maps"][i], np.float64)
        np.testing.assert_allclose(a[:, ::G.AMAP_STRIDE, ::G.AMAP_STRIDE, 0], GOLD[key + "/feat/amap%d" % i], rtol=1e-10, atol=1e-14)
        np.testing.assert_allclose(a.mean(), GOLD[key + "/feat/amap%d_mean" % i], rtol=1e-10)


@pytest.mark.parametrize("key", ["headline", "static", "couple_net_v0", "plain_decouple_net", "pix_mix_depthflow", "se_replace"])
def test_oracle_feature_mode_matches_the_reference_fixture(key):
    """O.davo_features against the digests of the reference's inference(mode='feature') (davo.py:1553-1564)."""
    ver, g = G.CASES[key], G.GOLDEN
    w = S.init_weights(ver, seed=g["weight_seed"], random_bias=True)
    img, flow, seg, depth = G.golden_inputs()
    f = O.davo_features(ver, img, flow, seg, w, torch.float64, depth=depth)
    for name, a in G.feature_digest(f).items():
        np.testing.assert_allclose(a, GOLD[key + "/feat/" + name], rtol=1e-9, atol=1e-12, err_msg=name)
    for name, a in G.colour_digest(f).items():
        if name.startswith("flow_px"):
            assert np.abs(a.astype(int) - GOLD[key + "/feat/" + name].astype(int)).max() <= 1, name
        elif name.startswith("flow_crc"):
            continue                     # last-bit atan2 differences flip single bytes; the strided pixels are compared
        else:
            assert a == GOLD[key + "/feat/" + name], name


def test_fp32_oracle_close_to_fp64_oracle():
    w = S.init_weights(HEADLINE, random_bias=True)
    img, flow, seg = S.make_inputs(1, 64, 96, seed=3)
    p64 = O.davo_forward(HEADLINE, img, flow, seg, w, torch.float64)
    p32 = O.davo_forward(HEADLINE, img, flow, seg, w, torch.float32)
    assert np.abs(p32 - p64).max() < 1e-6


def test_target_attention_is_ones_in_se_flow_mode():
    # davo.py:1404-1412: the SE result for the (all-zero) target flow is discarded
    w = S.init_weights(HEADLINE)
    img, flow, seg = S.make_inputs(1, 32, 48, seed=5)
    taps = {}
    O.davo_forward(HEADLINE, img, flow, seg, w, torch.float64, taps=taps)
    assert np.all(taps["attention_maps"][0] == 1.0)
    assert taps["attention_maps"][1].min() > 0 and taps["attention_maps"][1].max() < 1
    # tgt channels of the PoseNN input are the unmasked image; flow slots are zero
    x = taps["pair0"]["input"][0]
    assert np.all(x[..., 3:5] == 0)
    np.testing.assert_allclose(x[..., :3], img[0, :, 48:96].astype(np.float64) / 255 * 2 - 1, atol=1e-12)


def test_unsupported_variants_raise():
    w = S.init_weights(HEADLINE)
    img, flow, seg = S.make_inputs(1, 32, 48)
    with pytest.raises(NameError):
        O.davo_forward("v1-sharedNN-couplePoseNN", img, flow, seg, w)
    with pytest.raises(NameError):
        O.davo_forward("v1-sharedNN", img, flow, seg, w)


# ---- pins taken from the reference's own TF-free code (tests/golden/make_reference_pins.py) ----
def _pins():
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_pins.json")) as f:
        return json.load(f)


def test_batch_padding_and_sample_validity_match_the_reference_functions():
    """parallel.complete_batch_size / is_valid_sample against outputs of the reference's
    utils/common_utils.py:8-28 (imported from /root/reference when the fixture was made)."""
    from davo_b200 import parallel
    pins = _pins()
    for c in pins["complete_batch_size"]:
        out = parallel.complete_batch_size(list(range(c["n"])), c["batch"]) if c["n"] else []
        assert (len(out), out[-8:], int(sum(out))) == (c["len"], c["tail"], c["sum"]), c
    for case in pins["is_valid_sample"].values():
        fr = case["frames"]
        assert [parallel.is_valid_sample(fr, i, 3) for i in range(len(fr))] == case["seq3"]
        assert [parallel.is_valid_sample(fr, i, 5) for i in range(len(fr))] == case["seq5"]


def test_snippet_ate_matches_the_reference_compute_ate():
    """O.snippet_ate against data/kitti/pose_evaluation_utils.py:compute_ate run on TUM files."""
    for c in _pins()["compute_ate"]:
        assert abs(O.snippet_ate(np.array(c["gt"]), np.array(c["pred"])) - c["ate"]) <= 1e-6 * max(1.0, c["ate"])


def test_pose_vec2mat_convention_matches_the_reference_numpy_statement():
    """Host geo_utils.pose_vec2mat and the oracle's against data/kitti/pose_evaluation_utils.py:
    pose_vec2mat(vec, is_kitti_format=False): R = Rx.Ry.Rz of [rz, ry, rx], t = [tx, ty, tz]."""
    from davo_b200 import geo_utils
    for c in _pins()["pose_vec2mat"]:
        v = np.array([c["vec"]], np.float32)
        want = np.array(c["mat"])
        assert np.abs(geo_utils.pose_vec2mat(v)[0] - want).max() < 2e-6
        assert np.abs(O.pose_vec2mat(v)[0] - want).max() < 2e-6


def test_colour_helpers_match_the_reference_sources():
    """mode='feature' colourings: the oracle against the reference's make_color_wheel, Cityscapes colormap,
    label_to_color_image and flow_to_image (utils/flow_utils.py:240-272, 461-593;
    utils/seg_utils/get_dataset_colormap.py:208-234, 383-411), byte for byte."""
    pins = _pins()
    assert np.array_equal(np.array(pins["make_color_wheel"]), O.middlebury_wheel())
    assert np.array_equal(np.array(pins["cityscapes_colormap"]), O.cityscapes_colormap())
    lab = pins["label_to_color_image"]
    assert np.array_equal(O.label_to_color_image(np.array(lab["label"], np.float32)), np.array(lab["image"]))
    for c in pins["flow_to_image_uint8"]:
        assert np.array_equal(O.flow_to_uint8_image(np.array(c["flow"], np.float32)), np.array(c["image"], np.uint8))


def test_spatial_pyramid_pool_reads_the_left_square_of_every_cell():
    """nets/attention_module.py:137-167 with its ksize=[1,h_size,h_size,1] (:158): on a landscape map the window
    is h_size wide at a stride of w_size, zeros of tf.pad included; cells row-major, channel last."""
    H, W = 6, 20
    x = torch.arange(H * W * 2, dtype=torch.float64).reshape(1, H, W, 2)
    got = O.spatial_pyramid_pool(x, (2, 1))
    assert got.shape == (1, (4 + 1) * 2)
    hs, ws = 3, 10                                             # level 2
    want = [x[0, i * hs:(i + 1) * hs, j * ws:j * ws + hs].mean(dim=(0, 1)) for i in range(2) for j in range(2)]
    want.append(x[0, :6, :6].mean(dim=(0, 1)))                 # level 1: a 6 x 6 window out of 6 x 20
    assert torch.allclose(got[0], torch.cat(want))
    # a map that does not divide: H = 7 at level 2 -> h_size 4, one padded zero row counted in the lower cells
    y = torch.ones(1, 7, 20, 1, dtype=torch.float64)
    g = O.spatial_pyramid_pool(y, (2,))[0]
    assert torch.allclose(g, torch.tensor([1.0, 1.0, 0.75, 0.75], dtype=torch.float64))
    with pytest.raises(NotImplementedError):
        O.spatial_pyramid_pool(torch.ones(1, 20, 6, 1), (2,))


def test_resize_bilinear_tf1_rule():
    """TF 1.x resize_bilinear (align_corners=False): source = index * in/out, no half-pixel shift."""
    x = np.arange(2 * 3, dtype=np.float64).reshape(1, 2, 3, 1)
    same = O.resize_bilinear(x, 2, 3)
    assert np.array_equal(same, x)
    up = O.resize_bilinear(x, 4, 6)[0, :, :, 0]
    assert np.allclose(up[0], [0, 0.5, 1, 1.5, 2, 2])          # last column clamps: upper = min(lower+1, in-1)
    assert np.allclose(up[:, 0], [0, 1.5, 3, 3])               # rows likewise
    assert np.allclose(up[1, 1], 0.5 * (0 + 0.5 * 1) + 0.5 * (3 + 0.5 * 1))
    rng = np.random.default_rng(0)
    y = rng.normal(size=(2, 8, 26, 4))
    big = O.resize_bilinear(y, 32, 104)
    assert np.array_equal(big[:, ::4, ::4], y)                 # integer ratio: source samples are kept exactly


def test_feature_dict_has_the_reference_layout():
    """O.davo_features: keys and shapes of DAVO.inference(mode='feature') (davo.py:1553-1564)."""
    from davo_b200 import synthetic as S
    ver = "v1-decay100k-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_flow-abs_flow-fc_tanh"
    img, flow, seg = S.make_inputs(2, 32, 64, seed=3)
    f = O.davo_features(ver, img, flow, seg, S.init_weights(ver))
    assert sorted(f) == ["features", "flows", "images", "masks", "pose", "seg_19", "segs"]
    assert f["pose"].shape == (2, 2, 6) and len(f["flows"]) == 2
    assert [m.shape for m in f["masks"]["attention"]] == [(2, 32, 64, 1)] * 3
    assert np.all(f["masks"]["attention"][0] == 1)             # se_flow: the target map is ones (davo.py:1408-1412)
    assert np.allclose(f["masks"]["image"][0], f["images"][0])
    assert np.allclose(f["masks"]["image"][1], f["images"][1] * f["masks"]["attention"][1])
    assert f["features"]["rot"].shape == (2, 32, 64, 128) and f["seg_19"][0].shape == (2, 32, 64, 19)
    assert np.array_equal(f["seg_19"][1].argmax(-1), np.trunc(seg[:, 0, ..., 0]).astype(int))


def test_reference_pins_are_current_when_the_reference_is_here():
    """In the build container the reference tree exists: the committed pins equal a fresh run."""
    ref = os.environ.get("DAVO_REFERENCE_DIR", "/root/reference")
    if not os.path.isdir(ref):
        pytest.skip("reference tree not on this box")
    import subprocess, sys, json, tempfile, shutil
    here = os.path.dirname(os.path.abspath(__file__))
    with tempfile.TemporaryDirectory() as d:
        shutil.copy(os.path.join(here, "golden", "make_reference_pins.py"), d)
        subprocess.run([sys.executable, os.path.join(d, "make_reference_pins.py")], check=True, capture_output=True)
        with open(os.path.join(d, "reference_pins.json")) as f:
            fresh = json.load(f)
    assert fresh == _pins()


def test_trajectory_file_is_read_by_the_reference_devkit(tmp_path):
    """The reference's own KITTI evaluator (kitti_benchmark/cpp, built by oracle/ref_build.py into
    oracle/_ref/) reads the file geo_utils.write_kitti_trajectory writes: every pose is parsed, a
    result identical to the ground truth scores zero, and a 2 % translation scale error scores 2 %."""
    import subprocess
    from oracle import ref_build
    from davo_b200 import geo_utils
    devkit = ref_build.build()
    if devkit is None:
        pytest.skip("reference devkit not built and reference tree not on this box")
    rng = np.random.default_rng(5)
    gt_dir, res_dir = tmp_path / "data" / "odometry" / "poses", tmp_path / "results" / "x" / "data"
    gt_dir.mkdir(parents=True)
    res_dir.mkdir(parents=True)
    n = 900
    for seq in range(11):
        poses = np.zeros((n, 2, 6), np.float32)
        poses[:, 1, :3] = rng.normal(0, 2e-3, size=(n, 3))          # small rotations
        poses[:, 1, 3:] = [0.0, 0.0, -1.0]                           # tgt -> src1: one metre per frame
        poses[:, 1, 3:] += rng.normal(0, 0.02, size=(n, 3))
        traj = geo_utils.compose_trajectory(poses)
        geo_utils.write_kitti_trajectory(str(gt_dir / ("%02d.txt" % seq)), traj)
        scaled = traj.copy()
        scaled[:, :3, 3] *= 1.02 if seq == 10 else 1.0
        geo_utils.write_kitti_trajectory(str(res_dir / ("%02d.txt" % seq)), scaled)
    out = subprocess.run([devkit, "x"], cwd=str(tmp_path), capture_output=True, text=True, timeout=300).stdout
    for seq in range(11):
        assert "Processing: %02d.txt, poses: %d/%d" % (seq, n + 2, n + 2) in out, out[-2000:]
    assert "Done." in out
    stats = {seq: [float(v) for v in (tmp_path / "results" / "x" / ("%02d-stats.txt" % seq)).read_text().split()]
             for seq in (0, 10)}
    assert stats[0][0] < 1e-5 and stats[0][1] < 1e-7           # identical files: zero translation / rotation error
    assert abs(stats[10][0] - 0.02) < 2e-3                      # 2 % scale error reads as 2 % translation error


def _kitti_like_poses(n, rng, noise=0.0):
    poses = np.zeros((n, 2, 6), np.float32)
    poses[:, 1, :3] = rng.normal(0, 4e-3, size=(n, 3))                  # small rotations
    poses[:, 1, 3:] = np.array([0.0, 0.0, -1.0]) + rng.normal(0, 0.02, size=(n, 3))     # one metre per frame
    poses[:, 0, 3:] = [0.0, 0.0, 1.0]
    if noise:
        poses[:, 1] += rng.normal(0, noise, size=(n, 6)).astype(np.float32) * np.array([0.1, 0.1, 0.1, 1, 1, 1], np.float32)
    return poses


def test_kitti_eval_restatement_matches_the_reference_devkit(tmp_path):
    """oracle/kitti_eval.py (calcSequenceErrors + saveStats restated) against the reference's own C++ devkit run on
    the same files: per-segment records of errors/NN.txt and the two numbers of NN-stats.txt."""
    import subprocess
    from oracle import kitti_eval, ref_build
    from davo_b200 import geo_utils
    devkit = ref_build.build()
    if devkit is None:
        pytest.skip("reference devkit not built and reference tree not on this box")
    rng = np.random.default_rng(9)
    gt_dir, res_dir = tmp_path / "data" / "odometry" / "poses", tmp_path / "results" / "x" / "data"
    gt_dir.mkdir(parents=True)
    res_dir.mkdir(parents=True)
    trajs = {}
    for seq in range(11):
        n = 1100 if seq < 2 else 150                                  # two long sequences (all eight lengths), nine short ones
        base = _kitti_like_poses(n, rng)
        noisy = base.copy()
        noisy[:, 1] += (rng.normal(0, 1, size=(n, 6)) * np.array([2e-4, 2e-4, 2e-4, 0.01, 0.01, 0.01])).astype(np.float32)
        gt, res = geo_utils.compose_trajectory(base), geo_utils.compose_trajectory(noisy)
        geo_utils.write_kitti_trajectory(str(gt_dir / ("%02d.txt" % seq)), gt)
        geo_utils.write_kitti_trajectory(str(res_dir / ("%02d.txt" % seq)), res)
        trajs[seq] = (gt, res)
    out = subprocess.run([devkit, "x"], cwd=str(tmp_path), capture_output=True, text=True, timeout=300).stdout
    assert "Done." in out, out[-2000:]
    for seq in (0, 1, 5):
        errs = kitti_eval.sequence_errors(*trajs[seq])
        rows = [[float(v) for v in line.split()] for line in (tmp_path / "results" / "x" / "errors" / ("%02d.txt" % seq)).read_text().splitlines()]
        assert len(rows) == len(errs), (seq, len(rows), len(errs))
        for row, e in zip(rows, errs):                                # "%d %f %f %f %f": first_frame r_err t_err len speed
            assert int(row[0]) == e[0] and row[3] == e[4]
            assert abs(row[1] - e[2]) <= 1.5e-6 and abs(row[2] - e[3]) <= 1.5e-6 and abs(row[4] - e[5]) <= 1e-4 * e[5] + 1e-6
        if errs:
            t_mean, r_mean = kitti_eval.stats(errs)
            want = [float(v) for v in (tmp_path / "results" / "x" / ("%02d-stats.txt" % seq)).read_text().split()]
            assert abs(want[0] - t_mean) <= 1.5e-6 and abs(want[1] - r_mean) <= 1.5e-6
    assert len(kitti_eval.sequence_errors(*trajs[0])) > 200 and len(kitti_eval.sequence_errors(*trajs[5])) > 0


@pytest.mark.parametrize("size", G.SIZE_CASES[1:])
def test_oracle_matches_the_reference_at_other_frame_sizes(size):
    """The headline variant at other frame sizes (partial widened runs, widths that are not multiples of 16): the oracle
    against the poses the reference's own code computed there (tests/golden/make_golden.reference_size_case)."""
    h, w_, b = size
    w = S.init_weights(HEADLINE, random_bias=True)
    inputs = S.make_inputs(b, h, w_, seed=11, bad_label_frac=0.01)
    np.testing.assert_allclose(O.davo_forward(HEADLINE, *inputs, w, torch.float64), GOLD["size/%dx%d/pose" % (h, w_)], rtol=1e-9, atol=1e-13)
