"""Pins taken from the REFERENCE's own code: the TF-free modules next to the pose path are imported
from /root/reference in this container and their outputs on fixed inputs are committed as
``tests/golden/reference_pins.json`` (the reference tree does not travel to the GPU box).

  utils/common_utils.py: complete_batch_size, is_valid_sample   (sample list of test_kitti_pose.py:83-101,
                                                                 the padding rule of the rank shards)
  data/kitti/pose_evaluation_utils.py: compute_ate              (the snippet ATE of SURVEY 8c)
                                       pose_vec2mat(vec, False)   (numpy statement of the same
                                                                  [rz,ry,rx,tx,ty,tz] -> 4x4 convention as the
                                                                  TF utils/geo_utils.py:93-119 used on the path)

  utils/flow_utils.py: make_color_wheel (pure numpy), and flow_to_image + compute_color -- their source is
                       executed UNMODIFIED over a numpy stand-in for the dozen TensorFlow elementwise ops
                       they call (``_tf_shim``: float32 in, float32 out, one rounding per op), which pins the
                       structure and quirks of the colouring (whole-tensor max radius, col1 := col0);
                       the result goes through convert_image_dtype(uint8) as davo.py:989, 1530-1531 does
  utils/seg_utils/get_dataset_colormap.py: create_cityscapes_label_colormap (pure numpy)

Run:  python tests/golden/make_reference_pins.py
"""
import importlib.util
import json
import os
import sys
import tempfile

import numpy as np

REF = os.environ.get("DAVO_REFERENCE_DIR", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def _load(rel, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _Shape(tuple):
    def as_list(self):
        return list(self)


class _T(np.ndarray):
    """ndarray whose .shape has as_list(), like a TF static shape."""
    @property
    def shape(self):
        return _Shape(np.ndarray.shape.__get__(self))


def _t(x, dtype=None):
    return np.asarray(x, dtype=dtype).view(_T)


def _tf_shim():
    """The TensorFlow calls of flow_to_image / compute_color / label_to_color_image, on numpy."""
    import types
    tf = types.ModuleType("tensorflow")
    tf.float32, tf.int32, tf.uint8 = np.float32, np.int32, np.uint8
    tf.abs = lambda x: _t(np.abs(x))
    tf.sqrt = lambda x: _t(np.sqrt(x))
    tf.floor = lambda x: _t(np.floor(x))
    tf.atan2 = lambda y, x: _t(np.arctan2(y, x))
    tf.zeros_like = lambda x: _t(np.zeros_like(x))
    tf.ones_like = lambda x: _t(np.ones_like(x))
    tf.where = lambda c, a, b: _t(np.where(c, a, b))
    tf.stack = lambda xs, axis=0: _t(np.stack(xs, axis))
    tf.cast = lambda x, dt: _t(np.asarray(x).astype(dt))
    tf.gather = lambda params, idx: _t(np.asarray(params)[np.asarray(idx)])
    tf.reshape = lambda x, shape: _t(np.reshape(x, shape))
    tf.convert_to_tensor = lambda x, name=None: _t(x)
    tf.gather_nd = lambda params, idx: _t(np.asarray(params)[np.asarray(idx)[:, 0]])

    def reduce_max(x):
        if isinstance(x, (list, tuple)):          # TF packs [python int, float32 tensor] into a float32 tensor
            return np.float32(max(np.float32(v) for v in x))
        return np.float32(np.max(x))
    tf.reduce_max = reduce_max
    return tf


def _load_with_shim(rel, name):
    import types
    saved = {k: sys.modules.get(k) for k in ("tensorflow", "png", "matplotlib", "matplotlib.colors",
                                             "matplotlib.pyplot", "PIL", "PIL.Image")}
    sys.modules["tensorflow"] = _tf_shim()
    for k in list(saved)[1:]:
        sys.modules[k] = types.ModuleType(k)
    sys.modules["matplotlib"].colors = sys.modules["matplotlib.colors"]
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["PIL"].Image = sys.modules["PIL.Image"]
    try:
        return _load(rel, name)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def flow_cases():
    rng = np.random.default_rng(5)
    a = rng.normal(0.32, 15.38, size=(2, 6, 8, 2)).astype(np.float32)
    b = rng.normal(0, 0.01, size=(1, 4, 4, 2)).astype(np.float32)
    b[0, 0, 0] = (0.0, 0.0)
    b[0, 1, 1] = (2e7, 1.0)                        # above UNKNOWN_FLOW_THRESH: zeroed
    c = np.zeros((1, 2, 4, 2), np.float32)         # an all-zero flow (the target frame's)
    return [a, b, c]


def batch_cases():
    return [(n, b) for n in (0, 1, 2, 5, 7, 8, 39, 4539) for b in (1, 2, 4, 8, 64)]


def frame_lists():
    a = ["09 %.6d" % i for i in range(7)]
    b = ["09 %.6d" % i for i in range(4)] + ["10 %.6d" % i for i in range(5)]      # drive change inside
    return {"single_drive": a, "two_drives": b}


def ate_cases():
    rng = np.random.default_rng(11)
    out = []
    for n in (2, 3, 5, 40):
        t = np.arange(n) * 0.1
        gt = np.cumsum(rng.normal(0, 1, size=(n, 3)), axis=0)
        pred = gt * rng.uniform(0.5, 2.0) + rng.normal(0, 0.05, size=(n, 3)) + rng.normal(0, 1, size=(1, 3))
        out.append((t, gt, pred))
    return out


def write_tum(path, t, xyz):
    with open(path, "w") as f:
        for ti, p in zip(t, xyz):
            f.write("%.6f %.9f %.9f %.9f 0 0 0 1\n" % (ti, p[0], p[1], p[2]))


def main():
    cu = _load("utils/common_utils.py", "ref_common_utils")
    pe = _load("data/kitti/pose_evaluation_utils.py", "ref_pose_eval")
    pins = {"complete_batch_size": [], "is_valid_sample": {}, "compute_ate": [], "pose_vec2mat": []}
    rng = np.random.default_rng(23)
    vecs = np.concatenate([rng.normal(0, 0.02, size=(6, 6)), rng.uniform(-3.0, 3.0, size=(6, 6)),
                           np.array([[0, 0, 0, 1, 2, 3], [0.5, 0, 0, 0, 0, 0], [0, 0.5, 0, 0, 0, 0], [0, 0, 0.5, 0, 0, 0]], float)])
    for v in vecs:
        pins["pose_vec2mat"].append({"vec": v.tolist(), "mat": np.asarray(pe.pose_vec2mat(v, False), float).tolist()})
    for n, b in batch_cases():
        out = cu.complete_batch_size(list(range(n)), b) if n else []
        pins["complete_batch_size"].append({"n": n, "batch": b, "len": len(out), "tail": out[-8:], "sum": int(sum(out))})
    for key, frames in frame_lists().items():
        pins["is_valid_sample"][key] = {"frames": frames,
                                        "seq3": [bool(cu.is_valid_sample(frames, i, 3)) for i in range(len(frames))],
                                        "seq5": [bool(cu.is_valid_sample(frames, i, 5)) for i in range(len(frames))]}
    with tempfile.TemporaryDirectory() as d:
        for t, gt, pred in ate_cases():
            g, p = os.path.join(d, "g.txt"), os.path.join(d, "p.txt")
            write_tum(g, t, gt)
            write_tum(p, t, pred)
            pins["compute_ate"].append({"t": t.tolist(), "gt": gt.tolist(), "pred": pred.tolist(),
                                        "ate": float(pe.compute_ate(g, p))})
    fu = _load_with_shim("utils/flow_utils.py", "ref_flow_utils")
    cmod = _load_with_shim("utils/seg_utils/get_dataset_colormap.py", "ref_colormap")
    pins["make_color_wheel"] = np.asarray(fu.make_color_wheel(), float).tolist()
    pins["cityscapes_colormap"] = np.asarray(cmod.create_cityscapes_label_colormap(), int).tolist()
    pins["flow_to_image_uint8"] = []
    for fl in flow_cases():
        img01 = np.asarray(fu.flow_to_image(_t(fl)), np.float32)                 # range 0..1 (davo.py:988)
        u8 = np.clip(img01 * np.float32(255.5), 0, 255).astype(np.uint8)         # convert_image_dtype (davo.py:1530)
        pins["flow_to_image_uint8"].append({"flow": fl.tolist(), "image": u8.tolist()})
    lab = np.array([0, 5, 18, 19, 200, 255], np.float32).reshape(1, 2, 3, 1)
    pins["label_to_color_image"] = {"label": lab.tolist(),
                                    "image": np.asarray(cmod.label_to_color_image(_t(lab)), int).tolist()}
    with open(os.path.join(HERE, "reference_pins.json"), "w") as f:
        json.dump(pins, f, indent=1)
    print("wrote", len(pins["complete_batch_size"]), "batch cases,", len(pins["compute_ate"]), "ATE cases")


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference tree not found at " + REF)
    main()
