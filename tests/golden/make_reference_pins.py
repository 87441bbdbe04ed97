"""Pins taken from the REFERENCE's own code: the TF-free modules next to the pose path are imported
from /root/reference in this container and their outputs on fixed inputs are committed as
``tests/golden/reference_pins.json`` (the reference tree does not travel to the GPU box).

  utils/common_utils.py: complete_batch_size, is_valid_sample   (sample list of test_kitti_pose.py:83-101,
                                                                 the padding rule of the rank shards)
  data/kitti/pose_evaluation_utils.py: compute_ate              (the snippet ATE of SURVEY 8c)
                                       pose_vec2mat(vec, False)   (numpy statement of the same
                                                                  [rz,ry,rx,tx,ty,tz] -> 4x4 convention as the
                                                                  TF utils/geo_utils.py:93-119 used on the path)

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
    with open(os.path.join(HERE, "reference_pins.json"), "w") as f:
        json.dump(pins, f, indent=1)
    print("wrote", len(pins["complete_batch_size"]), "batch cases,", len(pins["compute_ate"]), "ATE cases")


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference tree not found at " + REF)
    main()
