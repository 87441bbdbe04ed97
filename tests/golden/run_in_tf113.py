"""Pin the oracle on the reference itself -- for whoever has TensorFlow 1.13.

NOT run in this repo's environment (there is no TensorFlow wheel for its interpreter): this
script exists so that one run of it, anywhere the reference runs, turns ``tests/golden/poses.npz`` from
"the oracle agrees with itself" into "the oracle agrees with the reference".

    # Python 3.6/3.7, tensorflow==1.13.1 (CPU is enough), numpy
    python tests/golden/run_in_tf113.py /path/to/BassyKuo/DAVO  [--out reference_poses.npz] [--case headline ...]

For every case of ``make_golden.CASES`` it builds the reference's OWN inference graph
(``DAVO(version).setup_inference(..., 'davo', ...)``, reference davo.py:1533-1551) on placeholders, assigns the
seeded weights of ``davo_b200.synthetic.init_weights`` to the trainable variables by name (what
``Saver.restore`` does, reference test_kitti_pose.py:129-131), feeds the seeded inputs of
``davo_b200.synthetic.make_inputs`` and fetches ``pred_poses`` (reference davo.py:1553-1569).  It prints the largest
difference to the committed oracle fixture and writes the reference's poses, which can then replace the fixture.
Only numpy-only modules of this repo are imported (synthetic.py, version.py, make_golden.py's CASES table).
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reference", help="checkout of BassyKuo/DAVO")
    ap.add_argument("--out", default="reference_poses.npz")
    ap.add_argument("--case", nargs="*", help="keys of make_golden.CASES (default: all)")
    args = ap.parse_args()
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.abspath(args.reference))
    import tensorflow as tf                                   # 1.13.1
    from davo import DAVO                                     # the REFERENCE's class
    from davo_b200 import synthetic as S                      # numpy only
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_cases", os.path.join(HERE, "make_golden.py"))
    src = open(spec.origin).read()
    ns = {}
    exec(src[src.index("CASES = {"):src.index("def main")], ns)   # the two tables only: the module itself imports torch
    cases, g = ns["CASES"], ns["GOLDEN"]
    gold = np.load(os.path.join(HERE, "poses.npz"))
    B, H, W = g["batch"], g["height"], g["width"]
    img, flow, seg = S.make_inputs(B, H, W, seed=g["input_seed"], bad_label_frac=g["bad_label_frac"])
    depth = S.make_depth(B, H, W)
    out = {}
    for key in (args.case or list(cases)):
        ver = cases[key]
        weights = S.init_weights(ver, seed=g["weight_seed"], random_bias=True)
        tf.reset_default_graph()
        ph_img = tf.placeholder(tf.uint8, [B, H, 3 * W, 3])
        ph_flow = tf.placeholder(tf.float32, [B, 4, H, W, 2])
        ph_seg = tf.placeholder(tf.float32, [B, 3, H, W, 1])
        ph_depth = tf.placeholder(tf.float32, [B, 3, H, W, 1])
        system = DAVO(version=ver)
        system.setup_inference(H, W, "davo", 3, B, ph_img, input_flow=ph_flow, input_depth=ph_depth, input_seglabel=ph_seg)
        variables = {v.name.split(":")[0]: v for v in tf.trainable_variables()}
        missing, extra = sorted(set(variables) - set(weights)), sorted(set(weights) - set(variables))
        if missing or extra:
            print("%-22s variable sets differ: graph-only %s, synthetic-only %s" % (key, missing, extra))
            continue
        with tf.Session(config=tf.ConfigProto(device_count={"GPU": 0})) as sess:
            sess.run(tf.global_variables_initializer())
            for name, var in variables.items():
                var.load(np.asarray(weights[name], np.float32).reshape(var.shape.as_list()), sess)
            pose = sess.run(system.pred_poses, {ph_img: img, ph_flow: flow, ph_seg: seg, ph_depth: depth})
        out[key + "/pose"] = np.asarray(pose, np.float64)
        want = gold[key + "/pose"]
        err = np.abs(out[key + "/pose"] - want)
        ok = np.all(err <= 1e-4 + 1e-3 * np.abs(want))
        print("%-22s max |reference - oracle fixture| = %.3e  (pose scale %.2e)  %s"
              % (key, err.max(), np.abs(want).max(), "within tolerance" if ok else "OUTSIDE the 1e-4 + 1e-3 rel tolerance"))
    np.savez_compressed(args.out, **out)
    print("wrote", args.out, "(%d cases)" % len(out))


if __name__ == "__main__":
    main()
