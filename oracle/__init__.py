"""CPU oracle for the DAVO pose-estimation forward path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``davo_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker
or the CPU baseline -- never as the thing shipped.

HOW IT IS PINNED: the reference (BassyKuo/DAVO) is TensorFlow 1.13 graph code and holds no golden vectors,
known-answer tests or checkpoints for this path, and TensorFlow 1.13 cannot be installed here (Python 3.12, no
wheel, no network).  But its graph code is plain Python: ``tests/golden/make_golden.py`` runs the reference's own,
unmodified ``davo.py`` / ``nets/posenn.py`` / ``nets/attention_module.py`` over ``tests/tf_shim`` (a test-only
stand-in that executes the TensorFlow ops the graph calls on torch-CPU tensors) and writes ``tests/golden/poses.npz``;
``tests/test_oracle.py`` holds this oracle to those fixtures at 1e-9 for every variant (poses, SE class weights,
per-layer statistics, attention maps, feature-mode tensors).  The wiring is therefore the reference's own; what stays
restated is the semantics of the individual TF ops (unit-tested in ``tests/test_tf_shim.py``).  A second, independent
plain-C restatement (``oracle/posenn_ref.c``) cross-checks the arithmetic.

Also pinned against the reference's own code (the TF-free pieces next to the path):
``tests/golden/reference_pins.json`` holds outputs of ``utils/common_utils.py``
(complete_batch_size, is_valid_sample) and ``data/kitti/pose_evaluation_utils.py``
(compute_ate) imported from /root/reference (``tests/golden/make_reference_pins.py``),
and ``oracle/ref_build.py`` compiles the reference's KITTI devkit from its sources into
``oracle/_ref/``, which the tests run on the trajectory file this repo writes.
"""
