"""CPU oracle for the DAVO pose-estimation forward path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``davo_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker
or the CPU baseline -- never as the thing shipped.

PARITY UNPINNED: the reference (BassyKuo/DAVO) is TensorFlow 1.13 graph code
and holds no golden vectors, known-answer tests or checkpoints for this path
(its only test file covers colour maps and cannot import).  TensorFlow 1.13
cannot be installed here (Python 3.12, no wheel, no network).  The oracle is
therefore a line-by-line restatement of the reference files (cited per
function) plus the documented TF 1.13 op semantics; it is cross-checked by a
second, independent plain-C restatement (``oracle/posenn_ref.c``), not by the
reference itself.

What IS pinned against the reference's own code (the TF-free pieces next to the path):
``tests/golden/reference_pins.json`` holds outputs of ``utils/common_utils.py``
(complete_batch_size, is_valid_sample) and ``data/kitti/pose_evaluation_utils.py``
(compute_ate) imported from /root/reference (``tests/golden/make_reference_pins.py``),
and ``oracle/ref_build.py`` compiles the reference's KITTI devkit from its sources into
``oracle/_ref/``, which the tests run on the trajectory file this repo writes.
"""
