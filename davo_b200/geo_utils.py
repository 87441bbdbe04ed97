"""Host-side pose geometry and trajectory composition (numpy).

The reference runs ``pose_vec2mat`` as a tiny fp32 TF graph once per sample
(reference ``utils/geo_utils.py:12-63, 93-119``; call site
``test_kitti_pose.py:122-123, 142``) and chains the relative poses in numpy
fp64 (``test_kitti_pose.py:136-153``).  The sequential composition stays on the
host here as well; only the matrix construction is batched.
"""
from __future__ import annotations

import numpy as np


def euler2mat(z, y, x):
    """[N] angles (radians) -> [N,3,3], R = Rx @ Ry @ Rz, angles clipped to [-pi, pi]
    (reference utils/geo_utils.py:29-31, 41-62).  fp32 like the TF graph."""
    f = np.float32
    z = np.clip(np.asarray(z, f), -np.pi, np.pi).astype(f)
    y = np.clip(np.asarray(y, f), -np.pi, np.pi).astype(f)
    x = np.clip(np.asarray(x, f), -np.pi, np.pi).astype(f)
    n = z.shape[0]
    one, zero = np.ones(n, f), np.zeros(n, f)
    cz, sz, cy, sy, cx, sx = np.cos(z), np.sin(z), np.cos(y), np.sin(y), np.cos(x), np.sin(x)
    zmat = np.stack([cz, -sz, zero, sz, cz, zero, zero, zero, one], 1).reshape(n, 3, 3)
    ymat = np.stack([cy, zero, sy, zero, one, zero, -sy, zero, cy], 1).reshape(n, 3, 3)
    xmat = np.stack([one, zero, zero, zero, cx, -sx, zero, sx, cx], 1).reshape(n, 3, 3)
    return np.matmul(np.matmul(xmat, ymat), zmat).astype(f)


def pose_vec2mat(vec):
    """[N,6] = [rz, ry, rx, tx, ty, tz] -> [N,4,4] (reference utils/geo_utils.py:105-119)."""
    vec = np.asarray(vec, np.float32)
    n = vec.shape[0]
    out = np.zeros((n, 4, 4), np.float32)
    out[:, :3, :3] = euler2mat(vec[:, 0], vec[:, 1], vec[:, 2])
    out[:, :3, 3] = vec[:, 3:6]
    out[:, 3, 3] = 1.0
    return out


def relative_pose_list(pred_poses, batch_size=1, reference_batch_semantics=False):
    """pred_poses [N,2,6] in sample order -> list of 4x4 relative motions.

    Default ("intended" loop): T(tgt->src0) of the first sample only, then inv(T(tgt->src1)) of every sample.
    That is what reference test_kitti_pose.py:136-145 produces at ``--batch_size 1`` (the setting of
    run_inference.sh:50), and it is the trajectory: N samples -> N+2 frames.

    ``reference_batch_semantics=True`` reproduces the reference's loop literally at ``batch_size`` > 1, where it
    differs: ``if i == 0`` (:143) tests the BATCH index inside the ``for j`` loop, so every sample of the first
    batch contributes its tgt->src0 pose in front of its inv(tgt->src1), and ``pred_poses`` is expected to hold
    the duplicated padding samples of ``complete_batch_size`` (:96-101), which the reference composes too.
    The file then has 1 + min(B, N_padded) + N_padded lines.
    """
    pred_poses = np.asarray(pred_poses, np.float32)
    t_src1 = np.linalg.inv(pose_vec2mat(pred_poses[:, 1]))     # one batched LAPACK call: the same bits as one call per matrix
    if reference_batch_semantics and batch_size > 1:
        nb = min(batch_size, len(pred_poses))
        t_src0 = pose_vec2mat(pred_poses[:nb, 0])
        rel = []
        for j in range(nb):                                    # batch i == 0: both poses of every sample, interleaved
            rel.append(t_src0[j])
            rel.append(t_src1[j])
        rel.extend(t_src1[nb:])
        return rel
    rel = [pose_vec2mat(pred_poses[:1, 0])[0]]                 # first sample, tgt->src0
    rel.extend(t_src1)
    return rel


def compose_trajectory(pred_poses, batch_size=1, reference_batch_semantics=False):
    """Absolute poses [len(rel)+1,4,4] fp64, chained by right multiplication
    (reference test_kitti_pose.py:118-119, 147-149)."""
    rel = relative_pose_list(pred_poses, batch_size, reference_batch_semantics)
    out = np.empty((len(rel) + 1, 4, 4), np.float64)
    out[0] = np.eye(4)
    rel64 = np.asarray(rel, np.float64)                # np.dot(float64, float32) promotes the same way
    for i in range(len(rel)):
        np.dot(out[i], rel64[i], out=out[i + 1])
    return out


def write_kitti_trajectory(path, traj):
    """12 ``str(float)`` per line (reference test_kitti_pose.py:150-153)."""
    with open(path, 'w') as f:
        for p in traj:
            f.write(' '.join(str(float(x)) for x in p[:3, :].reshape(12)) + '\n')
