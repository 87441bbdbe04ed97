"""On-device trajectory composition and KITTI relative-trajectory-error evaluation.

The reference composes the trajectory on the host, one ``sess.run(pose_vec2mat)`` and one 4x4 product per sample
(reference ``test_kitti_pose.py:136-149``), writes it to a text file and scores it with the KITTI devkit, a C++
program (``kitti_benchmark/cpp/test_odometry_all.cpp:44-125``, driven by ``kitti_benchmark/pose_kitti_eval.sh``).
For sweeps over many sequences or checkpoints both steps run on the GPU here, on the poses where the forward path
left them: ``davo_compose_trajectory`` (batched ``pose_vec2mat`` + blocked parallel prefix product in fp64) and
``davo_kitti_errors`` (the devkit's segment errors, one thread per (first frame, length)).  The host versions in
``geo_utils.py`` remain the default of the CLI (north_star keeps the sequential composition on the host).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi

KITTI_LENGTHS = (100, 200, 300, 400, 500, 600, 700, 800)     # test_odometry_all.cpp:13
KITTI_STEP = 10                                             # :86


def compose_trajectory_gpu(system, poses):
    """``poses``: CUDA float32 tensor ``[N,2,6]`` (or anything ``torch.as_tensor`` takes) -> CUDA float64 tensor
    ``[N+2,4,4]``, the absolute poses ``geo_utils.compose_trajectory`` computes on the host."""
    import torch
    dev = "cuda:%d" % system.device
    poses = torch.as_tensor(poses).to(device=dev, dtype=torch.float32).contiguous()
    n = int(poses.shape[0])
    traj = torch.empty((n + 2, 4, 4), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(system.device).cuda_stream
    system._check(system._lib.davo_compose_trajectory(system._h, C.c_void_p(poses.data_ptr()), n,
                                                      C.c_void_p(traj.data_ptr()), C.c_void_p(stream)),
                  "davo_compose_trajectory")
    return traj


def kitti_errors_gpu(system, gt, result, segments=False):
    """KITTI devkit errors of ``result`` against ``gt`` (both ``[n,4,4]`` or ``[n,3,4]``, numpy or CUDA tensors).

    Returns ``{'t_err': mean translational error per metre, 'r_err': mean rotational error in rad per metre,
    'num': segments}`` -- the two numbers of the devkit's ``<seq>-stats.txt`` (x100 = t_rel in %, x57.3 = deg/m as
    ``show_errors.py:38-39`` prints them) -- plus, with ``segments=True``, the per-segment records as a structured
    array (first_frame, last_frame, r_err, t_err, len, speed; last_frame = -1: sequence too short)."""
    import torch
    dev = "cuda:%d" % system.device

    def full(m):
        m = torch.as_tensor(m).to(device=dev, dtype=torch.float64)
        if m.shape[-2] == 3:
            tail = torch.tensor([0.0, 0.0, 0.0, 1.0], dtype=torch.float64, device=dev).expand(m.shape[0], 1, 4)
            m = torch.cat([m, tail], dim=1)
        return m.contiguous()

    gt, result = full(gt), full(result)
    n = int(gt.shape[0])
    if tuple(result.shape) != (n, 4, 4):
        raise ValueError("kitti_errors_gpu: %d result poses for %d ground-truth poses" % (result.shape[0], n))
    count = -(-n // KITTI_STEP) * len(KITTI_LENGTHS)
    seg = torch.empty((count, 24), dtype=torch.uint8, device=dev) if segments else None
    stats = (C.c_float * 3)()
    stream = torch.cuda.current_stream(system.device).cuda_stream
    system._check(system._lib.davo_kitti_errors(system._h, C.c_void_p(gt.data_ptr()), C.c_void_p(result.data_ptr()), n,
                                                C.c_void_p(seg.data_ptr()) if segments else None, stats, C.c_void_p(stream)),
                  "davo_kitti_errors")
    out = {"t_err": float(stats[0]), "r_err": float(stats[1]), "num": int(stats[2])}
    if segments:
        dt = np.dtype([("first_frame", "<i4"), ("last_frame", "<i4"), ("r_err", "<f4"), ("t_err", "<f4"), ("len", "<f4"), ("speed", "<f4")])
        out["segments"] = seg.cpu().numpy().view(dt).reshape(-1)
    return out
