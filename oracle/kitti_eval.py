"""CPU restatement (numpy) of the reference's KITTI devkit segment errors -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows kitti_benchmark/cpp/test_odometry_all.cpp of the reference: ``trajectoryDistances`` :44-56 (float accumulation
in frame order), ``lastFrameFromSegmentLength`` :58-63, ``rotationError`` :65-71, ``translationError`` :73-78,
``calcSequenceErrors`` :80-125 (lengths :13, step :86), ``saveStats`` :395-404.  Pinned against the devkit itself:
tests/test_oracle.py runs the reference's C++ program (compiled from its own sources into oracle/_ref/ by
oracle/ref_build.py) on the same trajectories and compares the stats files.
"""
import numpy as np

LENGTHS = np.array([100, 200, 300, 400, 500, 600, 700, 800], np.float32)      # :13
STEP = 10                                                                       # :86


def _full(m):
    m = np.asarray(m, np.float64)
    if m.shape[-2] == 3:
        m = np.concatenate([m, np.broadcast_to(np.array([[0.0, 0.0, 0.0, 1.0]]), (m.shape[0], 1, 4))], axis=1)
    return m


def trajectory_distances(gt):
    d = np.zeros(len(gt), np.float32)
    for i in range(1, len(gt)):                                                 # :47-55, all in float
        dx, dy, dz = (np.float32(gt[i - 1][k, 3] - gt[i][k, 3]) for k in range(3))
        d[i] = d[i - 1] + np.sqrt(dx * dx + dy * dy + dz * dz, dtype=np.float32)
    return d


def sequence_errors(gt, res):
    """-> list of (first_frame, last_frame, r_err / len, t_err / len, len, speed), float32 values as the devkit's."""
    gt, res = _full(gt), _full(res)
    dist = trajectory_distances(gt)
    out = []
    for ff in range(0, len(gt), STEP):
        for ln in LENGTHS:
            hit = np.nonzero(dist[ff:] > dist[ff] + ln)[0]                       # :58-63
            if len(hit) == 0:
                continue
            lf = ff + int(hit[0])
            dg = np.linalg.inv(gt[ff]) @ gt[lf]                                 # :104
            dr = np.linalg.inv(res[ff]) @ res[lf]                               # :105
            e = np.linalg.inv(dr) @ dg                                          # :106
            a, b, c = np.float32(e[0, 0]), np.float32(e[1, 1]), np.float32(e[2, 2])
            d = np.float32(0.5 * (np.float64(a + b + c) - 1.0))                 # :65-70
            r_err = np.arccos(np.maximum(np.minimum(d, np.float32(1)), np.float32(-1)), dtype=np.float32)
            t = e[:3, 3].astype(np.float32)
            t_err = np.sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2], dtype=np.float32)   # :73-78
            speed = np.float32(np.float64(ln) / (0.1 * np.float64(np.float32(lf - ff + 1))))   # :113-114
            out.append((ff, lf, np.float32(r_err / ln), np.float32(t_err / ln), ln, speed))
    return out


def stats(errs):
    """saveStats :395-404: (mean t_err, mean r_err), float accumulation in order."""
    t = r = np.float32(0)
    for e in errs:
        t = np.float32(t + e[3])
        r = np.float32(r + e[2])
    n = np.float32(len(errs))
    return (float(t / n), float(r / n)) if len(errs) else (0.0, 0.0)
