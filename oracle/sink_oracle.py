"""CPU checker for the sink's per-pair filter — TEST INFRASTRUCTURE, never on the product path.

Restates what CCloudOptimization::filter does to one pair's points before meshing
(/root/reference/CloudOptimization/CCloudOptimization.cpp:64-121):
    pcl::StatisticalOutlierRemoval (setMeanK, setStddevMulThresh)      :79-83
    pcl::NormalEstimationOMP with setRadiusSearch                      :99-106
    flip of every normal towards CamCenter[idx]                        :109-116
PCL (1.6 / 1.7, Readme.md:7) and FLANN are third-party dependencies that are NOT under /root/reference and are not installed
here: **parity unpinned** against PCL itself.  The published algorithms are restated (pcl/filters/impl/
statistical_outlier_removal.hpp, pcl/features/normal_3d.h, flann L2_Simple) with neighbour search by scipy's cKDTree; the
brute-force functions at the bottom pin this file in tests/test_sink_cpu.py.

Arithmetic: points float32 (pcl::PointXYZ); squared distance ((dx*dx + dy*dy) + dz*dz) in float32; sqrt in float32; sums of the
k distances in float64 (exact for floats of similar magnitude); mean / sample stddev over all points by sequential float64
sums (np.cumsum is sequential); covariance in float64 about the query point; np.linalg.eigh.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree


def _d2_f32(a, b):
    """FLANN L2_Simple in float32, one rounding per operation."""
    d = (a - b).astype(np.float32)
    sq = (d * d).astype(np.float32)
    return ((sq[..., 0] + sq[..., 1]).astype(np.float32) + sq[..., 2]).astype(np.float32)


def sor_mean_distances(p32, mean_k):
    n = len(p32)
    kq = min(n, mean_k + 1 + 8)
    tree = cKDTree(p32.astype(np.float64))
    _, idx = tree.query(p32.astype(np.float64), k=kq)
    idx = idx.reshape(n, kq)
    d2 = _d2_f32(p32[:, None, :], p32[idx])
    d2.sort(axis=1)
    take = min(mean_k + 1, kq)
    dist = np.sqrt(d2[:, :take]).astype(np.float32)  # correctly rounded float32 sqrt
    return dist.astype(np.float64).sum(axis=1) / float(mean_k)


def sor(p32, mean_k, std_mul):
    d = sor_mean_distances(p32, mean_k)
    n = len(d)
    s = float(np.cumsum(d)[-1])
    sq = float(np.cumsum(d * d)[-1])
    mean = s / n
    var = (sq - s * s / n) / (n - 1) if n > 1 else 0.0
    std = float(np.sqrt(max(var, 0.0)))
    thr = mean + std_mul * std
    keep = ~(d > thr)
    return keep, d, {"mean": mean, "stddev": std, "threshold": thr}


def normals(p32, radius, cam_center):
    """p32: the KEPT points.  Returns (normals [n,3] f32, curvature [n] f32, neighbour counts, eigen gap) — gap = (l1 - l0) / trace, the
    conditioning of the normal direction."""
    n = len(p32)
    tree = cKDTree(p32.astype(np.float64))
    r2 = np.float32(radius * radius)
    nb = tree.query_ball_point(p32.astype(np.float64), radius * (1 + 1e-5))
    out = np.full((n, 3), np.nan, np.float32)
    curv = np.full(n, np.nan, np.float32)
    cnt = np.zeros(n, np.int64)
    gap = np.zeros(n)
    cam = np.asarray(cam_center, np.float64).astype(np.float32)
    for i in range(n):
        j = np.asarray(nb[i], np.int64)
        j = j[_d2_f32(p32[j], p32[i]) < r2]
        cnt[i] = len(j)
        if len(j) < 3:
            continue
        d = p32[j].astype(np.float64) - p32[i].astype(np.float64)
        m = d.mean(axis=0)
        cov = d.T @ d / len(j) - np.outer(m, m)
        w, v = np.linalg.eigh(cov)
        nv = v[:, 0].astype(np.float32)
        tr = cov.trace()
        curv[i] = np.float32(abs(w[0] / tr)) if tr != 0 else np.float32(0)
        gap[i] = (w[1] - w[0]) / tr if tr != 0 else 0.0
        q = p32[i]
        if np.float32(-q[0] * nv[0] - q[1] * nv[1] - q[2] * nv[2]) < 0:  # default viewpoint (0, 0, 0)
            nv = -nv
        if float(np.dot(nv.astype(np.float64), (cam - q).astype(np.float64))) < 0:  # CCloudOptimization.cpp:109-116
            nv = -nv
        out[i] = nv
    return out, curv, cnt, gap


def sink_filter(xyz_f64, mean_k, std_mul, radius, cam_center):
    p32 = np.ascontiguousarray(xyz_f64, np.float64).astype(np.float32)
    keep, d, stats = sor(p32, mean_k, std_mul)
    kept = np.nonzero(keep)[0].astype(np.int32)
    nrm, curv, cnt, gap = normals(p32[kept], radius, cam_center)
    rec = np.concatenate([p32[kept], nrm, curv[:, None]], axis=1).astype(np.float32)
    return rec, kept, dict(stats, mean_dist=d, neighbours=cnt, eigen_gap=gap)


# ---- brute force (O(n^2), small n): pins the functions above -------------------------------------------------------------
def brute_mean_distances(p32, mean_k):
    out = np.zeros(len(p32))
    for i in range(len(p32)):
        d2 = np.sort(_d2_f32(p32, p32[i]))
        out[i] = np.sqrt(d2[: mean_k + 1]).astype(np.float32).astype(np.float64).sum() / mean_k
    return out


def brute_neighbour_counts(p32, radius):
    r2 = np.float32(radius * radius)
    return np.array([int((_d2_f32(p32, p32[i]) < r2).sum()) for i in range(len(p32))])
