"""The sink checker (oracle/sink_oracle.py) against brute force, on CPU.  PCL is absent (parity unpinned against it):
the k-d-tree based restatement is pinned by O(n^2) evaluation of the same definitions."""
import numpy as np

from oracle import sink_oracle as so


def cloud(n, seed, outliers=6):
    rng = np.random.default_rng(seed)
    u, v = rng.uniform(-20, 20, n), rng.uniform(-20, 20, n)
    z = 900 + 0.02 * (u * u + v * v) + rng.normal(0, 0.05, n)
    p = np.stack([u, v, z], 1)
    p[:outliers] += rng.uniform(30, 60, (outliers, 3))
    p[10] = p[11]  # a duplicate point
    return p


def test_sor_against_brute_force():
    p64 = cloud(500, 1)
    p32 = p64.astype(np.float32)
    for k in (4, 20):
        d = so.sor_mean_distances(p32, k)
        b = so.brute_mean_distances(p32, k)
        assert np.array_equal(d, b), k  # sums of float32 values in float64 are exact: order-independent
    keep, d, st = so.sor(p32, 20, 1.0)
    n = len(d)
    s = sq = 0.0
    for v in d.tolist():  # the sequential float64 sums of statistical_outlier_removal.hpp (sum() compensates since Python 3.12)
        s += v
        sq += v * v
    mean = s / n
    assert st["mean"] == mean
    var = (sq - s * s / n) / (n - 1)
    assert st["stddev"] == np.sqrt(var) and st["threshold"] == mean + np.sqrt(var)
    assert not keep[:6].any() and keep[6:].mean() > 0.9  # the far points go, the surface stays


def test_normals_against_brute_force():
    p64 = cloud(400, 2, outliers=3)
    p32 = p64.astype(np.float32)
    cam = np.array([0.0, 0.0, 0.0])
    nrm, curv, cnt, gap = so.normals(p32, 2.5, cam)
    assert np.array_equal(cnt, so.brute_neighbour_counts(p32, 2.5))
    few = cnt < 3
    assert few[:3].all() and np.isnan(nrm[few]).all() and np.isnan(curv[few]).all()
    ok = ~few & (gap > 1e-3)
    assert ok.sum() > 300
    assert np.allclose(np.linalg.norm(nrm[ok], axis=1), 1, atol=1e-6)
    # oriented towards the camera at the origin: the paraboloid z = 900 + ... faces -z
    assert (np.einsum("ij,ij->i", nrm[ok], cam[None] - p32[ok]) >= 0).all() and (nrm[ok][:, 2] < 0).mean() > 0.98
    # plane fit: the normal is orthogonal to the neighbourhood's principal directions (check one point by hand)
    i = int(np.nonzero(ok)[0][5])
    j = np.nonzero(so._d2_f32(p32, p32[i]) < np.float32(6.25))[0]
    d = p32[j].astype(np.float64) - p32[i]
    cov = np.cov(d.T, bias=True)
    w, v = np.linalg.eigh(cov)
    assert abs(abs(np.dot(v[:, 0], nrm[i])) - 1) < 1e-6 and abs(curv[i] - w[0] / w.sum()) < 1e-6


def test_sink_filter_record_layout():
    p64 = cloud(300, 3)
    rec, kept, info = so.sink_filter(p64, 20, 1.0, 2.5, [0, 0, 0])
    assert rec.dtype == np.float32 and rec.shape == (len(kept), 7) and np.array_equal(rec[:, :3], p64[kept].astype(np.float32))
    assert np.all(np.diff(kept) > 0) and 0 not in kept


def test_oracle_reproduces_the_committed_golden():
    """tests/golden/sink_small.npz (make_sink_golden.py): the checker has not drifted."""
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sink_small.npz"))
    rec, kept, info = so.sink_filter(g["xyz"], int(g["mean_k"]), float(g["std_mul"]), float(g["radius"]), g["cam"])
    assert np.array_equal(kept, g["kept"]) and np.array_equal(info["mean_dist"], g["mean_dist"])
    assert [info["mean"], info["stddev"], info["threshold"]] == g["stats"].tolist()
    assert np.array_equal(rec[:, :3].view(np.int32), g["records"][:, :3].view(np.int32))
    ok = ~np.isnan(g["records"][:, 3]) & (g["eigen_gap"] > 1e-3)
    assert np.abs(rec[ok, 3:] - g["records"][ok, 3:]).max() < 1e-6  # eigh across numpy / LAPACK builds
