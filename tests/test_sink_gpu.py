"""sb200_sink_filter (the sink's per-pair statistical outlier removal + normals + orientation, on the GPU) against
oracle/sink_oracle.py.  The kept set and the SOR distances must be identical (the sums involved are exact in float64); normals are
eigenvectors, compared within a tolerance where the plane fit is well conditioned."""
import os
import subprocess

import numpy as np
import pytest

from oracle import sink_oracle as so
from reconstruction_b200 import capi, stage, synth

pytestmark = pytest.mark.gpu


def surface_cloud(n_side, seed, outliers=40):
    """A disparity-map-like sampling of a bumpy surface ~1000 mm from the camera (0.27 mm spacing) + isolated far points."""
    rng = np.random.default_rng(seed)
    u, v = np.meshgrid(np.arange(n_side) * 0.27 - n_side * 0.135, np.arange(n_side) * 0.27 - n_side * 0.135)
    z = 1000 + 6 * np.sin(u / 9.0) * np.cos(v / 7.0) + rng.normal(0, 0.03, u.shape)
    p = np.stack([u + rng.normal(0, 0.02, u.shape), v + rng.normal(0, 0.02, u.shape), z], -1).reshape(-1, 3)
    far = rng.integers(0, len(p), outliers)
    p[far] += rng.uniform(-1, 1, (outliers, 3)) * np.array([[20, 20, 60]])
    p[5] = p[6]  # duplicate
    return p


def compare(p64, mean_k, std_mul, radius, cam, expect_normals=True):
    rec, kept, st = capi.sink_filter(p64, mean_k, std_mul, radius, cam)
    orec, okept, ost = so.sink_filter(p64, mean_k, std_mul, radius, cam)
    assert st["mean"] == ost["mean"] and st["stddev"] == ost["stddev"] and st["threshold"] == ost["threshold"]
    assert np.array_equal(kept, okept)
    assert np.array_equal(rec[:, :3].view(np.int32), orec[:, :3].view(np.int32))
    nan_o = np.isnan(orec[:, 3])
    assert np.array_equal(np.isnan(rec[:, 3]), nan_o) and np.array_equal(np.isnan(rec[:, 6]), nan_o)
    good = ~nan_o & (ost["eigen_gap"] > 1e-3)
    camdot = np.abs(np.einsum("ij,ij->i", orec[:, 3:6].astype(np.float64), np.asarray(cam)[None] - orec[:, :3].astype(np.float64)))
    sure = good & (camdot > 1e-3)  # away from grazing views the sign is determined
    assert good.sum() > 0.5 * len(kept) or not expect_normals
    if not good.any():
        return rec, kept, st
    assert np.abs(rec[sure, 3:6] - orec[sure, 3:6]).max() < 2e-5
    graze = good & ~sure
    if graze.any():
        assert (np.abs(np.abs(np.einsum("ij,ij->i", rec[graze, 3:6], orec[graze, 3:6])) - 1) < 1e-4).all()
    assert np.allclose(rec[good, 6], orec[good, 6], rtol=1e-4, atol=1e-7)
    return rec, kept, st


def test_surface_with_outliers():
    p = surface_cloud(200, 1)
    rec, kept, st = compare(p, 100, 1.0, 2.5, [30.0, -40.0, 0.0])  # the reference's parameters (CReconstruction.cpp:19)
    assert 0.8 * len(p) < len(kept) < len(p) and st["widened_queries"] > 0  # isolated points needed a wider ring
    # normals of a surface facing the camera point towards it
    ok = ~np.isnan(rec[:, 3])
    assert (np.einsum("ij,ij->i", rec[ok, 3:6], np.array([[30.0, -40.0, 0.0]]) - rec[ok, :3]) >= 0).all()


def test_small_and_degenerate_clouds():
    rng = np.random.default_rng(3)
    compare(rng.uniform(-5, 5, (60, 3)) + [0, 0, 500], 100, 1.0, 2.5, [0.0, 0.0, 0.0], False)   # fewer points than meanK + 1
    compare(rng.uniform(-50, 50, (3000, 3)) + [0, 0, 500], 8, 0.5, 2.5, [0.0, 0.0, 0.0], False)  # sparse: most rings widen, few normal neighbours
    p = surface_cloud(60, 4, outliers=5)
    p[100] = [np.nan, 0, 0]
    p[200] = [0, np.inf, 0]
    rec, kept, _ = capi.sink_filter(p, 20, 1.0, 2.5, [0.0, 0.0, 0.0])
    fin = np.isfinite(p).all(axis=1)
    orec, okept, _ = so.sink_filter(p[fin], 20, 1.0, 2.5, [0.0, 0.0, 0.0])
    assert 100 not in kept and 200 not in kept
    assert np.array_equal(kept, np.nonzero(fin)[0][okept]) and np.array_equal(rec[:, :3].view(np.int32), orec[:, :3].view(np.int32))


def test_points_of_a_matched_pair_and_cli(tmp_path):
    """The matcher's own output through the sink, and the CLI writing tmp/cloud_filter.ply in the PointNormal layout (:119)."""
    L, w0, h0 = 3, 64, 48
    cfg, pairs = stage.write_dataset(str(tmp_path), L, w0, h0, n_pairs=1)
    sp = pairs[0]
    g = capi.StereoB200(L, w0, h0)
    g.set_pair(*sp.image, *sp.mask)
    g.set_calib(sp.Q, sp.R_final, sp.T_final)
    n = g.match_pair()
    xyz, _, _ = g.get_points(n)
    cams = stage.rig_cameras(2, *sp.origin_size)
    center = -cams[0][1][:, :3].T @ cams[0][1][:, 3]
    compare(xyz, 30, 1.0, 15.0, center)  # a radius that suits this coarse sampling (~4 mm between points)
    rec, kept, st = compare(xyz, 100, 1.0, 2.5, center, False)  # the reference's parameters: nearly every neighbourhood is too small
    host = os.path.join(os.path.dirname(capi.HERE), "reconstruction_b200", "host")
    subprocess.run(["make", "-s", "-C", host], check=True)
    r = subprocess.run([os.path.join(host, "reconstruction"), cfg], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 0 and "Cloud after filtering" in r.stdout, r.stdout + r.stderr
    raw = open(tmp_path / "tmp" / "cloud_filter.ply", "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    assert b"property float normal_x" in head and b"property float curvature" in head and f"element vertex {len(kept)}".encode() in head
    got = np.frombuffer(body, np.float32).reshape(-1, 7)
    assert np.array_equal(got[:, :3].view(np.int32), rec[:, :3].view(np.int32))  # the same kept points
    both = ~np.isnan(got[:, 3]) & ~np.isnan(rec[:, 3])
    assert np.array_equal(np.isnan(got[:, 3]), np.isnan(rec[:, 3]))
    assert (np.abs(np.abs(np.einsum("ij,ij->i", got[both, 3:6], rec[both, 3:6])) - 1) < 1e-5).all()
    assert os.path.exists(str(tmp_path / "out.ply.normals.ply"))


def test_committed_golden():
    """The same call against tests/golden/sink_small.npz (written by the checker, make_sink_golden.py)."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sink_small.npz"))
    rec, kept, st = capi.sink_filter(g["xyz"], int(g["mean_k"]), float(g["std_mul"]), float(g["radius"]), g["cam"])
    assert np.array_equal(kept, g["kept"]) and [st["mean"], st["stddev"], st["threshold"]] == g["stats"].tolist()
    assert np.array_equal(rec[:, :3].view(np.int32), g["records"][:, :3].view(np.int32))
    ok = ~np.isnan(g["records"][:, 3]) & (g["eigen_gap"] > 1e-3)
    assert np.array_equal(np.isnan(rec[:, 3]), np.isnan(g["records"][:, 3]))
    assert np.abs(rec[ok, 3:6] - g["records"][ok, 3:6]).max() < 2e-5 and np.allclose(rec[ok, 6], g["records"][ok, 6], rtol=1e-4, atol=1e-7)
