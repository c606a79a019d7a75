"""Rectify, image half (CStereoMatching.cpp:144-158) on the B200 against OpenCV 4.13 vectors: the fixed-point maps of
initUndistortRectifyMap, remap (INTER_LINEAR) of image and mask, the ellipse erode; then the native chain config -> points with
the CPU oracle fed the GPU-rectified frames."""
import os

import numpy as np
import pytest

from reconstruction_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold(golden_dir):
    import torch

    assert torch.cuda.is_available(), "these tests need the B200"
    capi.build()
    return np.load(os.path.join(golden_dir, "rectify_cv2.npz"))


def _ctx(gold, c):
    w0, h0 = (int(v) for v in gold[c + "_lowest"])
    L = int(gold[c + "_pyrm_num"])
    ow, oh = (int(v) for v in gold[c + "_origin"])
    return capi.StereoB200(L, w0, h0, ow, oh), L, w0, h0


@pytest.mark.parametrize("c", ["a", "b"])
def test_remap_and_erode_with_opencv_maps(gold, c):
    """Same maps in, same bytes out: remap and erode are integer arithmetic."""
    g, L, _, _ = _ctx(gold, c)
    for j in (0, 1):
        g.set_rectify_maps(gold[c + f"_map1_{j}"], gold[c + f"_map2_{j}"])
        g.rectify_view(j, gold[c + f"_src_image{j}"], gold[c + f"_src_mask{j}"], use_given_maps=True)
        assert np.array_equal(g.get_remapped_mask(), gold[c + f"_mask_remapped{j}"]), f"view {j}: remapped mask"
        img, mask = g.get_level(L - 1, j)
        assert np.array_equal(img, gold[c + f"_image{j}"]), f"view {j}: remapped image"
        assert np.array_equal(mask, gold[c + f"_mask{j}"]), f"view {j}: eroded mask"


@pytest.mark.parametrize("c", ["a", "b"])
def test_maps_against_opencv(gold, c):
    """initUndistortRectifyMap: OpenCV accumulates along the row (and differently per SIMD path); the closed form agrees except
    where u*32 sits on a rounding boundary: at most one 1/32-pixel step, on a negligible share of the entries."""
    g, L, _, _ = _ctx(gold, c)
    o = capi.rectify_calib(gold[c + "_K0"], gold[c + "_Rt0"], gold[c + "_K1"], gold[c + "_Rt1"], gold[c + "_origin"], gold[c + "_lowest"][0], L)
    for j in (0, 1):
        g.rectify_view(j, gold[c + f"_src_image{j}"], gold[c + f"_src_mask{j}"], gold[c + f"_K{j}"], o["R_new"][j], o["P_scaled"][j])
        m1, m2 = g.get_rectify_maps()
        gm1, gm2 = gold[c + f"_map1_{j}"], gold[c + f"_map2_{j}"]
        fx = m1[..., 0].astype(np.int64) * 32 + (m2 & 31)
        fy = m1[..., 1].astype(np.int64) * 32 + ((m2 >> 5) & 31)
        gx = gm1[..., 0].astype(np.int64) * 32 + (gm2 & 31)
        gy = gm1[..., 1].astype(np.int64) * 32 + ((gm2 >> 5) & 31)
        d = np.maximum(np.abs(fx - gx), np.abs(fy - gy))
        assert d.max() <= 1, f"view {j}: map differs by {d.max()} fixed-point steps"
        assert (d != 0).mean() <= 1e-3, f"view {j}: {(d != 0).mean():.2e} of the map entries differ"
        img, mask = g.get_level(L - 1, j)
        assert (img != gold[c + f"_image{j}"]).any(axis=2).mean() <= 2e-3
        assert (mask != gold[c + f"_mask{j}"]).mean() <= 2e-3


def test_native_chain_matches_oracle(gold, oracle):
    """config -> Rectify on the device -> pyramid -> matcher -> points, all native; the CPU oracle gets the GPU-rectified frames
    and the host-computed calibration, everything after that must agree bit for bit."""
    c = "a"
    g, L, w0, h0 = _ctx(gold, c)
    cal = capi.rectify_calib(gold[c + "_K0"], gold[c + "_Rt0"], gold[c + "_K1"], gold[c + "_Rt1"], gold[c + "_origin"], w0, L)
    for j in (0, 1):
        g.rectify_view(j, gold[c + f"_src_image{j}"], gold[c + f"_src_mask{j}"], gold[c + f"_K{j}"], cal["R_new"][j], cal["P_scaled"][j])
    g.pair_build()
    g.set_calib(cal["Q"], cal["R_final"], cal["T_final"])
    n = g.match_pair()
    frames = [g.get_level(L - 1, j) for j in (0, 1)]
    o = oracle.CpuStereo("port", L, w0, h0, *(int(v) for v in gold[c + "_origin"]))
    o.set_pair(frames[0][0], frames[1][0], frames[0][1], frames[1][1])
    o.set_calib(cal["Q"], cal["R_final"], cal["T_final"])
    assert o.match_pair() == n
    for d in (0, 1):
        assert np.array_equal(g.get_disparity(d).view(np.int64), o.get_disparity(d, L - 1).view(np.int64))
    xyz, _, _ = g.get_points(n)
    assert np.array_equal(xyz.view(np.int64), o.to_cloud().view(np.int64))
