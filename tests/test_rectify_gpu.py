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
    """initUndistortRectifyMap: OpenCV accumulates the source position along each row; k_rectify_row_starts / k_rectify_maps replay
    that accumulation, so the fixed-point maps - and with them the remapped image and the eroded mask - are OpenCV's bit for bit."""
    g, L, _, _ = _ctx(gold, c)
    o = capi.rectify_calib(gold[c + "_K0"], gold[c + "_Rt0"], gold[c + "_K1"], gold[c + "_Rt1"], gold[c + "_origin"], gold[c + "_lowest"][0], L,
                           opencv_compat=413)
    for j in (0, 1):
        g.rectify_view(j, gold[c + f"_src_image{j}"], gold[c + f"_src_mask{j}"], gold[c + f"_K{j}"], o["R_new"][j], o["P_scaled"][j])
        m1, m2 = g.get_rectify_maps()
        assert np.array_equal(m1, gold[c + f"_map1_{j}"]), f"view {j}: {int((m1 != gold[c + f'_map1_{j}']).sum())} integer map entries differ"
        assert np.array_equal(m2, gold[c + f"_map2_{j}"]), f"view {j}: {int((m2 != gold[c + f'_map2_{j}']).sum())} fractional map entries differ"
        img, mask = g.get_level(L - 1, j)
        assert np.array_equal(img, gold[c + f"_image{j}"]), f"view {j}: rectified image"
        assert np.array_equal(mask, gold[c + f"_mask{j}"]), f"view {j}: rectified, eroded mask"


def sum_in_order(terms):
    s = 0.0
    for t in terms:
        s += t
    return s


def test_maps_wide_rows(gold):
    """a top level wider than one block pass (4096 columns) and not a multiple of the 32-column restart interval: the device maps
    against the same accumulation done on the host in numpy"""
    W, H = 4200 + 8, 40
    L, w0, h0 = 1, W, H
    g = capi.StereoB200(L, w0, h0, W, H)
    K = np.array([[3900.0, 0, W / 2 + 3.25], [0, 3950.0, H / 2 - 1.5], [0, 0, 1]])
    a = np.deg2rad(0.8)
    Rn = np.array([[np.cos(a), -np.sin(a), 0.002], [np.sin(a), np.cos(a), -0.001], [-0.002, 0.001, 1.0]])
    P = np.zeros((3, 4))
    P[:3, :3] = np.array([[3800.0, 0, W / 2], [0, 3800.0, H / 2], [0, 0, 1]])
    src = np.zeros((H, W, 3), np.uint8)
    g.rectify_view(0, src, src[..., 0].copy(), K, Rn, P)
    m1, m2 = g.get_rectify_maps()
    # P[:, :3] * R_new term by term in the library's order (s = 0; s += a*b), not through BLAS: the last bit matters
    m = np.array([[sum_in_order([float(P[r, k]) * float(Rn[k, q]) for k in range(3)]) for q in range(3)] for r in range(3)]).ravel()
    d = 1. / (m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]))
    ir = np.array([(m[4] * m[8] - m[5] * m[7]) * d, (m[2] * m[7] - m[1] * m[8]) * d, (m[1] * m[5] - m[2] * m[4]) * d,
                   (m[5] * m[6] - m[3] * m[8]) * d, (m[0] * m[8] - m[2] * m[6]) * d, (m[2] * m[3] - m[0] * m[5]) * d,
                   (m[3] * m[7] - m[4] * m[6]) * d, (m[1] * m[6] - m[0] * m[7]) * d, (m[0] * m[4] - m[1] * m[3]) * d])
    i = np.arange(H, dtype=np.float64)
    x, y, w = i * ir[1] + ir[2], i * ir[4] + ir[5], i * ir[7] + ir[8]
    for j in range(W):  # one addition per column, vectorised over the rows
        iw = 1. / w
        u, v = K[0, 0] * (x * iw) + K[0, 2], K[1, 1] * (y * iw) + K[1, 2]
        iu, iv = np.rint(u * 32).astype(np.int64), np.rint(v * 32).astype(np.int64)
        assert np.array_equal(m1[:, j, 0], (iu >> 5).astype(np.int16)) and np.array_equal(m1[:, j, 1], (iv >> 5).astype(np.int16)), j
        assert np.array_equal(m2[:, j], ((iv & 31) * 32 + (iu & 31)).astype(np.uint16)), j
        x, y, w = x + ir[0], y + ir[3], w + ir[6]
    g.close()


def test_native_chain_matches_oracle(gold, oracle):
    """config -> Rectify on the device -> pyramid -> matcher -> points, all native; the CPU oracle gets the GPU-rectified frames
    and the host-computed calibration, everything after that must agree bit for bit."""
    c = "a"
    g, L, w0, h0 = _ctx(gold, c)
    cal = capi.rectify_calib(gold[c + "_K0"], gold[c + "_Rt0"], gold[c + "_K1"], gold[c + "_Rt1"], gold[c + "_origin"], w0, L)  # default: 2.4.5 rules
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
