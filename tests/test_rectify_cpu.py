"""Rectify, calibration half (CStereoMatching.cpp:121-145) on the CPU: the host restatement of cv::stereoRectify and the matrix
bookkeeping around it against OpenCV 4.13 vectors (tests/golden/rectify_cv2.npz, made by tests/golden/make_rectify_golden.py)."""
import os

import numpy as np
import pytest

from reconstruction_b200 import capi


@pytest.fixture(scope="module")
def gold(golden_dir):
    capi.build()
    return np.load(os.path.join(golden_dir, "rectify_cv2.npz"))


@pytest.mark.parametrize("c", ["a", "b"])
def test_calibration_matches_opencv(gold, c):
    o = capi.rectify_calib(gold[c + "_K0"], gold[c + "_Rt0"], gold[c + "_K1"], gold[c + "_Rt1"], gold[c + "_origin"], gold[c + "_lowest"][0],
                           int(gold[c + "_pyrm_num"]), opencv_compat=413)  # the vectors come from OpenCV 4.13
    # rotations, projection matrices and Q: bit for bit (the principal points go through float32 inside OpenCV; reproduced)
    assert np.array_equal(o["R_new"][0], gold[c + "_R1"]) or np.abs(o["R_new"][0] - gold[c + "_R1"]).max() < 1e-15
    assert np.abs(o["R_new"][1] - gold[c + "_R2"]).max() < 1e-15
    assert np.array_equal(o["Q"], gold[c + "_Q"])
    for j in (0, 1):
        assert np.array_equal(o["P_scaled"][j], gold[c + f"_Pscaled{j}"])
        ref = gold[c + f"_P{j}"]
        assert np.abs(o["P_final"][j] - ref).max() <= 1e-13 * np.abs(ref).max()
    assert np.abs(o["R_final"] - gold[c + "_R_final"]).max() < 1e-15
    assert np.abs(o["T_final"] - gold[c + "_T_final"]).max() <= 1e-13 * np.abs(gold[c + "_T_final"]).max()


def test_stereo_rectify_alone(gold):
    c = "a"
    R1, R2, P1, P2, Q = capi.stereo_rectify_host(gold[c + "_K0"], gold[c + "_K1"], gold[c + "_origin"], gold[c + "_R"], gold[c + "_T"])
    assert np.abs(R1 - gold[c + "_R1"]).max() < 1e-15 and np.abs(R2 - gold[c + "_R2"]).max() < 1e-15
    assert np.abs(R1 @ R1.T - np.eye(3)).max() < 1e-14
    assert P1[0, 0] == P1[1, 1] == P2[0, 0] == (gold[c + "_K0"][1, 1] + gold[c + "_K1"][1, 1]) / 2
    assert P1[1, 2] == P2[1, 2] and P1[0, 3] == 0 and P2[0, 3] != 0  # horizontal pair: shared cy, baseline in P2
    Qg = gold[c + "_Q"].copy()
    Qg[3, 2] = -Qg[3, 2]  # the golden Q carries the reference's sign flip (:138)
    assert np.array_equal(Q, Qg)


def test_opencv_vintage_switch(gold):
    """cv::stereoRectify of the OpenCV the reference links (2.4.5) against today's (4.13): the smaller focal length instead of the
    mean, the far image corners at (nx, ny) instead of (nx-1, ny-1).  2.4.5 cannot be run here; what is asserted is that the two
    documented rules - and nothing else - separate the two modes, and that the library default is the reference's dependency."""
    c = "a"
    K0, K1, R, T, size = gold[c + "_K0"], gold[c + "_K1"], gold[c + "_R"], gold[c + "_T"], gold[c + "_origin"]
    new = capi.stereo_rectify_host(K0, K1, size, R, T, opencv_compat=413)
    old = capi.stereo_rectify_host(K0, K1, size, R, T, opencv_compat=245)
    assert np.array_equal(new[0], old[0]) and np.array_equal(new[1], old[1])  # the rotations do not depend on the version
    f_new, f_old = new[2][0, 0], old[2][0, 0]
    assert f_new == (K0[1, 1] + K1[1, 1]) / 2 and f_old == min(K0[1, 1], K1[1, 1]) and f_old < f_new
    # same rig with equal focal lengths: only the corner rule is left, and it moves the principal points by about half a pixel
    K1e = K1.copy()
    K1e[0, 0], K1e[1, 1] = K0[0, 0], K0[1, 1]
    n2 = capi.stereo_rectify_host(K0, K1e, size, R, T, opencv_compat=413)
    o2 = capi.stereo_rectify_host(K0, K1e, size, R, T, opencv_compat=245)
    assert n2[2][0, 0] == o2[2][0, 0]
    d = n2[2][:2, 2] - o2[2][:2, 2]
    assert np.all(np.abs(d - 0.5) < 0.05), d
    # the default of the one-call calibration entry is 2.4.5 (SB200_OPENCV_COMPAT overrides it)
    L = int(gold[c + "_pyrm_num"])
    args = (gold[c + "_K0"], gold[c + "_Rt0"], gold[c + "_K1"], gold[c + "_Rt1"], gold[c + "_origin"], gold[c + "_lowest"][0], L)
    if "SB200_OPENCV_COMPAT" not in os.environ:
        dflt, v245 = capi.rectify_calib(*args), capi.rectify_calib(*args, opencv_compat=245)
        assert all(np.array_equal(dflt[k], v245[k]) for k in dflt)
        assert not np.array_equal(dflt["Q"], capi.rectify_calib(*args, opencv_compat=413)["Q"])
