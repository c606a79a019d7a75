"""Golden vectors for the Rectify step (CStereoMatching.cpp:117-168), generated with OpenCV (cv2, the published
implementation of the third-party calls the reference makes there: stereoRectify, initUndistortRectifyMap, remap,
getStructuringElement + erode).  The reference pins nothing at this boundary, so these vectors pin the
B200 path against OpenCV 4.13 as installed in the build image.

    python tests/golden/make_rectify_golden.py        # writes tests/golden/rectify_cv2.npz
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from reconstruction_b200 import stage, synth  # noqa: E402


def reference_rectify(K0, Rt0, K1, Rt1, origin, lowest, pyrm_num, imgs, masks):
    """Python transcription of what the reference's Rectify computes, call for call."""
    ow, oh = origin
    largest = (lowest[0] << (pyrm_num - 1), lowest[1] << (pyrm_num - 1))
    R = Rt1[:, :3] @ Rt0[:, :3].T
    T = -R @ Rt0[:, 3] + Rt1[:, 3]
    dist = np.zeros((4, 1))
    R1, R2, P1, P2, Q, _, _ = cv2.stereoRectify(K0, dist, K1, dist, (ow, oh), R, T, flags=0, alpha=-1, newImageSize=(ow, oh))
    R_final = Rt0[:, :3].T @ R1.T
    T_final = -Rt0[:, :3].T @ Rt0[:, 3]
    E = np.zeros((4, 4))
    E[3, 3] = 1
    E[:3, :3] = R_final.T
    E[:3, 3] = -R_final.T @ T_final
    Q = Q.copy()
    Q[3, 2] = -Q[3, 2]
    scale = float(lowest[0]) / ow * (1 << (pyrm_num - 1))
    out = {"R": R, "T": T, "R1": R1, "R2": R2, "Q": Q, "R_final": R_final, "T_final": T_final, "scale": scale}
    ks = 3 * (1 << (pyrm_num - 1))
    el = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (ks, ks))
    for j, (K, Rn, P) in enumerate(((K0, R1, P1), (K1, R2, P2))):
        P = P.copy()
        P[:2] *= scale
        m1, m2 = cv2.initUndistortRectifyMap(K, dist, Rn, P, largest, cv2.CV_16SC2)
        out[f"Pscaled{j}"] = P
        out[f"P{j}"] = P @ E
        out[f"map1_{j}"], out[f"map2_{j}"] = m1, m2
        out[f"image{j}"] = cv2.remap(imgs[j], m1, m2, cv2.INTER_LINEAR)
        rm = cv2.remap(masks[j], m1, m2, cv2.INTER_LINEAR)
        out[f"mask_remapped{j}"] = rm
        out[f"mask{j}"] = cv2.erode(rm, el)
    return out


def make_case(seed, n_cam, origin, lowest, pyrm_num):
    rng = np.random.default_rng(seed)
    ow, oh = origin
    cams = stage.rig_cameras(n_cam, ow, oh)
    (K0, Rt0), (K1, Rt1) = cams[0], cams[1]
    # perturb so that nothing is axis-aligned: small roll/pitch on camera 1, different focal length and centre
    a, b = np.deg2rad(1.7), np.deg2rad(-0.9)
    rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    rx = np.array([[1, 0, 0], [0, np.cos(b), -np.sin(b)], [0, np.sin(b), np.cos(b)]])
    Rt1 = np.concatenate([rz @ rx @ Rt1[:, :3], (rz @ rx @ Rt1[:, 3])[:, None] + np.array([[0.0], [3.5], [-2.0]])], axis=1)
    K1 = K1.copy()
    K1[0, 0] *= 1.013
    K1[1, 1] *= 1.009
    K1[0, 2] += 4.25
    K1[1, 2] -= 2.5
    imgs, masks = [], []
    for j in range(2):
        tex = synth.value_noise(oh, ow, rng)
        imgs.append(np.ascontiguousarray(np.clip(128 + 52 * tex, 0, 255).astype(np.uint8)))
        yy, xx = np.mgrid[0:oh, 0:ow]
        m = (((xx - ow * 0.52) / (ow * 0.38)) ** 2 + ((yy - oh * 0.5) / (oh * 0.41)) ** 2 <= 1.0)
        masks.append(np.where(m, 255, 0).astype(np.uint8))
    out = reference_rectify(K0, Rt0, K1, Rt1, origin, lowest, pyrm_num, imgs, masks)
    out.update({"K0": K0, "K1": K1, "Rt0": Rt0, "Rt1": Rt1, "origin": np.array(origin), "lowest": np.array(lowest),
                "pyrm_num": np.array(pyrm_num), "src_image0": imgs[0], "src_image1": imgs[1], "src_mask0": masks[0], "src_mask1": masks[1]})
    return out


if __name__ == "__main__":
    cases = {"a": make_case(11, 3, (320, 240), (80, 60), 2), "b": make_case(12, 5, (200, 152), (50, 38), 3)}
    flat = {f"{c}_{k}": v for c, d in cases.items() for k, v in d.items()}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rectify_cv2.npz")
    np.savez_compressed(path, **flat)
    print(path, os.path.getsize(path) // 1024, "KiB", "cv2", cv2.__version__)
