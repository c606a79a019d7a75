"""Generate the committed golden fixtures (run in the BUILD container, where /root/reference exists).

    python tests/golden/make_golden.py

* stereo_small.npz — one synthetic staged pair (96x72 top level, 2 pyramid levels) pushed through
  oracle/_ref, i.e. the reference's OWN CStereoMatching.cpp / CManageData.cpp compiled unmodified
  (oracle/ref_build): every level's images/masks/margins, the disparity maps after every
  MatchOneLayer stage (CStereoMatching.cpp:51-109), the Rematch bounds, and the triangulated points.
* pyrdown_cv2.npz / erode_cv2.npz — cv2 4.13 outputs for pyrDown and ellipse erode (OpenCV is a
  third-party dependency of the reference, 2.4.5 in its tree, not buildable here).
* ncc_kat.npz — WindowToVec / NCC known answers from the reference's own primitive.

The fixtures travel to the GPU box (which has no /root/reference); tests compare both the oracle
port and the CUDA path with them.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import pyoracle as po  # noqa: E402
from reconstruction_b200 import synth  # noqa: E402


def stereo_fixture(path, lowest_w=48, lowest_h=36, levels=2, pair_id=3):
    sp = synth.make_pair(lowest_w, lowest_h, levels, pair_id=pair_id)
    o = po.CpuStereo("ref", levels, lowest_w, lowest_h)
    o.set_pair(*sp.image, *sp.mask)
    o.set_calib(sp.Q, sp.R_final, sp.T_final)
    out = {
        "lowest": np.array([lowest_w, lowest_h, levels], np.int32),
        "img0": sp.image[0], "img1": sp.image[1], "mask0": sp.mask[0], "mask1": sp.mask[1],
        "Q": sp.Q, "R_final": sp.R_final, "T_final": sp.T_final, "origin": np.array(sp.origin_size, np.int32),
    }
    for lv in range(levels):
        for v in (0, 1):
            img, mask = o.get_level(lv, v)
            out[f"L{lv}_img{v}"] = img
            out[f"L{lv}_mask{v}"] = mask
        for st in range(1, 11):
            o.run_stage(lv, st)
            if st == 1:
                out[f"L{lv}_margins"] = o.get_margins()
                continue
            for d in (0, 1):
                out[f"L{lv}_S{st}_d{d}"] = o.get_disparity(d, lv)
            if st == 6:
                for d in (0, 1):
                    bl, br = o.get_rematch_bounds(d, lv)
                    out[f"L{lv}_BL{d}"] = bl
                    out[f"L{lv}_BR{d}"] = br
    out["points"] = o.to_cloud()
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", len(out["points"]), "points")


def cv2_fixtures():
    import cv2

    rng = np.random.default_rng(7)
    d = {}
    for i, (h, w, cn) in enumerate([(37, 53, 3), (64, 48, 1), (5, 9, 3), (2, 2, 1), (1, 7, 1), (96, 128, 3)]):
        src = rng.integers(0, 256, (h, w, cn) if cn > 1 else (h, w), dtype=np.uint8)
        d[f"src{i}"] = src
        d[f"dst{i}"] = cv2.pyrDown(src)
    np.savez_compressed(os.path.join(HERE, "pyrdown_cv2.npz"), **d)
    e = {}
    for i, (h, w, ks) in enumerate([(40, 60, 3), (64, 64, 7), (50, 70, 12), (30, 30, 1), (48, 80, 6), (33, 21, 2)]):
        src = np.where(rng.random((h, w)) < 0.97, 255, 0).astype(np.uint8)
        src[rng.integers(0, h, 5), rng.integers(0, w, 5)] = rng.integers(0, 255, 5)
        k = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (ks, ks))
        e[f"src{i}"] = src
        e[f"ks{i}"] = np.int32(ks)
        e[f"kernel{i}"] = k
        e[f"dst{i}"] = cv2.erode(src, k)
    np.savez_compressed(os.path.join(HERE, "erode_cv2.npz"), **e)


def ncc_fixture():
    rng = np.random.default_rng(11)
    img_l = rng.integers(0, 256, (12, 40, 3), dtype=np.uint8)
    img_r = rng.integers(0, 256, (12, 40, 3), dtype=np.uint8)
    img_l[0:5, 0:5] = 77  # flat window -> norm 1, NCC exactly 0 (Q10)
    cases, norms, vecs, vals = [], [], [], []
    for ws in (3, 5):
        for (y0, xl, xr) in [(0, 0, 3), (2, 7, 9), (5, 20, 11), (6, 30, 30), (1, 13, 2)]:
            n, v = po.window_to_vec("ref", img_l, y0, xl, ws)
            cases.append((ws, y0, xl, xr))
            norms.append(n)
            vecs.append(np.pad(v, (0, 75 - v.size)))
            vals.append(po.ncc_match_value("ref", img_l, img_r, y0, xl, xr, ws))
    np.savez_compressed(
        os.path.join(HERE, "ncc_kat.npz"), img_l=img_l, img_r=img_r, cases=np.array(cases, np.int32),
        norms=np.array(norms), vecs=np.array(vecs), vals=np.array(vals),
    )


if __name__ == "__main__":
    po.build()
    assert po.available("ref"), "needs /root/reference to build oracle/_ref"
    stereo_fixture(os.path.join(HERE, "stereo_small.npz"))
    cv2_fixtures()
    ncc_fixture()
