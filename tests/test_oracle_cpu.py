"""Pin the CPU oracle (oracle/stereo_oracle.cpp, the restatement) against
  (1) the committed golden fixtures produced by the reference's own sources (tests/golden/make_golden.py),
  (2) oracle/_ref live, when it was built (build container; absent on the GPU box),
  (3) cv2 4.13 vectors for the two OpenCV pieces the path restates (pyrDown, ellipse erode),
  (4) WindowToVec / NCC known answers (CManageData.cpp:81-90).
Everything is bit-exact: s16 maps, f64 maps (compared as raw bits), points.
"""
import os

import numpy as np
import pytest

from reconstruction_b200 import synth

NOMATCH = -10000


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.int64) if a.dtype == np.float64 else a


def _same(a, b):
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(_bits(a), _bits(b))


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "stereo_small.npz"))


def _port_from_gold(oracle, gold):
    w0, h0, L = (int(v) for v in gold["lowest"])
    o = oracle.CpuStereo("port", L, w0, h0, int(gold["origin"][0]), int(gold["origin"][1]))
    o.set_pair(gold["img0"], gold["img1"], gold["mask0"], gold["mask1"])
    o.set_calib(gold["Q"], gold["R_final"], gold["T_final"])
    return o, L


def test_synth_is_reproducible(gold):
    w0, h0, L = (int(v) for v in gold["lowest"])
    sp = synth.make_pair(w0, h0, L, pair_id=3)
    assert np.array_equal(sp.image[0], gold["img0"]) and np.array_equal(sp.image[1], gold["img1"])
    assert np.array_equal(sp.mask[0], gold["mask0"]) and np.array_equal(sp.mask[1], gold["mask1"])


def test_port_free_running_matches_reference_golden(oracle, gold):
    """Whole pair, stage by stage, no teacher forcing: every dump point of MatchOneLayer."""
    o, L = _port_from_gold(oracle, gold)
    for lv in range(L):
        for v in (0, 1):
            img, mask = o.get_level(lv, v)
            assert np.array_equal(img, gold[f"L{lv}_img{v}"]), (lv, v)
            assert np.array_equal(mask, gold[f"L{lv}_mask{v}"]), (lv, v)
        for st in range(1, 11):
            o.run_stage(lv, st)
            if st == 1:
                assert np.array_equal(o.get_margins(), gold[f"L{lv}_margins"])
                continue
            for d in (0, 1):
                assert _same(o.get_disparity(d, lv), gold[f"L{lv}_S{st}_d{d}"]), (lv, st, d)
            if st == 6:
                for d in (0, 1):
                    bl, br = o.get_rematch_bounds(d, lv)
                    assert np.array_equal(bl, gold[f"L{lv}_BL{d}"]) and np.array_equal(br, gold[f"L{lv}_BR{d}"])
    pts = o.to_cloud()
    assert _same(pts, gold["points"])
    bgr, pix = o.get_point_attrs()
    assert len(pix) == len(pts) and np.all(np.diff(pix) > 0)  # row-major emission order (Q11)
    top = gold["img0"].reshape(-1, 3)
    assert np.array_equal(bgr, top[pix])


def test_port_teacher_forced_stages(oracle, gold):
    """Each stage fed with the reference's previous-stage output."""
    o, L = _port_from_gold(oracle, gold)
    for lv in range(L):
        o.run_stage(lv, 1)
        for st in range(2, 11):
            if st == 2 and lv > 0:
                for d in (0, 1):
                    o.set_disparity(d, gold[f"L{lv-1}_S10_d{d}"])
            elif st > 2:
                for d in (0, 1):
                    o.set_disparity(d, gold[f"L{lv}_S{st-1}_d{d}"])
            o.run_stage(lv, st)
            for d in (0, 1):
                assert _same(o.get_disparity(d, lv), gold[f"L{lv}_S{st}_d{d}"]), (lv, st, d)


def test_port_matches_reference_live(oracle):
    if not oracle.available("ref"):
        pytest.skip("oracle/_ref not built here (no /root/reference)")
    L, w0, h0 = 3, 40, 32
    sp = synth.make_pair(w0, h0, L, pair_id=5, origin_scale=1.5)
    objs = []
    for kind in ("ref", "port"):
        o = oracle.CpuStereo(kind, L, w0, h0, *sp.origin_size)
        o.set_pair(*sp.image, *sp.mask)
        o.set_calib(sp.Q, sp.R_final, sp.T_final)
        objs.append(o)
    r, p = objs
    for lv in range(L):
        for st in range(1, 11):
            r.run_stage(lv, st)
            p.run_stage(lv, st)
            if st == 1:
                assert np.array_equal(r.get_margins(), p.get_margins())
                continue
            for d in (0, 1):
                assert _same(r.get_disparity(d, lv), p.get_disparity(d, lv)), (lv, st, d)
    assert _same(r.to_cloud(), p.to_cloud())


def test_match_pair_entry_equals_staged_run(oracle, gold):
    o, L = _port_from_gold(oracle, gold)
    n = o.match_pair()
    assert n == len(gold["points"])
    assert _same(o.get_disparity(0, L - 1), gold[f"L{L-1}_S10_d0"])


def test_pyrdown_against_cv2(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "pyrdown_cv2.npz"))
    n = len([k for k in g.files if k.startswith("src")])
    for i in range(n):
        assert np.array_equal(oracle.pyrdown("port", g[f"src{i}"]), g[f"dst{i}"]), i


def test_erode_against_cv2(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "erode_cv2.npz"))
    n = len([k for k in g.files if k.startswith("src")])
    for i in range(n):
        ks = int(g[f"ks{i}"])
        assert np.array_equal(oracle.structuring_ellipse("port", ks), g[f"kernel{i}"]), i
        assert np.array_equal(oracle.erode_ellipse("port", g[f"src{i}"], ks), g[f"dst{i}"]), i


def test_ncc_known_answers(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "ncc_kat.npz"))
    for (ws, y0, xl, xr), norm, vec, val in zip(g["cases"], g["norms"], g["vecs"], g["vals"]):
        n, v = oracle.window_to_vec("port", g["img_l"], int(y0), int(xl), int(ws))
        assert n == norm and np.array_equal(_bits(v), _bits(vec[: v.size]))
        assert oracle.ncc_match_value("port", g["img_l"], g["img_r"], int(y0), int(xl), int(xr), int(ws)) == val
    # flat window: norm forced to 1, vector 0, NCC exactly 0 (quirk Q10)
    n, v = oracle.window_to_vec("port", g["img_l"], 0, 0, 5)
    assert n == 1.0 and not v.any()
    assert oracle.ncc_match_value("port", g["img_l"], g["img_r"], 0, 0, 3, 5) == 0.0


def test_quirk_negative_truncation_and_median(oracle):
    """Q2 / Q7 known answers on hand-made maps (values from the algorithm as written)."""
    L, w0, h0 = 1, 24, 16
    img = np.random.default_rng(0).integers(0, 256, (h0, w0, 3), dtype=np.uint8)
    mask = np.zeros((h0, w0), np.uint8)
    mask[3:13, 3:21] = 255
    o = oracle.CpuStereo("port", L, w0, h0)
    o.set_pair(img, img, mask, mask)
    o.run_stage(0, 1)
    assert o.get_margins().tolist() == [[3, 12, 3, 20, 18, 10]] * 2
    d = np.full((h0, w0), NOMATCH, np.int16)
    d[5:8, 6:9] = np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9]], np.int16)
    o.set_disparity(0, d)
    o.set_disparity(1, d)
    o.run_stage(0, 8)
    m = o.get_disparity(0, 0)
    # window = rows y-1..y+1, columns x-1..x only (Q7): at (6,7): {1,2,4,5,7,8} -> even count -> 4+(5-4)/2 = 4
    assert m[6, 7] == 4
    # centre missing with >= 4 valid neighbours in the two columns gets filled: (6,9) sees {3,6,9} only -> stays NOMATCH
    assert m[6, 9] == NOMATCH
    assert m[6, 6] == NOMATCH or m[6, 6] == 4  # (6,6): {1,4,7} k=3 centre valid -> median 4
    assert m[6, 6] == 4
