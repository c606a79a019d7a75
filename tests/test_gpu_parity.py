"""GPU parity tests (run on the B200): the CUDA path, called through the C ABI, against
  (1) the committed golden fixtures produced by the reference's own sources, and
  (2) the CPU oracle run live on the same seeded inputs,
free-running and teacher-forced, at every dump point of MatchOneLayer
(CStereoMatching.cpp:63-111).  Bars: integer maps (s16 disparity, BL/BR, margins, pixel indices,
colours) bit-exact; f64 maps and 3-D points are compared bit-for-bit as well — the refinement is
chaotic at the ulp level, so the kernels reproduce the reference's operation order and a
tolerance of 0 is the honest bar; the north-star tolerance (1e-4 relative) is asserted separately so
a failure report says which bar broke."""
import os

import numpy as np
import pytest

from reconstruction_b200 import capi, synth

pytestmark = pytest.mark.gpu
NOMATCH = -10000
REL_TOL = 1e-4  # BASELINE.json north_star: <= 1e-4 relative on depth / disparity


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.int64) if a.dtype == np.float64 else a


def _check(a, b, what):
    assert a.shape == b.shape and a.dtype == b.dtype, what
    if a.dtype == np.float64:
        nm_a, nm_b = a == NOMATCH, b == NOMATCH
        assert np.array_equal(nm_a, nm_b), f"{what}: NOMATCH sets differ at {int((nm_a != nm_b).sum())} pixels"
        v = ~nm_a
        rel = np.abs(a[v] - b[v]) / np.maximum(np.abs(b[v]), 1.0)
        assert rel.size == 0 or rel.max() <= REL_TOL, f"{what}: max rel err {rel.max():.3e} over {int((rel > REL_TOL).sum())} px"
    diff = _bits(a) != _bits(b)
    assert not diff.any(), f"{what}: {int(diff.sum())} of {diff.size} elements differ (first at {np.argwhere(diff)[:4].tolist()})"


@pytest.fixture(scope="module")
def lib():
    import torch

    assert torch.cuda.is_available(), "these tests need the B200"
    capi.build()
    return capi.load()


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "stereo_small.npz"))


def _gpu_from_gold(gold):
    w0, h0, L = (int(v) for v in gold["lowest"])
    g = capi.StereoB200(L, w0, h0, int(gold["origin"][0]), int(gold["origin"][1]))
    g.set_pair(gold["img0"], gold["img1"], gold["mask0"], gold["mask1"])
    g.set_calib(gold["Q"], gold["R_final"], gold["T_final"])
    return g, L


def test_golden_free_running(lib, gold):
    g, L = _gpu_from_gold(gold)
    for lv in range(L):
        for v in (0, 1):
            img, mask = g.get_level(lv, v)
            assert np.array_equal(img, gold[f"L{lv}_img{v}"]), ("pyrDown image", lv, v)
            assert np.array_equal(mask, gold[f"L{lv}_mask{v}"]), ("pyrDown mask", lv, v)
        assert np.array_equal(g.get_margins(lv), gold[f"L{lv}_margins"]), ("FindMargin", lv)
        for st in range(1, 11):
            g.run_stage(lv, st)
            if st == 1:
                continue
            for d in (0, 1):
                _check(g.get_disparity(d), gold[f"L{lv}_S{st}_d{d}"], f"level {lv} stage {st} ({capi.STAGE_NAMES[st]}) dir {d}")
            if st == 6:
                for d in (0, 1):
                    bl, br = g.get_rematch_bounds(d, lv)
                    _check(bl, gold[f"L{lv}_BL{d}"], f"level {lv} BL dir {d}")
                    _check(br, gold[f"L{lv}_BR{d}"], f"level {lv} BR dir {d}")
    xyz, bgr, pix = g.to_cloud()
    _check(xyz, gold["points"], "points")
    assert np.all(np.diff(pix) > 0)
    assert np.array_equal(bgr, gold["img0"].reshape(-1, 3)[pix])


def test_golden_teacher_forced(lib, gold):
    g, L = _gpu_from_gold(gold)
    for lv in range(L):
        g.run_stage(lv, 1)
        for st in range(2, 11):
            if st == 2 and lv > 0:
                for d in (0, 1):
                    g.set_disparity(d, gold[f"L{lv-1}_S10_d{d}"])
            elif st > 2:
                for d in (0, 1):
                    g.set_disparity(d, gold[f"L{lv}_S{st-1}_d{d}"])
            g.run_stage(lv, st)
            for d in (0, 1):
                _check(g.get_disparity(d), gold[f"L{lv}_S{st}_d{d}"], f"teacher-forced level {lv} stage {st} dir {d}")


@pytest.mark.parametrize("L,w0,h0,pair_id,scale", [(3, 64, 48, 1, 1.0), (2, 160, 120, 2, 1.5), (1, 200, 150, 4, 1.0),
                                                   (2, 75, 51, 5, 1.0), (3, 50, 37, 7, 1.25)])  # odd widths: unaligned row pitches
def test_live_oracle_free_running(lib, oracle, L, w0, h0, pair_id, scale):
    sp = synth.make_pair(w0, h0, L, pair_id=pair_id, origin_scale=scale)
    o = oracle.CpuStereo("port", L, w0, h0, *sp.origin_size)
    g = capi.StereoB200(L, w0, h0, *sp.origin_size)
    for e in (o, g):
        e.set_pair(*sp.image, *sp.mask)
        e.set_calib(sp.Q, sp.R_final, sp.T_final)
    for lv in range(L):
        for st in range(1, 11):
            o.run_stage(lv, st)
            g.run_stage(lv, st)
            if st == 1:
                assert np.array_equal(o.get_margins(), g.get_margins(lv))
                continue
            for d in (0, 1):
                _check(g.get_disparity(d), o.get_disparity(d, lv), f"level {lv} stage {st} ({capi.STAGE_NAMES[st]}) dir {d}")
            if st == 6:
                for d in (0, 1):
                    obl, obr = o.get_rematch_bounds(d, lv)
                    gbl, gbr = g.get_rematch_bounds(d, lv)
                    _check(gbl, obl, f"level {lv} BL dir {d}")
                    _check(gbr, obr, f"level {lv} BR dir {d}")
    oxyz = o.to_cloud()
    obgr, opix = o.get_point_attrs()
    xyz, bgr, pix = g.to_cloud()
    assert len(xyz) == len(oxyz)
    assert np.array_equal(pix, opix) and np.array_equal(bgr, obgr)
    _check(xyz, oxyz, "points")
    assert g.refine_counters()[1] >= 0


@pytest.mark.parametrize("kind,L,w0,h0", [("flat", 2, 96, 72), ("holes", 3, 64, 48), ("steps", 3, 64, 48), ("sat", 2, 120, 90), ("holes", 2, 75, 51)])
def test_torture_inputs_free_running(lib, oracle, kind, L, w0, h0):
    """Adversarial inputs (NCC ties and flat windows, holes, disparity jumps, saturation): every dump point of every level,
    free-running, bit for bit; the screening pass / table misses / generic refinement paths these inputs force are counted."""
    sp = synth.make_torture_pair(w0, h0, L, kind)
    o = oracle.CpuStereo("port", L, w0, h0, *sp.origin_size)
    g = capi.StereoB200(L, w0, h0, *sp.origin_size)
    for e in (o, g):
        e.set_pair(*sp.image, *sp.mask)
        e.set_calib(sp.Q, sp.R_final, sp.T_final)
    g.refine_counters(reset=True)
    for lv in range(L):
        for st in range(1, 11):
            try:
                o.run_stage(lv, st)
            except Exception:  # the reference exit(0)s on a degenerate margin; the ABI reports it
                with pytest.raises(capi.StereoError):
                    g.run_stage(lv, st)
                return
            g.run_stage(lv, st)
            if st == 1:
                assert np.array_equal(o.get_margins(), g.get_margins(lv))
                continue
            for d in (0, 1):
                _check(g.get_disparity(d), o.get_disparity(d, lv), f"{kind}: level {lv} stage {st} ({capi.STAGE_NAMES[st]}) dir {d}")
    xyz, bgr, pix = g.to_cloud()
    oxyz = o.to_cloud()
    assert len(xyz) == len(oxyz)
    if len(xyz):
        _check(xyz, oxyz, f"{kind}: points")
    fallbacks, misses = g.refine_counters()
    print(f"{kind}: {len(xyz)} points, {fallbacks} pixels left to the exact NCC pass, {misses} out-of-window refinement evaluations")
    if kind == "flat":
        assert fallbacks > 0  # ties must not be settled by the screening pass


@pytest.mark.parametrize("iters", [0, 1, 7, 13])
def test_refine_iteration_override(lib, oracle, iters):
    """sweep counts that are not a multiple of the sweeps fused per launch (remainder launch), including none at all"""
    L, w0, h0 = 2, 96, 72
    sp = synth.make_pair(w0, h0, L, pair_id=8)
    o = oracle.CpuStereo("port", L, w0, h0)
    g = capi.StereoB200(L, w0, h0)
    for e in (o, g):
        e.set_pair(*sp.image, *sp.mask)
        e.set_calib(sp.Q, sp.R_final, sp.T_final)
        e.set_refine_iters(iters)
        for lv in range(L):
            e.match_one_layer(lv)
    for d in (0, 1):
        _check(g.get_disparity(d), o.get_disparity(d, L - 1), f"{iters} sweeps, dir {d}")


def test_match_pair_one_call(lib, oracle):
    """sb200_match_pair_host (what the C++ mirror calls) == staged run == oracle."""
    L, w0, h0 = 3, 80, 60
    sp = synth.make_pair(w0, h0, L, pair_id=6)
    o = oracle.CpuStereo("port", L, w0, h0)
    o.set_pair(*sp.image, *sp.mask)
    o.set_calib(sp.Q, sp.R_final, sp.T_final)
    n_ref = o.match_pair()
    oxyz = np.empty((n_ref, 3))
    o._f("get_points", None, [__import__("ctypes").c_void_p] * 2)(o.h, oxyz.ctypes.data_as(__import__("ctypes").c_void_p))
    g = capi.StereoB200(L, w0, h0)
    cap = sp.top_size[0] * sp.top_size[1]
    xyz = np.empty((cap, 3))
    bgr = np.empty((cap, 3), np.uint8)
    pix = np.empty(cap, np.int32)
    n = g.match_pair_host(*sp.image, *sp.mask, sp.Q, sp.R_final, sp.T_final, xyz, bgr, pix, cap)
    assert n == n_ref
    _check(xyz[:n], oxyz, "points via match_pair_host")
    _check(g.get_disparity(0), o.get_disparity(0, L - 1), "final disparity[0]")
    _check(g.get_disparity(1), o.get_disparity(1, L - 1), "final disparity[1]")
    assert g.launch_count() > 0


def test_error_paths(lib):
    g = capi.StereoB200(2, 48, 36)
    with pytest.raises(capi.StereoError):
        g.run_stage(0, 2)  # nothing uploaded
    img = np.zeros((72, 96, 3), np.uint8)
    empty = np.zeros((72, 96), np.uint8)
    g.set_pair(img, img, empty, empty)  # empty masks: inverted margins (:1014-1017)
    m = g.get_margins(1)
    assert m[0].tolist() == [72 - 1 - 2, 2, 96 - 1 - 2, 2, 2 - (96 - 1 - 2) + 1, 2 - (72 - 1 - 2) + 1]
    for st in range(1, 6):
        g.run_stage(0, st)
    with pytest.raises(capi.StereoError):  # the reference exit(0)s in SetBoundary_smooth
        g.run_stage(0, 6)


def test_pairs_in_flight_are_independent(lib):
    """Several contexts on one GPU, one host thread each (what bench.py and the C++ mirror do): every pair's
    disparity maps and points equal, bit for bit, those of the same pair processed alone."""
    import threading

    L, w0, h0 = 3, 96, 72
    pairs = [synth.make_pair(w0, h0, L, pair_id=20 + k) for k in range(3)]
    cap = pairs[0].top_size[0] * pairs[0].top_size[1]

    def run(g, sp, out):
        xyz = np.empty((cap, 3))
        pix = np.empty(cap, np.int32)
        for _ in range(3):  # keep the contexts overlapping for a while
            n = g.match_pair_host(*sp.image, *sp.mask, sp.Q, sp.R_final, sp.T_final, xyz, None, pix, cap)
        out.append((n, xyz[:n].copy(), pix[:n].copy(), g.get_disparity(0).copy(), g.get_disparity(1).copy()))

    alone = []
    g = capi.StereoB200(L, w0, h0)
    for sp in pairs:
        run(g, sp, alone)
    g.close()
    ctxs = [capi.StereoB200(L, w0, h0) for _ in pairs]
    outs = [[] for _ in pairs]
    th = [threading.Thread(target=run, args=(ctxs[k], pairs[k], outs[k])) for k in range(3)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for k in range(3):
        assert outs[k], f"context {k} failed"
        n, xyz, pix, d0, d1 = outs[k][0]
        rn, rxyz, rpix, rd0, rd1 = alone[k]
        assert n == rn and n > 0
        _check(xyz, rxyz, f"pair {k}: points")
        assert np.array_equal(pix, rpix)
        _check(d0, rd0, f"pair {k}: disparity[0]")
        _check(d1, rd1, f"pair {k}: disparity[1]")
    for c in ctxs:
        c.close()
