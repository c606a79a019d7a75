"""Oracle parity at BASELINE sizes, with the production kernel variants.

The small-size tests (test_gpu_parity.py) never reach the tile variant that produces the headline
(`k_refine_fused<64,80,512,2>` + TMA is chosen for interiors of 400 000 px and more, refine.cu) nor the 120 / 150-sweep
levels of a 5-level pyramid (CStereoMatching.cpp:95).  Here every dump point of MatchOneLayer (CStereoMatching.cpp:63-111:
S2..S10, BL/BR) and the triangulated points are compared bit for bit against `oracle/_ref` — the reference's own
CStereoMatching.cpp + CManageData.cpp compiled unmodified — on

  * BASELINE configs[1]: one pair of the 2048x1536, 3-level rig, the whole pair;
  * a 5-level 1536x1152 pair (120 and 150 sweeps, the >= 400k-pixel tile + TMA path);
  * a 5-level pair whose row pitches are not multiples of 16 bytes (lowest 94x63: the non-TMA fall-back at size).

`oracle/_ref/libstereo_ref.so` travels to the GPU box prebuilt; when it is missing the restatement (`port`, pinned to
`_ref` in test_oracle_cpu.py) is used instead and the test says so.  One small pair is also run under every tile variant x
TMA on/off x screening on/off."""
import itertools
import os

import numpy as np
import pytest

from reconstruction_b200 import capi, synth

pytestmark = pytest.mark.gpu
NOMATCH = -10000
REL_TOL = 1e-4  # BASELINE.json north_star


@pytest.fixture(scope="module")
def lib():
    import torch

    assert torch.cuda.is_available(), "these tests need the B200"
    capi.build()
    return capi.load()


def _kind(oracle):
    if oracle.available("ref"):
        return "ref"
    print("oracle/_ref/libstereo_ref.so is absent: comparing against the restatement (port)")
    return "port"


def _same(a, b, what):
    assert a.shape == b.shape and a.dtype == b.dtype, what
    if a.dtype == np.float64:
        nm_a, nm_b = a == NOMATCH, b == NOMATCH
        assert np.array_equal(nm_a, nm_b), f"{what}: NOMATCH sets differ at {int((nm_a != nm_b).sum())} pixels"
        v = ~nm_a
        rel = np.abs(a[v] - b[v]) / np.maximum(np.abs(b[v]), 1.0)
        assert rel.size == 0 or rel.max() <= REL_TOL, f"{what}: max rel err {rel.max():.3e} over {int((rel > REL_TOL).sum())} px"
        a, b = a.view(np.int64), b.view(np.int64)
    diff = a != b
    assert not diff.any(), f"{what}: {int(diff.sum())} of {diff.size} elements differ (first at {np.argwhere(diff)[:4].tolist()})"
    return a.size


def _env_ctx(env, *args):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return capi.StereoB200(*args)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("name,L,w0,h0,pair_id", [("B 2048x1536 L3", 3, 512, 384, 3), ("1536x1152 L5", 5, 96, 72, 3),
                                                  ("odd 1504x1008 L5", 5, 94, 63, 5)])
def test_every_dump_point_at_size(lib, oracle, name, L, w0, h0, pair_id):
    kind = _kind(oracle)
    sp = synth.make_pair(w0, h0, L, pair_id=pair_id)
    o = oracle.CpuStereo(kind, L, w0, h0, *sp.origin_size)
    g = capi.StereoB200(L, w0, h0, *sp.origin_size)
    for e in (o, g):
        e.set_pair(*sp.image, *sp.mask)
        e.set_calib(sp.Q, sp.R_final, sp.T_final)
    compared = 0
    for lv in range(L):
        for st in range(1, 11):
            o.run_stage(lv, st)
            g.run_stage(lv, st)
            if st == 1:
                assert np.array_equal(o.get_margins(), g.get_margins(lv)), f"{name}: FindMargin level {lv}"
                continue
            for d in (0, 1):
                compared += _same(g.get_disparity(d), o.get_disparity(d, lv), f"{name} [{kind}]: level {lv} stage {st} ({capi.STAGE_NAMES[st]}) dir {d}")
            if st == 6:
                for d in (0, 1):
                    obl, obr = o.get_rematch_bounds(d, lv)
                    gbl, gbr = g.get_rematch_bounds(d, lv)
                    _same(gbl, obl, f"{name}: level {lv} BL dir {d}")
                    _same(gbr, obr, f"{name}: level {lv} BR dir {d}")
    oxyz = o.to_cloud()
    xyz, bgr, pix = g.to_cloud()
    assert len(xyz) == len(oxyz) and len(xyz) > 0
    _same(xyz, oxyz, f"{name}: points")
    assert np.all(np.diff(pix) > 0), "points are emitted in row-major order (Q11)"
    assert np.array_equal(bgr, sp.image[0].reshape(-1, 3)[pix])
    if kind == "port":
        obgr, opix = o.get_point_attrs()
        assert np.array_equal(pix, opix) and np.array_equal(bgr, obgr)
    fallbacks, misses = g.refine_counters()
    print(f"{name}: oracle {kind}, {compared} map elements and {len(xyz)} points compared bit for bit; {fallbacks} pixels went to the "
          f"exact NCC pass, {misses} out-of-window refinement evaluations")
    g.close()
    o.close()


@pytest.fixture(scope="module")
def small_case(oracle):
    L, w0, h0 = 2, 160, 120
    sp = synth.make_pair(w0, h0, L, pair_id=11)
    kind = "ref" if oracle.available("ref") else "port"
    o = oracle.CpuStereo(kind, L, w0, h0, *sp.origin_size)
    o.set_pair(*sp.image, *sp.mask)
    o.set_calib(sp.Q, sp.R_final, sp.T_final)
    for lv in range(L):
        o.match_one_layer(lv)
    d = [o.get_disparity(k, L - 1) for k in (0, 1)]
    xyz = o.to_cloud()
    return sp, L, w0, h0, d, xyz


@pytest.mark.parametrize("tile,tma,screen", list(itertools.product(range(8), (0, 1), (0, 1, 2))))
def test_every_tile_variant_against_the_oracle(lib, small_case, tile, tma, screen):
    """All 8 tile shapes of the fused refinement x TMA / plain tile loads x the NCC search path (0 = every candidate in exact
    arithmetic, 1 = integer screening with the TMA band kernel for K3, 2 = integer screening with the register-resident K3
    kernel of round 1): final maps and points against the oracle, bit for bit (not against each other)."""
    sp, L, w0, h0, d, oxyz = small_case
    g = _env_ctx({"SB200_REFINE_TILE": tile, "SB200_REFINE_TMA": tma, "SB200_SCREEN": int(screen > 0), "SB200_BAND": int(screen != 2)},
                 L, w0, h0, *sp.origin_size)
    g.set_pair(*sp.image, *sp.mask)
    g.set_calib(sp.Q, sp.R_final, sp.T_final)
    n = g.match_pair()
    for k in (0, 1):
        _same(g.get_disparity(k), d[k], f"tile {tile} tma {tma} screen {screen}: disparity[{k}]")
    xyz, _, _ = g.get_points(n)
    _same(xyz, oxyz, f"tile {tile} tma {tma} screen {screen}: points")
    g.close()
