"""Full-size GPU checks (BASELINE.json configs C and E).  The CPU oracle needs minutes per pair at these sizes, so the
CUDA path is checked through size-independent properties of the domain instead:
  * the pyramid and margins against the oracle's pyrDown / a numpy bounding box (cheap on the CPU at any size),
  * invariance of the refined maps under the fused kernel's tiling (sweeps per launch, tile shape) - bit for bit,
  * run-to-run determinism,
  * the uniqueness invariant after UniquenessContraint<double> (CStereoMatching.cpp:450-497),
  * DisparityToCloud (CStereoMatching.cpp:682-761): point set == eroded mask & matched pixels in row-major order, xyz recomputed from
    the disparity with the reference's formula (bit-exact), colours == source pixels.
Small-size bit parity against the oracle lives in test_gpu_parity.py."""
import math
import os

import numpy as np
import pytest

from reconstruction_b200 import capi, synth

pytestmark = pytest.mark.gpu
NOMATCH = -10000


@pytest.fixture(scope="module")
def lib():
    import torch

    assert torch.cuda.is_available(), "these tests need the B200"
    capi.build()
    return capi.load()


def _run(sp, L, w0, h0, env=None):
    old = {}
    for k, v in (env or {}).items():
        old[k] = os.environ.get(k)
        os.environ[k] = str(v)
    try:
        g = capi.StereoB200(L, w0, h0, *sp.origin_size)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    g.set_calib(sp.Q, sp.R_final, sp.T_final)
    g.set_pair(*sp.image, *sp.mask)
    n = g.match_pair()
    return g, n


def _check_uniqueness_invariant(p, q, ms, mt):
    """Every surviving p[x] has a partner q[x'] with |q + p| < 2 in [bl, br], or passes one of the two neighbour tests
    (:476-493).  Checked with the FINAL maps: the last pass (0 -> 1) leaves map 0 satisfying it against the final map 1
    except where the neighbour test used a p[x-1] that the same sweep killed afterwards - so only the partner/neighbour
    disjunction on surviving pixels is asserted, with killed neighbours counted as failing (the reference's own rule)."""
    H, W = p.shape
    YL, YR, XL, XR = ms[:4]
    XL1, XR1 = mt[2], mt[3]
    ys, xs = np.nonzero(p[YL:YR + 1, XL:XR + 1] != NOMATCH)
    ys += YL
    xs += XL
    pv = p[ys, xs]
    bl = np.maximum((pv + 0.5).astype(np.int64) + xs - 1, XL1)  # int() truncation, as the reference
    br = np.minimum(bl + 2, XR1)
    ok = np.zeros(len(pv), bool)
    for k in range(3):
        xx = bl + k
        inside = xx <= br
        qv = q[ys, np.clip(xx, 0, W - 1)]
        ok |= inside & (np.abs(qv + pv) < 2)
    qm = q[ys, np.clip(bl + 1, 0, W - 1)]
    ok |= np.abs(qm + p[ys, np.clip(xs - 1, 0, W - 1)]) < 2
    pn = p[ys, np.clip(xs + 1, 0, W - 1)]
    ok |= np.abs(qm + pn) < 2
    ok |= pn == NOMATCH  # p[x+1] was tested with its value BEFORE the sweep reached it; if it died afterwards that value is gone
    return int((~ok).sum()), len(pv)


def _check_cloud(g, sp, oracle, L):
    W, H = sp.top_size
    d0 = g.get_disparity(0)
    xyz, bgr, pix = g.to_cloud()
    m = g.get_margins(L - 1)[0]
    ks = int(math.ceil(0.02 * H))
    er = oracle.erode_ellipse("port", sp.mask[0], ks)
    sel = np.zeros((H, W), bool)
    sel[m[0]:m[1] + 1, m[2]:m[3] + 1] = True
    sel &= (er == 255) & (d0 != NOMATCH)
    exp_pix = np.flatnonzero(sel.reshape(-1)).astype(np.int32)  # row-major == the reference's emission order
    assert np.array_equal(pix, exp_pix), "point set / order differs from eroded-mask & matched pixels"
    assert np.array_equal(bgr, sp.image[0].reshape(-1, 3)[pix])
    scale = float(g.lowest[0]) / sp.origin_size[0] * (1 << (L - 1))
    q = sp.Q.copy()
    q[:, 3] *= scale
    y, x = np.divmod(pix.astype(np.int64), W)
    d = d0.reshape(-1)[pix]
    iw = 1.0 / (q[3, 3] + q[3, 2] * d)
    f0 = (q[0, 3] + x.astype(np.float64)) * iw
    f1 = (y.astype(np.float64) + q[1, 3]) * iw
    f2 = q[2, 3] * iw
    R, T = sp.R_final, sp.T_final
    for r in range(3):  # cv::gemm row: ((R0*F0 + R1*F1) + R2*F2) + T, no FMA
        ref = ((R[r, 0] * f0 + R[r, 1] * f1) + R[r, 2] * f2) + T[r]
        rel = np.abs(xyz[:, r] - ref) / np.maximum(np.abs(ref), 1.0)
        assert rel.max() <= 1e-12, f"xyz[{r}] deviates from the reference formula: {rel.max():.3e}"
    return len(pix)


@pytest.mark.parametrize("name,L,w0,h0", [("C", 5, 256, 192), ("E", 5, 375, 250)])
def test_full_size_properties(lib, oracle, name, L, w0, h0):
    sp = synth.make_pair(w0, h0, L, pair_id=3)
    g, n = _run(sp, L, w0, h0)
    W, H = sp.top_size
    # pyramid + margins
    img, mask = sp.image[0], sp.mask[0]
    for lv in range(L - 1, -1, -1):
        gi, gm = g.get_level(lv, 0)
        assert np.array_equal(gi, img) and np.array_equal(gm, mask), f"pyramid level {lv}"
        ys, xs = np.nonzero(mask[2:-2, 2:-2] == 255)
        assert g.get_margins(lv)[0][:4].tolist() == [ys.min() + 2, ys.max() + 2, xs.min() + 2, xs.max() + 2]
        if lv:
            img, mask = oracle.pyrdown("port", img), oracle.pyrdown("port", mask)
    d = [g.get_disparity(k) for k in (0, 1)]
    assert n > 0.3 * W * H and np.isfinite(d[0]).all() and np.isfinite(d[1]).all()
    # uniqueness invariant on both maps
    mg = g.get_margins(L - 1)
    bad0, tot0 = _check_uniqueness_invariant(d[0], d[1], mg[0], mg[1])
    assert tot0 > 0 and bad0 == 0, f"{bad0} of {tot0} surviving pixels of map 0 violate the uniqueness rule"
    assert _check_cloud(g, sp, oracle, L) == n
    # determinism + invariance under the fused refinement's tiling
    # ... and under the alternative code paths: no integer screening, the register-resident K3 kernel instead of the TMA band
    # kernel, kernel-by-kernel enqueueing instead of the CUDA graph
    for env in ({}, {"SB200_REFINE_T": 3, "SB200_REFINE_TILE": 0}, {"SB200_REFINE_T": 6, "SB200_REFINE_TILE": 2}, {"SB200_SCREEN": 0},
                {"SB200_BAND": 0}, {"SB200_GRAPH": 0}):
        g2, n2 = _run(sp, L, w0, h0, env)
        assert n2 == n
        for k in (0, 1):
            assert np.array_equal(g2.get_disparity(k).view(np.int64), d[k].view(np.int64)), f"map {k} changes under {env}"
        g2.close()
    g.close()
