"""Synthetic staged inputs for the stereo-matching hot path.

The parity boundary sits AFTER the reference's Rectify (CStereoMatching.cpp:117-168,
OpenCV arithmetic, un-pinned — SURVEY.md §8c), so what is generated here is exactly
what Rectify leaves behind for one camera pair: two rectified top-level BGR images,
two 0/255 masks, and Q / R_final / T_final.  Content follows SURVEY.md §8d: a
high-entropy multi-octave value-noise texture (three decorrelated channels, full
0..255 range) seen through a smooth, bumpy disparity field, masks a few pixels away
from the image border, everything seeded with 20260000 + pair id.

Nothing here is on the timed path; bench.py and the tests call it to make inputs.
"""
from __future__ import annotations

import dataclasses

import numpy as np


@dataclasses.dataclass
class StagedPair:
    """What Rectify hands to ConstructPyrm / MatchOneLayer / DisparityToCloud."""

    image: tuple  # (view0, view1) uint8 [H, W, 3] BGR, C-contiguous
    mask: tuple  # (view0, view1) uint8 [H, W]
    Q: np.ndarray  # 4x4 f64, AFTER the sign flip of Q[3][2] (CStereoMatching.cpp:138)
    R_final: np.ndarray  # 3x3 f64 (CStereoMatching.cpp:132)
    T_final: np.ndarray  # 3 f64   (CStereoMatching.cpp:133)
    origin_size: tuple  # (W_origin, H_origin): m_OriginSize, sets `scale` (CStereoMatching.cpp:692)
    pyrm_num: int
    lowest_size: tuple  # (W0, H0)

    @property
    def top_size(self):
        return (self.lowest_size[0] << (self.pyrm_num - 1), self.lowest_size[1] << (self.pyrm_num - 1))


def _bilinear_upsample(grid: np.ndarray, out_h: int, out_w: int, cell: float) -> np.ndarray:
    """grid [gh, gw, C] sampled at (y/cell, x/cell), bilinear, float32."""
    ys = np.arange(out_h, dtype=np.float32) / cell
    xs = np.arange(out_w, dtype=np.float32) / cell
    y0 = np.floor(ys).astype(np.int64)
    x0 = np.floor(xs).astype(np.int64)
    fy = (ys - y0)[:, None, None]
    fx = (xs - x0)[None, :, None]
    rows0 = grid[y0]
    rows1 = grid[y0 + 1]
    top = rows0[:, x0] * (1 - fx) + rows0[:, x0 + 1] * fx
    bot = rows1[:, x0] * (1 - fx) + rows1[:, x0 + 1] * fx
    return top * (1 - fy) + bot * fy


def value_noise(h: int, w: int, rng: np.random.Generator, octaves=(2, 4, 8, 16, 32, 64), channels=3) -> np.ndarray:
    """Multi-octave value noise, float32 [h, w, channels], roughly zero-mean unit-variance."""
    acc = np.zeros((h, w, channels), dtype=np.float32)
    for lam in octaves:
        gh, gw = int(h / lam) + 3, int(w / lam) + 3
        grid = rng.standard_normal((gh, gw, channels), dtype=np.float32)
        acc += _bilinear_upsample(grid, h, w, float(lam)) * np.float32(lam**0.35)
    acc -= acc.mean(axis=(0, 1), keepdims=True)
    acc /= acc.std(axis=(0, 1), keepdims=True) + 1e-6
    return acc


def synth_calibration(top_w: int, top_h: int, origin_scale: float = 1.0):
    """Q (post sign flip), R_final, T_final and the origin size of a plausible rectified rig."""
    wo, ho = int(round(top_w * origin_scale)), int(round(top_h * origin_scale))
    f = 0.9 * wo
    cx, cy = wo / 2 + 3.7, ho / 2 - 2.2
    cx1 = cx - 0.015 * wo
    tx = -120.0  # mm, stereoRectify's Tx for a left->right pair
    q = np.zeros((4, 4))
    q[0, 0] = q[1, 1] = 1.0
    q[0, 3] = -cx
    q[1, 3] = -cy
    q[2, 3] = f
    q[3, 2] = -(-1.0 / tx)  # stereoRectify gives -1/Tx; the reference flips the sign (:138)
    q[3, 3] = (cx - cx1) / tx
    ay, ax = np.deg2rad(12.0), np.deg2rad(-5.0)
    ry = np.array([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]])
    rx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
    r_final = ry @ rx
    t_final = np.array([210.5, -35.25, -980.125])
    return q, r_final, t_final, (wo, ho)


def make_pair(
    lowest_w: int,
    lowest_h: int,
    pyrm_num: int,
    pair_id: int = 0,
    origin_scale: float = 1.0,
    mask_hole: bool = True,
    noise_sigma: float = 1.5,
) -> StagedPair:
    """One synthetic rectified pair at the top-level size lowest * 2**(pyrm_num-1)."""
    W, H = lowest_w << (pyrm_num - 1), lowest_h << (pyrm_num - 1)
    rng = np.random.default_rng(20260000 + pair_id)

    d_mean, d_amp = -0.08 * W, 0.03 * W
    pad = int(abs(d_mean) + 2 * d_amp) + 16
    canvas = value_noise(H, W + pad + 8, rng)  # T(u, y)
    tex = np.clip(128.0 + 52.0 * canvas, 0, 255).astype(np.float32)

    yy, xx = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    nx, ny = (xx - W / 2) / (W / 2), (yy - H / 2) / (H / 2)
    r2 = nx * nx + ny * ny
    # right-grid disparity: a dome plus gentle ripples (|d/dx| << 1 so no fold-over)
    d_r = (
        d_mean
        - d_amp * np.sqrt(np.clip(1.2 - r2, 0.05, None))
        + 0.12 * d_amp * np.sin(5.1 * nx + 0.7 * pair_id) * np.cos(4.3 * ny)
    ).astype(np.float32)

    # left view: canvas column x ; right view: canvas column x' - d_r (linear interpolation in u)
    img0 = tex[:, :W, :]
    u = xx - d_r
    u0 = np.floor(u).astype(np.int64)
    fu = (u - u0)[..., None]
    u0 = np.clip(u0, 0, tex.shape[1] - 2)
    rows = np.arange(H)[:, None]
    img1 = tex[rows, u0] * (1 - fu) + tex[rows, u0 + 1] * fu
    img1 = 0.97 * img1 + 2.0
    if noise_sigma > 0:
        img0 = img0 + rng.standard_normal(img0.shape, dtype=np.float32) * noise_sigma
        img1 = img1 + rng.standard_normal(img1.shape, dtype=np.float32) * noise_sigma
    img0 = np.ascontiguousarray(np.clip(np.rint(img0), 0, 255).astype(np.uint8))
    img1 = np.ascontiguousarray(np.clip(np.rint(img1), 0, 255).astype(np.uint8))

    def silhouette(px, py):
        ex, ey = (px - W * 0.5) / (W * 0.43), (py - H * 0.5) / (H * 0.45)
        inside = ex * ex + ey * ey <= 1.0
        if mask_hole:
            hx, hy = (px - W * 0.62) / (W * 0.05), (py - H * 0.40) / (H * 0.04)
            inside &= hx * hx + hy * hy > 1.0
        return inside

    m0 = silhouette(xx, yy)
    m1 = silhouette(u, yy)  # the same surface region seen from view 1
    border = max(8, 2 * 3)
    for m in (m0, m1):
        m[:border, :] = False
        m[-border:, :] = False
        m[:, :border] = False
        m[:, -border:] = False
    mask0 = np.where(m0, 255, 0).astype(np.uint8)
    mask1 = np.where(m1, 255, 0).astype(np.uint8)

    q, r_final, t_final, origin = synth_calibration(W, H, origin_scale)
    return StagedPair(
        image=(img0, img1),
        mask=(np.ascontiguousarray(mask0), np.ascontiguousarray(mask1)),
        Q=q,
        R_final=r_final,
        T_final=t_final,
        origin_size=origin,
        pyrm_num=pyrm_num,
        lowest_size=(lowest_w, lowest_h),
    )


def make_torture_pair(lowest_w: int, lowest_h: int, pyrm_num: int, kind: str, seed: int = 0) -> StagedPair:
    """Adversarial staged pairs for the parity tests (small sizes, CPU oracle alongside):
      "flat"   two grey levels only (values 100 / 101, blocks of constant colour): NCC ties and zero-variance windows everywhere
               (quirk Q10, the screening pass must hand these to the exact search)
      "holes"  a mask riddled with holes and thin bridges: mode 1 / 2 refinement pixels, hole look-ahead ranges (Q3), wide Rematch ranges
      "steps"  piecewise-constant disparity with jumps of 6-40 px: refinement pixels leaving their table window, exp() arguments
               beyond -512, uniqueness / order constraint removals
      "sat"    heavily saturated texture (large areas clipped to 0 / 255) with fine detail elsewhere
    """
    W, H = lowest_w << (pyrm_num - 1), lowest_h << (pyrm_num - 1)
    rng = np.random.default_rng(777000 + seed)
    base = make_pair(lowest_w, lowest_h, pyrm_num, pair_id=100 + seed)
    img0, img1 = base.image[0].copy(), base.image[1].copy()
    m0, m1 = base.mask[0].copy(), base.mask[1].copy()
    if kind == "flat":
        blk = 6
        g0 = rng.integers(100, 102, (H // blk + 2, W // blk + 40, 3), dtype=np.uint8)
        canvas = np.repeat(np.repeat(g0, blk, axis=0), blk, axis=1)
        shift = int(0.08 * W)
        img0 = np.ascontiguousarray(canvas[:H, shift:shift + W])
        img1 = np.ascontiguousarray(canvas[:H, :W])  # constant disparity -shift
        m1 = np.roll(m0, -shift, axis=1)
        m1[:, -shift - 8:] = 0
    elif kind == "holes":
        holes = rng.random((H // 8 + 1, W // 8 + 1)) < 0.22
        hm = np.repeat(np.repeat(holes, 8, axis=0), 8, axis=1)[:H, :W]
        thin = (np.arange(W)[None, :] % 37 < 2) | (np.arange(H)[:, None] % 41 < 2)
        m0[hm | thin] = 0
        m1[np.roll(hm, -int(0.08 * W), axis=1)] = 0
    elif kind == "steps":
        canvas = np.clip(128 + 52 * value_noise(H, W + 200, rng), 0, 255).astype(np.uint8)
        img0 = np.ascontiguousarray(canvas[:, 100:100 + W])
        img1 = np.zeros_like(img0)
        band = H // 6
        for b, d in enumerate([0, 6, -14, 25, -40, 9, 0]):
            y0, y1 = b * band, min(H, (b + 1) * band)
            if y0 >= H:
                break
            img1[y0:y1] = canvas[y0:y1, 100 + d:100 + d + W]
        m1 = m0.copy()
    elif kind == "sat":
        t = value_noise(H, W + 64, rng)
        canvas = np.clip(128 + 400 * t, 0, 255).astype(np.uint8)
        shift = int(0.05 * W)
        img0 = np.ascontiguousarray(canvas[:, shift:shift + W])
        img1 = np.ascontiguousarray(canvas[:, :W])
        m1 = np.roll(m0, -shift, axis=1)
        m1[:, -shift - 8:] = 0
    else:
        raise ValueError(kind)
    border = 8
    for m in (m0, m1):
        m[:border], m[-border:], m[:, :border], m[:, -border:] = 0, 0, 0, 0
    return dataclasses.replace(base, image=(np.ascontiguousarray(img0), np.ascontiguousarray(img1)),
                               mask=(np.ascontiguousarray(m0), np.ascontiguousarray(m1)))
