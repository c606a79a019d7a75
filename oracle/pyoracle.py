"""TEST INFRASTRUCTURE — ctypes front-end for the two CPU checkers.

* kind="ref"  -> oracle/_ref/libstereo_ref.so : the reference's own CStereoMatching.cpp +
                 CManageData.cpp compiled unmodified (oracle/ref_build/ref_harness.cpp).
* kind="port" -> oracle/libstereo_oracle.so   : the restatement (oracle/stereo_oracle.cpp).

Both export the same C functions with a different prefix (ref_ / orc_), so one class
drives either.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {"ref": os.path.join(HERE, "_ref", "libstereo_ref.so"), "port": os.path.join(HERE, "libstereo_oracle.so")}
_PREFIX = {"ref": "ref_", "port": "orc_"}

STAGE_NAMES = {
    1: "FindMargin",
    2: "InitialMatch",
    3: "SmoothConstraint",
    4: "OrderConstraint",
    5: "Uniqueness<short>#1",
    6: "Rematch",
    7: "Uniqueness<short>#2",
    8: "MedianFilter",
    9: "DisparityRefine",
    10: "Uniqueness<double>",
}


def build(kind: str | None = None) -> None:
    """Run the oracle Makefile (ref is skipped by make itself when /root/reference is absent)."""
    targets = ["port", "ref"] if kind is None else [kind]
    # make's chatter goes to stderr: bench.py's stdout carries exactly one JSON line
    subprocess.run(["make", "-s", "-C", HERE] + targets, check=True, stdout=sys.stderr)


def available(kind: str) -> bool:
    return os.path.exists(_LIBS[kind])


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class CpuStereo:
    """One camera pair through the CPU checker, stage by stage."""

    def __init__(self, kind, pyrm_num, lowest_w, lowest_h, origin_w=None, origin_h=None, radius=2, ws=0.03, offset=2):
        if not available(kind):
            raise FileNotFoundError(f"{_LIBS[kind]} missing — run `make -C oracle {kind}`")
        self.kind = kind
        self.lib = C.CDLL(_LIBS[kind])
        self.pre = _PREFIX[kind]
        self.L = pyrm_num
        self.lowest = (lowest_w, lowest_h)
        top_w, top_h = lowest_w << (pyrm_num - 1), lowest_h << (pyrm_num - 1)
        f = self._f("create", C.c_void_p, [C.c_int] * 6 + [C.c_double, C.c_int])
        self.h = C.c_void_p(f(pyrm_num, lowest_w, lowest_h, origin_w or top_w, origin_h or top_h, radius, ws, offset))

    def _f(self, name, restype, argtypes):
        fn = getattr(self.lib, self.pre + name)
        fn.restype = restype
        fn.argtypes = argtypes
        return fn

    def close(self):
        if self.h:
            self._f("destroy", None, [C.c_void_p])(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- configuration ---------------------------------------------------
    def set_threads(self, n):
        self._f("set_threads", None, [C.c_int])(n)

    def max_threads(self):
        return self._f("max_threads", C.c_int, [])()

    def level_size(self, level):
        return (self.lowest[0] << level, self.lowest[1] << level)

    def set_pair(self, img0, img1, mask0, mask1):
        arrs = [np.ascontiguousarray(a, dtype=np.uint8) for a in (img0, img1, mask0, mask1)]
        self._keep = arrs
        self._f("set_pair", None, [C.c_void_p] * 5)(self.h, *[_p(a) for a in arrs])

    def set_calib(self, Q, R, T):
        q, r, t = (np.ascontiguousarray(a, dtype=np.float64) for a in (Q, R, T))
        self._f("set_calib", None, [C.c_void_p] * 4)(self.h, _p(q), _p(r), _p(t))

    def set_refine_iters(self, n):
        self._f("set_refine_iters", None, [C.c_void_p, C.c_int])(self.h, n)

    # -- pyramid / margins -----------------------------------------------
    def get_level(self, level, view):
        w, h = self.level_size(level)
        img = np.empty((h, w, 3), np.uint8)
        mask = np.empty((h, w), np.uint8)
        self._f("get_level", None, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p])(
            self.h, level, view, _p(img), _p(mask)
        )
        return img, mask

    def get_margins(self):
        out = np.zeros(12, np.int32)
        self._f("get_margins", None, [C.c_void_p, C.c_void_p])(self.h, _p(out))
        return out.reshape(2, 6)

    # -- stages ------------------------------------------------------------
    def run_stage(self, level, stage):
        rc = self._f("run_stage", C.c_int, [C.c_void_p, C.c_int, C.c_int])(self.h, level, stage)
        if rc != 0:
            raise RuntimeError(f"{self.kind}: run_stage({level},{stage}) -> {rc}")

    def match_one_layer(self, level):
        self._f("match_one_layer", None, [C.c_void_p, C.c_int])(self.h, level)

    def match_pair(self):
        return int(self._f("match_pair", C.c_long, [C.c_void_p])(self.h))

    def get_disparity(self, dir_, level):
        es = self._f("disp_elem_size", C.c_int, [C.c_void_p, C.c_int])(self.h, dir_)
        if es == 0:
            return None
        w, h = self.level_size(level)
        out = np.empty((h, w), np.int16 if es == 2 else np.float64)
        self._f("get_disparity", None, [C.c_void_p, C.c_int, C.c_void_p])(self.h, dir_, _p(out))
        return out

    def set_disparity(self, dir_, arr):
        a = np.ascontiguousarray(arr)
        assert a.dtype in (np.int16, np.float64)
        self._f("set_disparity", None, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int])(
            self.h, dir_, _p(a), a.shape[0], a.shape[1], a.dtype.itemsize
        )

    def get_rematch_bounds(self, dir_, level):
        w, h = self.level_size(level)
        bl = np.empty((h, w), np.int16)
        br = np.empty((h, w), np.int16)
        self._f("get_rematch_bounds", None, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p])(self.h, dir_, _p(bl), _p(br))
        return bl, br

    # -- triangulation ------------------------------------------------------
    def to_cloud(self):
        n = int(self._f("to_cloud", C.c_long, [C.c_void_p])(self.h))
        xyz = np.empty((n, 3), np.float64)
        if n:
            self._f("get_points", None, [C.c_void_p, C.c_void_p])(self.h, _p(xyz))
        return xyz

    def get_point_attrs(self):
        """port only: per-point BGR and flat pixel index (y*W+x) in emission order."""
        assert self.kind == "port"
        n = int(self._f("num_points", C.c_long, [C.c_void_p])(self.h))
        bgr = np.empty((n, 3), np.uint8)
        pix = np.empty(n, np.int32)
        if n:
            self._f("get_point_attrs", None, [C.c_void_p, C.c_void_p, C.c_void_p])(self.h, _p(bgr), _p(pix))
        return bgr, pix


# ---- primitives (known-answer tests) ----------------------------------------
def _lib(kind):
    return C.CDLL(_LIBS[kind]), _PREFIX[kind]


def window_to_vec(kind, img, y0, x, ws):
    lib, pre = _lib(kind)
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty(ws * ws * 3, np.float64)
    f = getattr(lib, pre + "window_to_vec")
    f.restype = C.c_double
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    n = f(_p(img), img.shape[1] * 3, y0, x, ws, _p(out))
    return n, out


def ncc_match_value(kind, img_l, img_r, y0, xl, xr, ws):
    lib, pre = _lib(kind)
    a = np.ascontiguousarray(img_l, np.uint8)
    b = np.ascontiguousarray(img_r, np.uint8)
    f = getattr(lib, pre + "ncc_match_value")
    f.restype = C.c_double
    f.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 5
    return f(_p(a), _p(b), a.shape[1] * 3, y0, xl, xr, ws)


def pyrdown(kind, src):
    lib, pre = _lib(kind)
    src = np.ascontiguousarray(src, np.uint8)
    h, w = src.shape[:2]
    cn = 1 if src.ndim == 2 else src.shape[2]
    dst = np.empty(((h + 1) // 2, (w + 1) // 2) + (() if cn == 1 else (cn,)), np.uint8)
    f = getattr(lib, pre + "pyrdown")
    f.restype = None
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    f(_p(src), w, h, cn, _p(dst))
    return dst


def erode_ellipse(kind, src, ksize):
    lib, pre = _lib(kind)
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty_like(src)
    f = getattr(lib, pre + "erode_ellipse")
    f.restype = None
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    f(_p(src), src.shape[1], src.shape[0], ksize, _p(dst))
    return dst


def structuring_ellipse(kind, ksize):
    lib, pre = _lib(kind)
    dst = np.empty((ksize, ksize), np.uint8)
    f = getattr(lib, pre + "structuring_ellipse")
    f.restype = None
    f.argtypes = [C.c_int, C.c_void_p]
    f(ksize, _p(dst))
    return dst
