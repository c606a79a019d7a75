"""ctypes binding of the C ABI in include/stereo_b200.h (libstereo_b200.so).

Python here is plumbing for tests and bench.py: host buffers in, host buffers out, exactly what the
C++ mirror classes (reconstruction_b200/host) pass through the same ABI.  There is no fallback:
if the library is missing, or no B200 is present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libstereo_b200.so")
CSRC = os.path.join(HERE, "csrc")

NOMATCH = -10000
STAGE_NAMES = {
    1: "FindMargin", 2: "InitialMatch", 3: "SmoothConstraint", 4: "OrderConstraint", 5: "Uniqueness<short>#1",
    6: "Rematch", 7: "Uniqueness<short>#2", 8: "MedianFilter", 9: "DisparityRefine", 10: "Uniqueness<double>",
}

# every symbol include/stereo_b200.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = [
    "sb200_ctx_create", "sb200_ctx_destroy", "sb200_last_error", "sb200_status_string", "sb200_pair_upload",
    "sb200_pair_stage_device", "sb200_pair_set_calib", "sb200_match_pair", "sb200_match_one_layer", "sb200_run_stage", "sb200_set_refine_iters",
    "sb200_disparity_info", "sb200_get_disparity", "sb200_set_disparity", "sb200_get_rematch_bounds", "sb200_get_level",
    "sb200_get_margin", "sb200_triangulate", "sb200_get_points", "sb200_points_device", "sb200_match_pair_host",
    "sb200_stream", "sb200_launch_count", "sb200_graph_info", "sb200_set_profiling", "sb200_get_stage_ms", "sb200_get_refine_counters",
    "sb200_get_refine_profile", "sb200_get_stage_level_ms", "sb200_get_search_counters",
    "sb200_exp_host",
    "sb200_rectify_calib", "sb200_rectify_calib_compat", "sb200_stereo_rectify_host", "sb200_rectify_view", "sb200_pair_build", "sb200_get_rectify_maps",
    "sb200_set_rectify_maps", "sb200_get_remapped_mask",
    "sb200_sink_filter", "sb200_sink_last_error",
    "sb200_comm_unique_id", "sb200_comm_init", "sb200_comm_destroy", "sb200_comm_last_error", "sb200_allgather_points",
    "sb200_exchange_submit", "sb200_exchange_wait", "sb200_exchange_device", "sb200_exchange_drain", "sb200_comm_stats", "sb200_comm_set_consumer",
]
UNIQUE_ID_BYTES = 128


class StereoError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile the CUDA library in tree (nvcc, sm_100a); no-op when up to date."""
    r = subprocess.run(["make", "-s", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        raise StereoError("building libstereo_b200.so failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)
    return LIB_PATH


_lib = None


def load():
    """dlopen the library and declare the prototypes; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise StereoError(f"{LIB_PATH} is missing - run `make -C reconstruction_b200/csrc` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    P = C.POINTER
    protos = {
        "sb200_ctx_create": (i32, [P(vp), i32, i32, i32, i32, i32, i32, i32, dbl, i32]),
        "sb200_ctx_destroy": (None, [vp]),
        "sb200_last_error": (C.c_char_p, [vp]),
        "sb200_status_string": (C.c_char_p, [i32]),
        "sb200_pair_upload": (i32, [vp, vp, vp, vp, vp]),
        "sb200_pair_stage_device": (i32, [vp, vp, vp, vp, vp]),
        "sb200_pair_set_calib": (i32, [vp, vp, vp, vp]),
        "sb200_match_pair": (i32, [vp, P(i64)]),
        "sb200_match_one_layer": (i32, [vp, i32]),
        "sb200_run_stage": (i32, [vp, i32, i32]),
        "sb200_set_refine_iters": (i32, [vp, i32]),
        "sb200_disparity_info": (i32, [vp, P(i32), P(i32), P(i32)]),
        "sb200_get_disparity": (i32, [vp, i32, vp]),
        "sb200_set_disparity": (i32, [vp, i32, vp, i32, i32, i32]),
        "sb200_get_rematch_bounds": (i32, [vp, i32, vp, vp]),
        "sb200_get_level": (i32, [vp, i32, i32, vp, vp]),
        "sb200_get_margin": (i32, [vp, i32, i32, vp]),
        "sb200_triangulate": (i32, [vp, P(i64)]),
        "sb200_get_points": (i32, [vp, vp, vp, vp]),
        "sb200_points_device": (i32, [vp, P(vp), P(vp), P(vp), P(i64)]),
        "sb200_match_pair_host": (i32, [vp] * 11 + [i64, P(i64)]),
        "sb200_stream": (vp, [vp]),
        "sb200_launch_count": (i64, [vp]),
        "sb200_graph_info": (i32, [vp, P(i64), P(i64)]),
        "sb200_set_profiling": (i32, [vp, i32]),
        "sb200_get_stage_ms": (i32, [vp, vp, i32]),
        "sb200_get_refine_counters": (i32, [vp, vp, i32]),
        "sb200_get_search_counters": (i32, [vp, vp, i32]),
        "sb200_get_refine_profile": (i32, [vp, i32, P(dbl), P(i64), P(i64), i32]),
        "sb200_get_stage_level_ms": (i32, [vp, i32, i32, P(dbl), i32]),
        "sb200_exp_host": (dbl, [dbl]),
        "sb200_rectify_calib": (i32, [vp] * 4 + [i32] * 4 + [vp] * 6),
        "sb200_rectify_calib_compat": (i32, [vp] * 4 + [i32] * 5 + [vp] * 6),
        "sb200_stereo_rectify_host": (i32, [vp, vp, i32, i32, vp, vp, i32] + [vp] * 5),
        "sb200_rectify_view": (i32, [vp, i32, vp, vp, i32, i32, vp, vp, vp, i32]),
        "sb200_pair_build": (i32, [vp]),
        "sb200_get_rectify_maps": (i32, [vp, vp, vp]),
        "sb200_set_rectify_maps": (i32, [vp, vp, vp]),
        "sb200_get_remapped_mask": (i32, [vp, vp]),
        "sb200_sink_filter": (i32, [i32, vp, i64, i32, dbl, dbl, vp, vp, vp, i64, P(i64), vp]),
        "sb200_sink_last_error": (C.c_char_p, []),
        "sb200_comm_unique_id": (i32, [vp]),
        "sb200_comm_init": (i32, [P(vp), i32, i32, i32, vp, i32, i32]),
        "sb200_comm_destroy": (None, [vp]),
        "sb200_comm_last_error": (C.c_char_p, [vp]),
        "sb200_allgather_points": (i32, [vp, vp, vp, vp, vp, vp, i64, P(i64)]),
        "sb200_exchange_submit": (i32, [vp, vp, i32, i64]),
        "sb200_exchange_wait": (i32, [vp, i64, vp, vp, vp, vp, i64, P(i64)]),
        "sb200_exchange_device": (i32, [vp, i64, P(vp), P(vp), P(vp), P(i64)]),
        "sb200_exchange_drain": (i32, [vp, i64]),
        "sb200_comm_stats": (i32, [vp, P(dbl), P(i64), P(i64), i32]),
        "sb200_comm_set_consumer": (i32, [vp, i32]),
    }
    for name, (res, args) in protos.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _p(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if hasattr(a, "data_ptr"):  # torch tensor (pinned host memory in bench.py)
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


class StereoB200:
    """One camera pair on one GPU; mirrors oracle.pyoracle.CpuStereo method for method."""

    def __init__(self, pyrm_num, lowest_w, lowest_h, origin_w=None, origin_h=None, radius=2, ws=0.03, offset=2, device=0):
        self.lib = load()
        self.L = pyrm_num
        self.lowest = (lowest_w, lowest_h)
        self.h = C.c_void_p()
        rc = self.lib.sb200_ctx_create(C.byref(self.h), device, pyrm_num, lowest_w, lowest_h, origin_w or 0, origin_h or 0,
                                       radius, ws, offset)
        if rc != 0:
            msg = self.lib.sb200_last_error(self.h).decode() if self.h else ""
            status = self.lib.sb200_status_string(rc).decode()
            if self.h:
                self.lib.sb200_ctx_destroy(self.h)
                self.h = None
            raise StereoError(f"sb200_ctx_create: {status} {msg}")

    # ---- helpers ---------------------------------------------------------
    def _ck(self, rc, what):
        if rc != 0:
            raise StereoError(f"{what}: {self.lib.sb200_status_string(rc).decode()} - {self.lib.sb200_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.sb200_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def level_size(self, level):
        return (self.lowest[0] << level, self.lowest[1] << level)

    @property
    def top_size(self):
        return self.level_size(self.L - 1)

    # ---- staging ----------------------------------------------------------
    def set_pair(self, img0, img1, mask0, mask1):
        self._ck(self.lib.sb200_pair_upload(self.h, _p(img0), _p(img1), _p(mask0), _p(mask1)), "pair_upload")

    def stage_device(self, img0, img1, mask0, mask1):
        """Same as set_pair, from CUDA tensors / device pointers already resident on this GPU."""
        ptr = [C.c_void_p(a.data_ptr()) if hasattr(a, "data_ptr") else C.c_void_p(int(a)) for a in (img0, img1, mask0, mask1)]
        self._ck(self.lib.sb200_pair_stage_device(self.h, *ptr), "pair_stage_device")

    # ---- Rectify (device image half) -------------------------------------------------
    def rectify_view(self, view, src_bgr, src_mask, K=None, R_new=None, P_scaled=None, use_given_maps=False):
        src_bgr, src_mask = np.ascontiguousarray(src_bgr, np.uint8), np.ascontiguousarray(src_mask, np.uint8)
        mats = [None if a is None else np.ascontiguousarray(a, np.float64) for a in (K, R_new, P_scaled)]
        self._ck(self.lib.sb200_rectify_view(self.h, view, _p(src_bgr), _p(src_mask), src_mask.shape[1], src_mask.shape[0],
                                             *[_p(m) for m in mats], int(use_given_maps)), "rectify_view")

    def pair_build(self):
        self._ck(self.lib.sb200_pair_build(self.h), "pair_build")

    def get_rectify_maps(self):
        w, h = self.top_size
        m1, m2 = np.empty((h, w, 2), np.int16), np.empty((h, w), np.uint16)
        self._ck(self.lib.sb200_get_rectify_maps(self.h, _p(m1), _p(m2)), "get_rectify_maps")
        return m1, m2

    def set_rectify_maps(self, m1, m2):
        m1, m2 = np.ascontiguousarray(m1, np.int16), np.ascontiguousarray(m2, np.uint16)
        self._ck(self.lib.sb200_set_rectify_maps(self.h, _p(m1), _p(m2)), "set_rectify_maps")

    def get_remapped_mask(self):
        w, h = self.top_size
        out = np.empty((h, w), np.uint8)
        self._ck(self.lib.sb200_get_remapped_mask(self.h, _p(out)), "get_remapped_mask")
        return out

    def set_calib(self, Q, R, T):
        q, r, t = (np.ascontiguousarray(a, dtype=np.float64) for a in (Q, R, T))
        self._ck(self.lib.sb200_pair_set_calib(self.h, _p(q), _p(r), _p(t)), "pair_set_calib")

    def set_refine_iters(self, n):
        self._ck(self.lib.sb200_set_refine_iters(self.h, n), "set_refine_iters")

    def get_level(self, level, view):
        w, h = self.level_size(level)
        img = np.empty((h, w, 3), np.uint8)
        mask = np.empty((h, w), np.uint8)
        self._ck(self.lib.sb200_get_level(self.h, level, view, _p(img), _p(mask)), "get_level")
        return img, mask

    def get_margins(self, level):
        out = np.zeros((2, 6), np.int32)
        for v in (0, 1):
            self._ck(self.lib.sb200_get_margin(self.h, level, v, _p(out[v])), "get_margin")
        return out

    # ---- stages -------------------------------------------------------------
    def run_stage(self, level, stage):
        self._ck(self.lib.sb200_run_stage(self.h, level, stage), f"run_stage({level},{stage})")

    def match_one_layer(self, level):
        self._ck(self.lib.sb200_match_one_layer(self.h, level), "match_one_layer")

    def match_pair(self):
        n = C.c_int64()
        self._ck(self.lib.sb200_match_pair(self.h, C.byref(n)), "match_pair")
        return n.value

    def disparity_info(self):
        w, h, e = C.c_int(), C.c_int(), C.c_int()
        self.lib.sb200_disparity_info(self.h, C.byref(w), C.byref(h), C.byref(e))
        return w.value, h.value, e.value

    def get_disparity(self, dir_):
        w, h, e = self.disparity_info()
        if e == 0:
            return None
        out = np.empty((h, w), np.int16 if e == 2 else np.float64)
        self._ck(self.lib.sb200_get_disparity(self.h, dir_, _p(out)), "get_disparity")
        return out

    def set_disparity(self, dir_, arr):
        a = np.ascontiguousarray(arr)
        assert a.dtype in (np.int16, np.float64)
        self._ck(self.lib.sb200_set_disparity(self.h, dir_, _p(a), a.shape[1], a.shape[0], a.dtype.itemsize), "set_disparity")

    def get_rematch_bounds(self, dir_, level):
        w, h = self.level_size(level)
        bl = np.empty((h, w), np.int16)
        br = np.empty((h, w), np.int16)
        self._ck(self.lib.sb200_get_rematch_bounds(self.h, dir_, _p(bl), _p(br)), "get_rematch_bounds")
        return bl, br

    # ---- triangulation ----------------------------------------------------
    def to_cloud(self):
        n = C.c_int64()
        self._ck(self.lib.sb200_triangulate(self.h, C.byref(n)), "triangulate")
        return self.get_points(n.value)

    def get_points(self, n):
        xyz = np.empty((n, 3), np.float64)
        bgr = np.empty((n, 3), np.uint8)
        pix = np.empty(n, np.int32)
        self._ck(self.lib.sb200_get_points(self.h, _p(xyz), _p(bgr), _p(pix)), "get_points")
        return xyz, bgr, pix

    def points_device(self):
        x, b, p, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int64()
        self._ck(self.lib.sb200_points_device(self.h, C.byref(x), C.byref(b), C.byref(p), C.byref(n)), "points_device")
        return x.value, b.value, p.value, n.value

    def match_pair_host(self, img0, img1, mask0, mask1, Q, R, T, xyz_out, bgr_out, pix_out, capacity):
        """The one-call entry the C++ mirror uses: host buffers in, points out (H2D + D2H inside)."""
        n = C.c_int64()
        q, r, t = (np.ascontiguousarray(a, dtype=np.float64) for a in (Q, R, T))
        rc = self.lib.sb200_match_pair_host(self.h, _p(img0), _p(img1), _p(mask0), _p(mask1), _p(q), _p(r), _p(t),
                                            _p(xyz_out), _p(bgr_out), _p(pix_out), capacity, C.byref(n))
        self._ck(rc, "match_pair_host")
        return n.value

    # ---- instrumentation ----------------------------------------------------
    def stream(self):
        return self.lib.sb200_stream(self.h)

    def launch_count(self):
        return int(self.lib.sb200_launch_count(self.h))

    def graph_info(self):
        """(graph launches, graph instantiations) of sb200_match_pair on this context"""
        a, b = C.c_int64(), C.c_int64()
        self._ck(self.lib.sb200_graph_info(self.h, C.byref(a), C.byref(b)), "graph_info")
        return a.value, b.value

    def set_profiling(self, on):
        self._ck(self.lib.sb200_set_profiling(self.h, int(on)), "set_profiling")

    def stage_ms(self, reset=True):
        out = np.zeros(16, np.float64)
        self._ck(self.lib.sb200_get_stage_ms(self.h, _p(out), int(reset)), "get_stage_ms")
        return out

    def stage_level_ms(self, stage, level, reset=True):
        ms = C.c_double()
        self._ck(self.lib.sb200_get_stage_level_ms(self.h, stage, level, C.byref(ms), int(reset)), "get_stage_level_ms")
        return ms.value

    def refine_profile(self, level=-1, reset=True):
        """(sweep_ms, sweep_launches, px_iters) of the DisparityRefine sweep kernel since the last reset."""
        ms, n, px = C.c_double(), C.c_int64(), C.c_int64()
        self._ck(self.lib.sb200_get_refine_profile(self.h, level, C.byref(ms), C.byref(n), C.byref(px), int(reset)), "get_refine_profile")
        return ms.value, n.value, px.value

    def search_counters(self, reset=True):
        """(pixels handed to the list kernels by the K3 tile kernel, pixels that reached the exact FP64 pass)"""
        out = np.zeros(2, np.int64)
        self._ck(self.lib.sb200_get_search_counters(self.h, _p(out), int(reset)), "get_search_counters")
        return out

    def refine_counters(self, reset=True):
        out = np.zeros(2, np.int64)
        self._ck(self.lib.sb200_get_refine_counters(self.h, _p(out), int(reset)), "get_refine_counters")
        return out


def rectify_calib(K0, Rt0, K1, Rt1, origin, lowest_w, pyrm_num, opencv_compat=None):
    """Host half of Rectify (no GPU needed): dict with R_new [2,3,3], P_scaled [2,3,4], P_final [2,3,4], Q, R_final, T_final.
    opencv_compat: 245 (the OpenCV the reference links) or 413 (the OpenCV the golden vectors were made with); None = the
    library default (245 unless SB200_OPENCV_COMPAT says otherwise)."""
    a = [np.ascontiguousarray(m, np.float64) for m in (K0, Rt0, K1, Rt1)]
    out = {"R_new": np.zeros((2, 3, 3)), "P_scaled": np.zeros((2, 3, 4)), "P_final": np.zeros((2, 3, 4)), "Q": np.zeros((4, 4)),
           "R_final": np.zeros((3, 3)), "T_final": np.zeros(3)}
    outs = [_p(out[k]) for k in ("R_new", "P_scaled", "P_final", "Q", "R_final", "T_final")]
    if opencv_compat is None:
        rc = load().sb200_rectify_calib(*[_p(m) for m in a], int(origin[0]), int(origin[1]), int(lowest_w), int(pyrm_num), *outs)
    else:
        rc = load().sb200_rectify_calib_compat(*[_p(m) for m in a], int(origin[0]), int(origin[1]), int(lowest_w), int(pyrm_num),
                                               int(opencv_compat), *outs)
    if rc != 0:
        raise StereoError("rectify_calib: bad argument")
    return out


def stereo_rectify_host(K1, K2, size, R, T, opencv_compat=413):
    a = [np.ascontiguousarray(m, np.float64) for m in (K1, K2, R, T)]
    R1, R2, P1, P2, Q = np.zeros((3, 3)), np.zeros((3, 3)), np.zeros((3, 4)), np.zeros((3, 4)), np.zeros((4, 4))
    rc = load().sb200_stereo_rectify_host(_p(a[0]), _p(a[1]), int(size[0]), int(size[1]), _p(a[2]), _p(a[3]), int(opencv_compat), _p(R1),
                                          _p(R2), _p(P1), _p(P2), _p(Q))
    if rc != 0:
        raise StereoError("stereo_rectify_host: bad argument")
    return R1, R2, P1, P2, Q


def exp_host(x: float) -> float:
    return load().sb200_exp_host(float(x))


def sink_filter(xyz, sor_meank, sor_std_mul, normal_radius, cam_center, device=0):
    """CCloudOptimization::filter's point processing on the GPU (sb200_sink_filter): returns (records [m,7] f32 =
    x y z nx ny nz curvature of the kept points in input order, kept_index [m] i32, stats dict)."""
    xyz = np.ascontiguousarray(xyz, np.float64)
    n = xyz.shape[0]
    out = np.empty((n, 7), np.float32)
    kept = np.empty(n, np.int32)
    cam = np.ascontiguousarray(cam_center, np.float64)
    m = C.c_int64()
    stats = np.zeros(5)
    lib = load()
    rc = lib.sb200_sink_filter(device, _p(xyz), n, int(sor_meank), float(sor_std_mul), float(normal_radius), _p(cam), _p(out), _p(kept), n,
                               C.byref(m), _p(stats))
    if rc != 0:
        raise StereoError(f"sink_filter: {lib.sb200_status_string(rc).decode()} - {lib.sb200_sink_last_error().decode()}")
    return out[:m.value].copy(), kept[:m.value].copy(), {"mean": stats[0], "stddev": stats[1], "threshold": stats[2], "device_ms": stats[3],
                                                        "widened_queries": int(stats[4])}


def comm_unique_id() -> bytes:
    """NCCL unique id (call on one rank, hand the bytes to the others over the launcher's bootstrap)."""
    buf = C.create_string_buffer(UNIQUE_ID_BYTES)
    lib = load()
    rc = lib.sb200_comm_unique_id(buf)
    if rc != 0:
        raise StereoError(f"comm_unique_id: {lib.sb200_status_string(rc).decode()} - {lib.sb200_comm_last_error(None).decode()}")
    return buf.raw


class PointComm:
    """The exchange step behind the C ABI (sb200_comm_*): one communicator per GPU."""

    def __init__(self, device, rank, nranks, unique_id: bytes, producers=1, slots=2):
        self.lib = load()
        self.rank, self.nranks, self.producers = rank, nranks, producers
        self.h = C.c_void_p()
        idb = C.create_string_buffer(unique_id, UNIQUE_ID_BYTES)
        rc = self.lib.sb200_comm_init(C.byref(self.h), device, rank, nranks, idb, producers, slots)
        if rc != 0:
            msg = self.lib.sb200_comm_last_error(self.h).decode()
            if self.h:
                self.lib.sb200_comm_destroy(self.h)
                self.h = None
            raise StereoError(f"sb200_comm_init: {self.lib.sb200_status_string(rc).decode()} - {msg}")

    def _ck(self, rc, what):
        if rc != 0:
            raise StereoError(f"{what}: {self.lib.sb200_status_string(rc).decode()} - {self.lib.sb200_comm_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.sb200_comm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def allgather_points(self, ctx: StereoB200, capacity: int, want_host=True):
        """Synchronous gather of ctx's points from every rank: (counts [nranks], xyz, bgr, pix) in rank-major order."""
        counts = np.zeros(self.nranks, np.int64)
        total = C.c_int64()
        if want_host:
            xyz, bgr, pix = np.empty((capacity, 3)), np.empty((capacity, 3), np.uint8), np.empty(capacity, np.int32)
        else:
            xyz = bgr = pix = None
        self._ck(self.lib.sb200_allgather_points(self.h, ctx.h, _p(counts), _p(xyz), _p(bgr), _p(pix), capacity, C.byref(total)), "allgather_points")
        n = total.value
        if not want_host:
            return counts, None, None, None
        return counts, xyz[:n], bgr[:n], pix[:n]

    def submit(self, ctx: StereoB200, producer: int, seq: int):
        self._ck(self.lib.sb200_exchange_submit(self.h, ctx.h, producer, seq), "exchange_submit")

    def wait(self, ticket: int, capacity: int = 0, want_host=False):
        counts = np.zeros(self.nranks, np.int64)
        total = C.c_int64()
        if want_host:
            xyz, bgr, pix = np.empty((capacity, 3)), np.empty((capacity, 3), np.uint8), np.empty(capacity, np.int32)
        else:
            xyz = bgr = pix = None
        self._ck(self.lib.sb200_exchange_wait(self.h, ticket, _p(counts), _p(xyz), _p(bgr), _p(pix), capacity, C.byref(total)), "exchange_wait")
        n = total.value
        if not want_host:
            return counts, n
        return counts, xyz[:n], bgr[:n], pix[:n]

    def set_consumer(self, on=True):
        self._ck(self.lib.sb200_comm_set_consumer(self.h, int(on)), "comm_set_consumer")

    def drain(self, n_tickets: int):
        self._ck(self.lib.sb200_exchange_drain(self.h, n_tickets), "exchange_drain")

    def stats(self, reset=True):
        ms, nb, nx = C.c_double(), C.c_int64(), C.c_int64()
        self._ck(self.lib.sb200_comm_stats(self.h, C.byref(ms), C.byref(nb), C.byref(nx), int(reset)), "comm_stats")
        return {"collective_ms": ms.value, "bytes_received": nb.value, "exchanges": nx.value}
