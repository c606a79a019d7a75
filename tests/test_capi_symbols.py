"""CPU-side checks of the C ABI library: it loads, exports every symbol include/stereo_b200.h
declares, refuses to work without a GPU (no CPU fallback), and its exp() twin returns the C
library's bits on this machine (the refinement is chaotic at the ulp level)."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

from reconstruction_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    capi.build()
    return capi.load()


def test_header_and_library_agree(lib):
    hdr = open(os.path.join(ROOT, "include", "stereo_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(sb200_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == sorted(capi.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_no_cpu_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = lib.sb200_ctx_create(C.byref(h), 0, 2, 48, 36, 0, 0, 2, 0.03, 2)
    assert rc == 1  # SB200_ERR_NO_DEVICE
    assert not h
    with pytest.raises(capi.StereoError):
        capi.StereoB200(2, 48, 36)


def test_exp_twin_matches_libm(lib):
    rng = np.random.default_rng(123)
    xs = np.concatenate([
        -rng.random(200000) * 12, -rng.random(50000) * 800, -(rng.random(50000) ** 3) * 1e-3,
        -np.arange(0, 800, dtype=np.float64), -(np.arange(0, 60) * 0.5) ** 2,
        -np.square(rng.integers(-30, 30, 20000) + rng.random(20000)),
        np.array([0.0, -0.0, -1e-300, -1e-20, -745.0, -745.2, -744.9, -708.3, -708.5, -1023.9, -1024.0, -1e9, -np.inf]),
    ])
    f = lib.sb200_exp_host
    bad = [x for x in xs if np.float64(f(float(x))).view(np.int64) != np.float64(math.exp(x)).view(np.int64)]
    assert not bad, (len(bad), bad[:5])
