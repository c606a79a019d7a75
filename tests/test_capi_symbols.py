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


def test_header_is_plain_c_and_links(tmp_path):
    """include/stereo_b200.h is a C header (no C++ / torch types in the boundary): a C99 program that includes it and takes the
    address of entry points compiles with gcc -std=c99 -pedantic and links against libstereo_b200.so; without a GPU the calls
    report SB200_ERR_NO_DEVICE instead of computing anything."""
    import subprocess

    capi.build()
    root = os.path.dirname(capi.HERE)
    src = tmp_path / "abi.c"
    src.write_text(
        '#include "stereo_b200.h"\n#include <stdio.h>\n'
        "int main(void) {\n"
        "  sb200_ctx* ctx = NULL;\n"
        "  int rc = sb200_ctx_create(&ctx, 0, 3, 64, 48, 0, 0, 2, 0.03, 2);\n"
        "  double xyz[9] = {0, 0, 1, 1, 0, 1, 0, 1, 1}, cam[3] = {0, 0, 0};\n"
        "  float rec[21]; int64_t kept = 0;\n"
        "  int rc2 = sb200_sink_filter(0, xyz, 3, 2, 1.0, 2.5, cam, rec, NULL, 3, &kept, NULL);\n"
        '  printf("%d %s | %d %s\\n", rc, sb200_status_string(rc), rc2, sb200_sink_last_error());\n'
        "  if (ctx) sb200_ctx_destroy(ctx);\n"
        "  return 0;\n}\n")
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(root, "include"), str(src), "-o", str(exe),
                    "-L", capi.HERE, "-lstereo_b200", "-Wl,-rpath," + capi.HERE], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except ImportError:
        has_gpu = False
    if not has_gpu:
        assert r.stdout.startswith("1 no CUDA device") and "| 1 " in r.stdout, r.stdout
