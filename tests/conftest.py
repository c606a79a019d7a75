"""Shared pytest configuration: the `gpu` marker and the oracle build.

`-m "not gpu"` runs on the CPU-only build container (oracle vs golden vectors, host logic,
C-ABI symbol check, gloo world_size-2).  `-m gpu` runs on a B200 and calls the CUDA path
through the C ABI, comparing it with the oracle.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The ctypes front-end of the CPU checkers, built on demand."""
    from oracle import pyoracle

    pyoracle.build()  # `ref` is skipped by make when /root/reference is absent (GPU box)
    return pyoracle


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
