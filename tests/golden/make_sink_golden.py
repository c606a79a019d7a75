"""Golden vector for the sink filter: a small surface cloud with isolated points and what
oracle/sink_oracle.py makes of it with the reference's parameters scaled to the sampling (meanK 30, 1 sigma, radius 1.5).
    python tests/golden/make_sink_golden.py  ->  tests/golden/sink_small.npz
PCL is absent (parity unpinned against it); the fixture pins the checker against drift and gives the GPU test a committed
expectation next to the live one."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import sink_oracle as so  # noqa: E402


def cloud():
    rng = np.random.default_rng(20260202)
    n = 56
    u, v = np.meshgrid(np.arange(n) * 0.3 - n * 0.15, np.arange(n) * 0.3 - n * 0.15)
    z = 800 + 3 * np.sin(u / 4.0) * np.cos(v / 3.0) + rng.normal(0, 0.02, u.shape)
    p = np.stack([u + rng.normal(0, 0.02, u.shape), v + rng.normal(0, 0.02, u.shape), z], -1).reshape(-1, 3)
    far = rng.integers(0, len(p), 25)
    p[far] += rng.uniform(-1, 1, (25, 3)) * np.array([[8, 8, 30]])
    p[7] = p[8]
    return p


def main():
    p = cloud()
    cam = np.array([40.0, -25.0, 0.0])
    rec, kept, info = so.sink_filter(p, 30, 1.0, 1.5, cam)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sink_small.npz")
    np.savez_compressed(out, xyz=p, cam=cam, mean_k=30, std_mul=1.0, radius=1.5, records=rec, kept=kept, mean_dist=info["mean_dist"],
                        stats=np.array([info["mean"], info["stddev"], info["threshold"]]), eigen_gap=info["eigen_gap"])
    print(len(p), "points,", len(kept), "kept,", os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
