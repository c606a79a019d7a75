"""The sink filter on a disparity-map-like cloud (for ncu / timing): python tools/prof_sink.py [n_side=2700] [spacing=0.4]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from reconstruction_b200 import capi

n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 2700
sp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
rng = np.random.default_rng(1)
u, v = np.meshgrid(np.arange(n_side) * sp - n_side * sp / 2, np.arange(n_side) * sp - n_side * sp / 2)
z = 1000 + 40 * np.sin(u / 90.0) * np.cos(v / 70.0) + rng.normal(0, 0.03, u.shape)
p = np.stack([u + rng.normal(0, 0.02, u.shape), v + rng.normal(0, 0.02, u.shape), z], -1).reshape(-1, 3)
far = rng.integers(0, len(p), 2000)
p[far] += rng.uniform(-1, 1, (2000, 3)) * np.array([[20, 20, 60]])
for rep in range(2):
    t = time.time()
    rec, kept, st = capi.sink_filter(p, 100, 1.0, 2.5, [0.0, 0.0, 0.0])
    print("points %d kept %d wall %.3f s" % (len(p), len(kept), time.time() - t), st, flush=True)
