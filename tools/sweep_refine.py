"""Time the fused DisparityRefine for several (sweeps per launch T, tile variant) settings on one synthetic pair.
usage: python tools/sweep_refine.py L lowest_w lowest_h "T:tile,T:tile,..." """
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from reconstruction_b200 import capi, synth

L, w0, h0 = (int(a) for a in sys.argv[1:4])
combos = [tuple(int(v) for v in c.split(":")) for c in sys.argv[4].split(",")]
t = time.time(); sp = synth.make_pair(w0, h0, L, pair_id=0); print("synth %.1fs" % (time.time() - t), sp.top_size, flush=True)
ref = None
for T, tile in combos:
    os.environ["SB200_REFINE_T"] = str(T)
    os.environ["SB200_REFINE_TILE"] = str(tile)
    g = capi.StereoB200(L, w0, h0)
    g.set_calib(sp.Q, sp.R_final, sp.T_final)
    for r in range(2):
        g.set_profiling(r == 1)
        g.set_pair(*sp.image, *sp.mask)
        t1 = time.time(); n = g.match_pair(); t2 = time.time()
    ms = g.stage_ms()
    top = g.refine_profile(level=L - 1, reset=False)
    sm, sl, spx = g.refine_profile()
    d0 = g.get_disparity(0)
    same = "ref" if ref is None else ("same" if np.array_equal(ref.view(np.int64), d0.view(np.int64)) else "DIFFERENT")
    if ref is None:
        ref = d0
    print(f"T={T} tile={tile}: match_pair {1e3*(t2-t1):.1f} ms, refine stage {ms[9]:.2f} ms, sweeps {sm:.2f} ms "
          f"(top level {top[0]:.2f} ms, {22*top[2]/(top[0]*1e-3)/1e9:.0f} GB/s algorithmic), initial {ms[2]:.2f} ms, "
          f"miss evals {g.refine_counters()[1]}, points {n}, launches {g.launch_count()}, result {same}", flush=True)
    g.close()
