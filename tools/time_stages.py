"""Per-stage device time of one pair (CUDA events around every MatchOneLayer stage).
usage: python tools/time_stages.py L lowest_w lowest_h [reps]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from reconstruction_b200 import capi, synth

L, w0, h0 = (int(a) for a in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
t = time.time(); sp = synth.make_pair(w0, h0, L, pair_id=0); print("synth %.1fs" % (time.time() - t), sp.top_size, flush=True)
g = capi.StereoB200(L, w0, h0)
g.set_calib(sp.Q, sp.R_final, sp.T_final)
names = ["pyramid"] + [capi.STAGE_NAMES[i] for i in range(1, 11)] + ["DisparityToCloud"]
for r in range(reps):
    g.set_profiling(r == reps - 1)
    t = time.time(); g.set_pair(*sp.image, *sp.mask); t1 = time.time(); n = g.match_pair(); t2 = time.time()
    print(f"rep {r}: upload {1e3*(t1-t):.1f} ms, match_pair {1e3*(t2-t1):.1f} ms, points {n}, launches {g.launch_count()}", flush=True)
sm, sl, spx = g.refine_profile()
lvl = [[g.stage_level_ms(st, lv, reset=False) for lv in range(L)] for st in range(13)]
ms = g.stage_ms()
for i, nm in enumerate(names):
    print(f"  {i:2d} {nm:22s} {ms[i]:9.3f} ms")
print("  per level (ms): stage " + " ".join(f"L{lv:<6d}" for lv in range(L)))
for st in list(range(2, 11)) + [12]:
    print(f"  {st:2d} {(capi.STAGE_NAMES.get(st) or 'RefineSweeps'):22s} " + " ".join(f"{lvl[st][lv]:7.3f}" for lv in range(L)))
print("     level totals           " + " ".join(f"{sum(lvl[st][lv] for st in range(2, 11)):7.3f}" for lv in range(L)))
sc = g.search_counters(reset=False)
print("  total stages %.3f ms; refine out-of-table evals %d; NCC search: %d px listed by the band kernel, %d px to the exact pass (all reps)" % (ms[:12].sum(), g.refine_counters()[1], sc[0], sc[1]))
if sl:
    print("  refine sweeps: %.3f ms over %d sweeps (%.1f us/sweep), %.3f G px-iter, %.1f GB/s algorithmic (22 B/px-iter)" % (sm, sl, 1e3 * sm / sl, spx / 1e9, 22 * spx / (sm * 1e-3) / 1e9))
W, H = sp.top_size
print("  Mpix/s (match_pair wall): %.1f" % (W * H / (t2 - t1) / 1e6))
