"""One pair through match_pair (for ncu): python tools/prof_pair.py L lowest_w lowest_h [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reconstruction_b200 import capi, synth

L, w0, h0 = (int(a) for a in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
sp = synth.make_pair(w0, h0, L, pair_id=0)
g = capi.StereoB200(L, w0, h0)
g.set_calib(sp.Q, sp.R_final, sp.T_final)
for _ in range(reps):
    g.set_pair(*sp.image, *sp.mask)
    print("points", g.match_pair(), "launches", g.launch_count(), flush=True)
