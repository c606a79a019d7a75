"""Stage a synthetic calibrated rig on disk in the reference's schema:
    python tools/stage_rig.py OUT_DIR PyrmNum LowestLevelWidth LowestLevelHeight [n_pairs]
then  reconstruction_b200/host/reconstruction OUT_DIR/config.yml"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reconstruction_b200 import stage

out, L, w0, h0 = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
n = int(sys.argv[5]) if len(sys.argv) > 5 else 1
cfg, _ = stage.write_dataset(out, L, w0, h0, n_pairs=n)
print(cfg)
