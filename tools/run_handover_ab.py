"""A/B of the host mirror's hand-over to the sink on ONE GPU: the same staged data set (n_views-1 pairs of 4096x3072, 5 levels,
2 distinct pairs repeated) through two CLI binaries and, for the current one, through the gather path as well.

    python tools/run_handover_ab.py OUT_JSON n_views BIN[:ENV=VAL,...] ...
"""
import json
import os
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reconstruction_b200 import stage  # noqa: E402

out_json, n_views = sys.argv[1], int(sys.argv[2])
root = tempfile.mkdtemp(prefix="handover_")
t0 = time.time()
cfg, _ = stage.write_dataset(root, 5, 256, 192, n_pairs=n_views - 1, isoutput=0, distinct=2)
res = {"dataset_s": time.time() - t0, "pairs": n_views - 1, "frame": "4096x3072, 5 levels", "runs": []}
for spec in sys.argv[3:]:
    binp, _, envs = spec.partition(":")
    env = dict(os.environ, SB200_DEVICES="0")
    for kv in filter(None, envs.split(",")):
        k, v = kv.split("=")
        env[k] = v
    t0 = time.time()
    r = subprocess.run([os.path.abspath(binp), cfg], cwd=root, capture_output=True, text=True, env=env)
    wall = time.time() - t0
    keep = [ln for ln in r.stdout.splitlines() if any(k in ln for k in ("Matching time", "total time", "point all-gather", "wrote", "failed", "error"))]
    res["runs"].append({"binary": binp, "env": envs, "returncode": r.returncode, "wall_s": wall, "lines": keep, "stderr": r.stderr[-300:]})
    json.dump(res, open(out_json, "w"), indent=1)
print(json.dumps(res)[:1500])
subprocess.run(["rm", "-rf", root])
