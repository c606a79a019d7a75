"""BASELINE.json configs[4] end to end: a 20-view 6000x4000 synthetic rig (19 adjacent pairs, 5-level pyramid, lowest 375x250)
through the C++ CLI `reconstruction config.yml` on every visible GPU: staged frames -> matcher on all devices -> NCCL point
all-gather (C ABI) -> sink filter per pair -> merged PLY.  Two pairs are synthesised and repeated (hard links): the content is
synthetic anyway and the data-set generation would otherwise cost more box time than the run.

    python tools/run_config_e.py OUT_JSON [n_views] [lowest_w lowest_h levels]
"""
import json
import os
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reconstruction_b200 import capi, stage  # noqa: E402

out_json = sys.argv[1]
n_views = int(sys.argv[2]) if len(sys.argv) > 2 else 20
w0, h0, L = (int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (375, 250, 5)
HOST = os.path.join(os.path.dirname(capi.HERE), "reconstruction_b200", "host")
root = tempfile.mkdtemp(prefix="configE_")
t0 = time.time()
cfg, pairs = stage.write_dataset(root, L, w0, h0, n_pairs=n_views - 1, isoutput=0, distinct=2)
t_data = time.time() - t0
import torch  # noqa: E402

ngpu = torch.cuda.device_count()
env = dict(os.environ, SB200_CTX_PER_DEVICE=os.environ.get("SB200_CTX_PER_DEVICE", "2"))
t0 = time.time()
r = subprocess.run([os.path.join(HOST, "reconstruction"), cfg], cwd=root, capture_output=True, text=True, env=env)
wall = time.time() - t0
lines = r.stdout.splitlines()
keep = [ln for ln in lines if any(k in ln for k in ("Matching time", "total time", "point all-gather", "points", "sink", "failed", "error"))]
ply = os.path.join(root, "out.ply")
n_pts = None
if os.path.exists(ply):
    with open(ply, "rb") as f:
        for _ in range(20):
            ln = f.readline()
            if ln.startswith(b"element vertex"):
                n_pts = int(ln.split()[-1])
                break
W, H = w0 << (L - 1), h0 << (L - 1)
res = {"config": f"{n_views}-view {W}x{H} synthetic rig, {L}-level pyramid, {n_views - 1} adjacent pairs (2 distinct, repeated), {ngpu} GPU(s)",
       "command": "reconstruction config.yml (C++ host mirror; staged frames; SB200_CTX_PER_DEVICE=%s)" % env["SB200_CTX_PER_DEVICE"],
       "returncode": r.returncode, "wall_s": wall, "dataset_s": t_data, "merged_points": n_pts,
       "Mpix_per_s_wall": (n_views - 1) * W * H / wall / 1e6, "stdout_lines": keep[-40:], "stderr_tail": r.stderr[-600:]}
json.dump(res, open(out_json, "w"), indent=1)
print(json.dumps({k: res[k] for k in ("config", "returncode", "wall_s", "merged_points", "Mpix_per_s_wall")}))
subprocess.run(["rm", "-rf", root])
