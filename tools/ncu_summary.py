"""Summarise ncu output into small JSON files for profiles/.
    python tools/ncu_summary.py raw    <raw.csv>      <out.json> "<command>"      # ncu -i X.ncu-rep --page raw --csv
    python tools/ncu_summary.py launch <launches.csv> <out.json> "<command>"      # ncu --metrics gpu__time_duration.sum --csv
"""
import csv
import json
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum"]


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return v


def raw(path, out, cmd):
    r = list(csv.reader(open(path)))
    hdr, units = r[0], r[1]
    res = []
    for vals in r[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        k = {"kernel": d.get("Kernel Name"), "metrics": {}, "stall_ratio_per_issue": {}}
        for key in KEYS:
            if key in d:
                k["metrics"][key] = {"value": num(d[key]), "unit": u.get(key, "")}
        for key in hdr:
            if key.startswith("smsp__average_warps_issue_stalled_") and key.endswith("_per_issue_active.ratio"):
                v = num(d[key])
                if isinstance(v, float) and v >= 0.1:
                    k["stall_ratio_per_issue"][key[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(v, 3)
        res.append(k)
    json.dump({"command": cmd, "kernels": res}, open(out, "w"), indent=1)


def launch(path, out, cmd):
    rows, hdr = [], None
    with open(path) as f:
        for line in f:
            if line.startswith('"ID"'):
                hdr = next(csv.reader([line]))
                break
        for r in csv.reader(f):
            if len(r) >= len(hdr):
                rows.append(dict(zip(hdr, r)))
    t, n = defaultdict(float), defaultdict(int)
    for r in rows:
        name = r["Kernel Name"].split("(")[0]
        t[name] += num(r["Metric Value"]) / 1000.0
        n[name] += 1
    tot = sum(t.values())
    ks = [{"kernel": k, "launches": n[k], "us": round(v, 1), "share": round(v / tot, 4)} for k, v in sorted(t.items(), key=lambda kv: -kv[1])]
    json.dump({"command": cmd, "total_us": round(tot, 1), "launches": len(rows), "kernels": ks}, open(out, "w"), indent=1)


if __name__ == "__main__":
    {"raw": raw, "launch": launch}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
