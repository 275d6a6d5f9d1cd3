#!/usr/bin/env python
"""Sums dram__bytes_read/write and durations per kernel name over an ncu CSV of ONE training step (bench.py --ncu-step),
writes profiles/<out>.json: per-step DRAM traffic (roofline.step_traffic) and the traffic of the dominant GEMM launch."""
import csv
import json
import sys
from collections import defaultdict


def main(path, out):
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    per = defaultdict(lambda: {"launches": 0, "us": 0.0, "read": 0.0, "write": 0.0})
    launches = defaultdict(dict)
    for r in csv.DictReader(lines):
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        unit = r.get("Metric Unit", "")
        name = r["Kernel Name"].split("(")[0][:80]
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        elif m.startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        launches[(r["ID"], name)][m] = v
    for (_, name), d in launches.items():
        p = per[name]
        p["launches"] += 1
        p["us"] += d.get("gpu__time_duration.sum", 0.0)
        p["read"] += d.get("dram__bytes_read.sum", 0.0)
        p["write"] += d.get("dram__bytes_write.sum", 0.0)
    tot_r = sum(p["read"] for p in per.values())
    tot_w = sum(p["write"] for p in per.values())
    pair = [(d.get("gpu__time_duration.sum", 0.0), d) for (i, n), d in launches.items() if "gemm_tc_kernel<0, 2" in n]
    dom = max(pair, key=lambda x: x[0])[1] if pair else {}
    res = {"step_dram_bytes": tot_r + tot_w, "step_dram_read": tot_r, "step_dram_write": tot_w,
           "step_kernel_us_serialised": sum(p["us"] for p in per.values()),
           "dominant_gemm_launch": {"us": dom.get("gpu__time_duration.sum"),
                                    "dram_bytes": dom.get("dram__bytes_read.sum", 0.0) + dom.get("dram__bytes_write.sum", 0.0)},
           "per_kernel": {k: v for k, v in sorted(per.items(), key=lambda kv: -kv[1]["us"])}}
    json.dump(res, open(out, "w"), indent=1)
    print("step DRAM traffic %.2f GB (read %.2f, write %.2f), serialised kernel time %.1f us" % ((tot_r + tot_w) / 1e9, tot_r / 1e9, tot_w / 1e9, res["step_kernel_us_serialised"]))
    for k, v in list(res["per_kernel"].items())[:12]:
        print("%-70s %4d %9.1f us %8.1f MB" % (k, v["launches"], v["us"], (v["read"] + v["write"]) / 1e6))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
