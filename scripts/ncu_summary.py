#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import csv
import sys
from collections import defaultdict


def main(path, skip_before=None):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        rows.append((r["Kernel Name"], v * scale))
    tot = sum(v for _, v in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for k, v in rows:
        k = k.split("(")[0][:90]
        agg[k][0] += 1
        agg[k][1] += v
    print("launches %d total %.1f us" % (len(rows), tot))
    print("%-90s %7s %12s %7s" % ("kernel", "count", "total_us", "share"))
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-90s %7d %12.1f %6.1f%%" % (k, n, v, 100 * v / max(tot, 1e-9)))


if __name__ == "__main__":
    main(sys.argv[1])
