#!/usr/bin/env python
"""Per-source-line totals (instructions executed, stall samples) from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = []
fname = "?"
for r in rows:
    if len(r) >= 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]
    if len(r) > 7 and r[0].isdigit():
        try:
            out.append((int(r[7]), int(r[6]) if r[6].isdigit() else 0, fname, int(r[0]), r[1].strip()[:110]))
        except ValueError:
            pass
ti = sum(o[0] for o in out); ts = sum(o[1] for o in out)
print("total instr %d samples %d" % (ti, ts))
for o in sorted(out, key=lambda x: -x[0])[:top]:
    print("%5.1f%% instr %5.1f%% stall  %s:%d  %s" % (100.0 * o[0] / ti, 100.0 * o[1] / max(ts, 1), o[2], o[3], o[4]))
