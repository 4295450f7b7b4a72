#!/usr/bin/env python
"""Per-CUDA-line instruction / stall-sample shares from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.

    ncu -i rep --page source --csv --print-source cuda,sass > lines.csv;  python tools/ncu_lines.py lines.csv [min_pct]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], newline="")))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
data = []
fname, hdr = "", None
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
        iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
        iX = hdr.index("L1 Wavefronts Shared Excessive")
    elif hdr and len(r) >= len(hdr) and r[0].isdigit():
        try:
            data.append((fname, int(r[0]), r[1], float(r[iI] or 0), float(r[iS] or 0), float(r[iX] or 0)))
        except ValueError:
            pass
ti, ts = sum(d[3] for d in data), sum(d[4] for d in data)
print("total warp instructions %.4g, stall samples %.4g" % (ti, ts))
for fn, ln, src, ie, smp, ex in data:
    if 100 * ie / ti >= thr or 100 * smp / ts >= thr:
        print("%-14s %5d  %5.1f%% inst  %5.1f%% samples  %9.3g xs-smem-wavefronts | %s" % (fn[:14], ln, 100 * ie / ti, 100 * smp / ts, ex, src.strip()[:110]))
