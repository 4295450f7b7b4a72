#!/usr/bin/env python
"""DRAM traffic per launch of the dominant kernels from `ncu --set full` raw-page exports -> profiles/ncu_traffic.json,
the file bench.py quotes in `roofline.traffic` (a bench run cannot measure it itself: the counters need ncu).

    python tools/ncu_traffic.py profiles/r2_ncu_full_chain_and_head.csv:16384 profiles/r2_ncu_full_knn_pca_wls.csv:262144

Each argument is <raw csv>:<patches per launch in that capture>.  For every kernel family the mean of
dram__bytes_read.sum + dram__bytes_write.sum over its captured launches is stored in bytes, with the capture it came from."""
import csv
import json
import os
import sys

FAMILIES = {"chain_max_kernel": "chain_max_kernel", "head_fused_kernel": "head_fused_kernel", "wls_kernel": "wls_kernel",
            "knn_kernel": "knn_kernel", "pca_kernel": "pca_kernel"}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    out = {}
    for arg in sys.argv[1:]:
        path, per = arg.rsplit(":", 1)
        rows = list(csv.reader(open(path, newline="")))
        hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
        hdr, units = rows[hi], rows[hi + 1]
        col = {n: i for i, n in enumerate(hdr)}
        acc = {}
        for r in rows[hi + 2:]:
            if len(r) < len(hdr):
                continue
            fam = next((f for f, pat in FAMILIES.items() if pat in r[col["Kernel Name"]]), None)
            if fam is None:
                continue
            b = 0.0
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                b += float(r[col[m]]) * SCALE[units[col[m]]]
            acc.setdefault(fam, []).append((b, float(r[col["gpu__time_duration.sum"]])))
        for fam, v in acc.items():
            out[fam] = {"dram_bytes_per_launch": sum(x[0] for x in v) / len(v), "launches_captured": len(v),
                        "patches_per_launch": int(per), "ncu_duration_ms_mean": sum(x[1] for x in v) / len(v),
                        "source": os.path.relpath(path)}
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "ncu_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
