#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2g; mkdir -p $O
python tools/wls_probe.py --n 262144 --orders 3 > $O/wls_probe_plain.log 2>&1; echo "plain rc $?"
ncu --set full --clock-control none --import-source on -k regex:"wls_kernel" -s 3 -c 1 -o $O/wls -f python tools/wls_probe.py --n 262144 --orders 3 > $O/ncu_wls.log 2>&1; echo "ncu rc $?"
ncu -i $O/wls.ncu-rep --page raw --csv > $O/wls_raw.csv 2>/dev/null
ls -la $O
