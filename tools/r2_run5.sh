#!/bin/bash
# round 2, run 5: source-level ncu capture of knn_kernel (classic workload, one 65 536-query chunk of a 262 144-point torus)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2e; mkdir -p $O
python bench.py --workload config2 --steps 1 --warmup 1 --points 262144 --no-cpu-baseline > $O/bench_small_classic.jsonl 2>&1; echo "small classic rc $?"
ncu --set full --clock-control none --import-source on -k regex:"knn_kernel" -s 2 -c 1 -o $O/knn -f python bench.py --workload config2 --steps 1 --warmup 1 --points 262144 --no-cpu-baseline > $O/ncu.log 2>&1; echo "ncu knn rc $?"
ncu -i $O/knn.ncu-rep --page source --csv > $O/knn_source.csv 2>/dev/null; echo "export rc $?"
ls -la $O
