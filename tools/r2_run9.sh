#!/bin/bash
# round 2, run 9 (one B200): source-level ncu capture of the kNN kernel on config 2 (one 262 144-query chunk in Morton
# order through the pipeline), plus the launch list of the classic path.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2n; mkdir -p $O
A="--workload config2 --order 3 --steps 1 --warmup 1 --no-cpu-baseline --no-library-baseline --no-check"
timeout 300 python bench.py $A > $O/bench.jsonl 2> $O/bench.err; echo "bench rc $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches.csv python bench.py $A > $O/ncu1.log 2>&1; echo "ncu list rc $?"
ncu --set full --clock-control none --import-source on -k regex:"knn_kernel" -s 3 -c 1 -o $O/knn -f python bench.py $A > $O/ncu2.log 2>&1; echo "ncu knn rc $?"
ncu -i $O/knn.ncu-rep --page raw --csv > $O/knn_raw.csv 2>/dev/null
ncu -i $O/knn.ncu-rep --page source --csv --print-source cuda > $O/knn_lines.csv 2>/dev/null; echo "export rc $?"
ls -la $O
