#!/bin/bash
# round 2, run 7 (one B200): full GPU tests of the sub-grid kNN / config-3 tree, kNN register-budget A/B, config 3 at size,
# default + classic bench lines
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2k; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -x -p no:cacheprovider --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -14 $O/pytest_gpu.log
for v in "" ab_libs/lib_knn_minb4.so ab_libs/lib_knn_head.so; do
  echo "== lib=$v"
  for a in "--points 1000000 --cloud torus --k 256 --order 3 --queries 1000000 --morton" "--points 30000000 --cloud density --k 256 --order 3 --queries 1000000 --morton" "--points 1000000 --cloud torus --k 256 --order 3 --queries 1000000"; do
    DEEPFIT_B200_LIB=$v timeout 300 python tools/scale_probe.py $a 2>&1 | tail -1 | cut -c1-420
  done
done 2>&1 | tee $O/knn_ab2.log
timeout 900 python tools/config3_run.py > $O/config3.jsonl 2> $O/config3.err; echo "config3 rc $?"; cat $O/config3.jsonl; tail -3 $O/config3.err
timeout 600 python bench.py --workload config2 --order 3 --steps 3 --no-cpu-baseline > $O/bench_config2_o3.jsonl 2> $O/bench_config2_o3.err; echo "bench config2 rc $?"; cut -c1-200 $O/bench_config2_o3.jsonl; tail -3 $O/bench_config2_o3.err
timeout 600 python bench.py --steps 5 --no-cpu-baseline > $O/bench_default.jsonl 2> $O/bench_default.err; echo "bench default rc $?"; cut -c1-200 $O/bench_default.jsonl; tail -3 $O/bench_default.err
