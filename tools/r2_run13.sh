#!/bin/bash
# round 2, run 13 (one B200): kNN with one selection function, no early selection under a radius guess, 5 CTAs per SM:
# tests, then build variants twice in alternating order on the same box.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2r; mkdir -p $O
timeout 1800 python -m pytest tests/test_gpu_knn.py tests/test_gpu_pipeline.py -m gpu -q -x -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -3 $O/pytest_gpu.log
probe() { DEEPFIT_B200_LIB=$1 timeout 300 python tools/scale_probe.py $2 2>&1 | tail -1 | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(l['cloud'], l['k'], 'morton' if l['morton'] else 'file', 'knn M q/s', round(l['knn_queries_per_s']/1e6,2), l['properties_ok'])"; }
for rep in 1 2; do
for v in "" ab_libs/lib_minb4.so ab_libs/lib_minb6.so ab_libs/lib_selg.so ab_libs/lib_bu2.so ab_libs/lib_bu8.so ab_libs/lib_minb6bu2.so; do
  echo "== lib=$v"
  probe "$v" "--points 1000000 --cloud torus --k 256 --order 3 --queries 1000000 --morton --chunk 262144"
  probe "$v" "--points 1000000 --cloud torus --k 256 --order 3 --queries 1000000 --chunk 262144"
  probe "$v" "--points 1000000 --cloud sphere --k 128 --order 3 --queries 1000000 --morton --chunk 262144"
done
done 2>&1 | tee $O/knn_ab.log
for v in "" ab_libs/lib_minb4.so ab_libs/lib_minb6.so; do
  echo "== lib=$v"
  probe "$v" "--points 30000000 --cloud density --k 256 --order 3 --queries 1000000 --morton"
  probe "$v" "--points 10000000 --cloud multi --k 512 --order 4 --queries 262144 --morton"
  probe "$v" "--points 1000000 --cloud sphere --k 64 --order 3 --queries 1000000 --morton --chunk 262144"
done 2>&1 | tee -a $O/knn_ab.log
