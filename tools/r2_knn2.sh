#!/bin/bash
# kNN with sub-grids under heavy cells: exactness tests, then stage timings (file order / Morton order) on the uniform,
# multi-surface and variable-density clouds; $VARIANTS = older builds of the library for a same-box A/B
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2j; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_knn.py -q -x -p no:cacheprovider > $O/pytest_knn.log 2>&1; echo "pytest knn rc $?"; tail -5 $O/pytest_knn.log
for v in "" $VARIANTS; do
  echo "== lib=$v"
  for m in "" "--morton"; do
  for a in "--points 1000000 --cloud torus --k 256 --order 3 --queries 1000000" "--points 10000000 --cloud multi --k 512 --order 4 --queries 262144" "--points 30000000 --cloud density --k 256 --order 3 --queries 1000000"; do
    DEEPFIT_B200_LIB=$v timeout 300 python tools/scale_probe.py $a $m 2>&1 | tail -1 | cut -c1-420
  done
  done
done 2>&1 | tee $O/knn_ab.log
