#!/bin/bash
# round 2, run 12 (one B200): kNN build variants (block_cell as a function, cold-branch hints, register budgets), same box.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2q; mkdir -p $O
for v in "" ab_libs/lib_bc.so ab_libs/lib_ex.so ab_libs/lib_bcex.so ab_libs/lib_minb5.so ab_libs/lib_minb5bcex.so ab_libs/lib_minb3.so ""; do
  echo "== lib=$v"
  for a in "--points 1000000 --cloud torus --k 256 --order 3 --queries 1000000 --morton --chunk 262144" "--points 1000000 --cloud torus --k 256 --order 3 --queries 1000000 --chunk 262144" "--points 10000000 --cloud multi --k 512 --order 4 --queries 262144 --morton"; do
    DEEPFIT_B200_LIB=$v timeout 300 python tools/scale_probe.py $a 2>&1 | tail -1 | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(l['cloud'], l['k'], 'morton' if l['morton'] else 'file', 'knn M q/s', round(l['knn_queries_per_s']/1e6,2), l['stage_ms'], l['properties_ok'])"
  done
done 2>&1 | tee $O/knn_ab.log
