#!/bin/bash
# round 2, run 11 (one B200): kNN final step v2 (rolled counting sort by distance bucket, helpers as real functions):
# tests, then same-box A/B of five builds of knn.cu.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2p; mkdir -p $O
timeout 1800 python -m pytest tests/test_gpu_knn.py tests/test_gpu_pipeline.py tests/test_gpu_config3.py -m gpu -q -x -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -4 $O/pytest_gpu.log
for v in "" ab_libs/lib_sortnet.so ab_libs/lib_bucket4.so ab_libs/lib_inl.so ab_libs/lib_sortnet_inl.so; do
  echo "== lib=$v"
  for a in "--points 1000000 --cloud torus --k 256 --order 3 --queries 1000000 --morton --chunk 262144" "--points 1000000 --cloud torus --k 256 --order 3 --queries 1000000 --chunk 262144" "--points 1000000 --cloud sphere --k 128 --order 3 --queries 1000000 --morton --chunk 262144" "--points 10000000 --cloud multi --k 512 --order 4 --queries 262144 --morton"; do
    DEEPFIT_B200_LIB=$v timeout 300 python tools/scale_probe.py $a 2>&1 | tail -1 | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(l['cloud'], l['k'], 'morton' if l['morton'] else 'file', 'knn M q/s', round(l['knn_queries_per_s']/1e6,2), l['stage_ms'], l['properties_ok'])"
  done
done 2>&1 | tee $O/knn_ab.log
for v in "" ab_libs/lib_sortnet_inl.so; do
  DEEPFIT_B200_LIB=$v timeout 300 python tools/scale_probe.py --points 30000000 --cloud density --k 256 --order 3 --queries 1000000 --morton 2>&1 | tail -1 | cut -c1-330
done 2>&1 | tee -a $O/knn_ab.log
summ() { python -c "import sys,json; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(l['value']), round(l['ms_per_step'],3), l.get('stages_one_chunk'))"; }
for cfg in "cur:" "sortnet:ab_libs/lib_sortnet.so" "sortnet_inl:ab_libs/lib_sortnet_inl.so" "cur:" "sortnet:ab_libs/lib_sortnet.so" "sortnet_inl:ab_libs/lib_sortnet_inl.so"; do
  name=${cfg%%:*}; lib=${cfg#*:}
  DEEPFIT_B200_LIB=$lib timeout 300 python bench.py --workload config2 --order 3 --steps 5 --no-cpu-baseline --no-library-baseline --no-check 2> $O/c2_$name.err | tail -1 > $O/c2_${name}.jsonl
  summ "config2 o3 $name" < $O/c2_${name}.jsonl
done 2>&1 | tee $O/config2_ab.log
