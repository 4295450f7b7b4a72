#!/bin/bash
# round 2, run 10 (one B200): kNN final step as a counting sort by distance bucket instead of the 36-stage network:
# kNN / pipeline tests, then same-box A/B (lib_sortnet = the same tree with -DDFB_KNN_SORTNET).
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2o; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -x -p no:cacheprovider --durations=5 > $O/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -12 $O/pytest_gpu.log
summ() { python -c "import sys,json; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(l['value']), round(l['ms_per_step'],3), l.get('stages_one_chunk'), l.get('check'))"; }
for rep in 1 2; do
for cfg in "cur:" "sortnet:ab_libs/lib_sortnet.so"; do
  name=${cfg%%:*}; lib=${cfg#*:}
  DEEPFIT_B200_LIB=$lib timeout 300 python bench.py --workload config2 --order 3 --steps 5 --no-cpu-baseline --no-library-baseline 2> $O/c2_$name.err | tail -1 > $O/c2_${name}_$rep.jsonl
  summ "config2 o3 $name" < $O/c2_${name}_$rep.jsonl
done
done 2>&1 | tee $O/config2_ab.log
for v in "" ab_libs/lib_sortnet.so; do
  echo "== lib=$v"
  for a in "--points 1000000 --cloud torus --k 256 --order 3 --queries 1000000 --morton" "--points 30000000 --cloud density --k 256 --order 3 --queries 1000000 --morton" "--points 10000000 --cloud multi --k 512 --order 4 --queries 262144 --morton" "--points 1000000 --cloud torus --k 256 --order 3 --queries 1000000" "--points 1000000 --cloud sphere --k 64 --order 3 --queries 1000000 --morton" "--points 1000000 --cloud sphere --k 128 --order 3 --queries 1000000 --morton"; do
    DEEPFIT_B200_LIB=$v timeout 300 python tools/scale_probe.py $a 2>&1 | tail -1 | cut -c1-420
  done
done 2>&1 | tee $O/knn_ab.log
for cfg in "cur:" "sortnet:ab_libs/lib_sortnet.so"; do
  name=${cfg%%:*}; lib=${cfg#*:}
  DEEPFIT_B200_LIB=$lib timeout 400 python bench.py --steps 5 --no-cpu-baseline --no-library-baseline 2> $O/def_$name.err | tail -1 > $O/def_$name.jsonl
  summ "default $name" < $O/def_$name.jsonl
done 2>&1 | tee $O/default_ab.log
