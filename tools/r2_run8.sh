#!/bin/bash
# round 2, run 8 (one B200): split stage 2 for the classic path (gather + covariance | one-thread-per-patch eigen-solve |
# rotation inside the jet fit), float4 gathers, all loads of a patch in flight: full GPU tests, then same-box A/B.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2m; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -x -p no:cacheprovider --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -16 $O/pytest_gpu.log
summ() { python -c "import sys,json; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(l['value']), round(l['ms_per_step'],3), l.get('stages_one_chunk'), l.get('check'))"; }
for rep in 1 2; do
for cfg in "cur::" "fusedpca:DFB_PIPE_FUSED_PCA=1:" "gather3:DFB_PIPE_GATHER3=1:" "fused_gather3:DFB_PIPE_FUSED_PCA=1 DFB_PIPE_GATHER3=1:" "prev::ab_libs/lib_prev.so"; do
  name=${cfg%%:*}; rest=${cfg#*:}; envs=${rest%%:*}; lib=${rest#*:}
  env $envs DEEPFIT_B200_LIB=$lib timeout 300 python bench.py --workload config2 --order 3 --steps 5 --no-cpu-baseline --no-library-baseline 2> $O/c2_$name.err | tail -1 > $O/c2_${name}_$rep.jsonl
  summ "config2 o3 $name" < $O/c2_${name}_$rep.jsonl
done
done 2>&1 | tee $O/config2_ab.log
for o in 1 4; do
  for cfg in "cur:" "prev:ab_libs/lib_prev.so"; do
    name=${cfg%%:*}; lib=${cfg#*:}
    DEEPFIT_B200_LIB=$lib timeout 300 python bench.py --workload config2 --order $o --steps 5 --no-cpu-baseline --no-library-baseline 2>> $O/c2_o.err | tail -1 > $O/c2_o${o}_$name.jsonl
    summ "config2 o$o $name" < $O/c2_o${o}_$name.jsonl
  done
done 2>&1 | tee -a $O/config2_ab.log
for cfg in "cur:" "prev:ab_libs/lib_prev.so"; do
  name=${cfg%%:*}; lib=${cfg#*:}
  DEEPFIT_B200_LIB=$lib timeout 400 python bench.py --steps 5 --no-cpu-baseline --no-library-baseline 2> $O/def_$name.err | tail -1 > $O/def_$name.jsonl
  summ "default $name" < $O/def_$name.jsonl
done 2>&1 | tee $O/default_ab.log
