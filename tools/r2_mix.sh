#!/bin/bash
# round 2, per-layer precision ladder on the GPU (one B200): which chain layers can run with the FP16 hi*hi product only
# (f16mix) and still hold every per-row gate of the 512 golden patches per cloud, and what each candidate is worth.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2M; mkdir -p $O
B="--steps 5 --no-cpu-baseline --no-library-baseline"
line() { python -c "import sys,json; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=l['roofline']; print('$1', round(l['value']), 'frac %.3f' % r['frac'], 'exec %.3f' % r['executed_frac'], l['kernel_ms'], l['clocks']['sm_mhz'], l['check'])"; }
python tools/mix_probe.py f16x3 f16mix:qstn_max f16mix:stn_max f16mix:enc_max f16mix:qstn_max+stn_max \
  f16mix:qstn_max+stn_max+qstn_small+stn_small f16mix:qstn_small+stn_small+enc_small f16mix:qstn_max+stn_max+enc_max \
  f16mix:qstn_max+stn_max+enc_max+qstn_small+stn_small+enc_small bf16x3 f16f8 > $O/mix_probe.log 2>&1; echo "probe rc $?"; cat $O/mix_probe.log
for rep in 1 2; do
DEEPFIT_B200_LIB=ab_libs/lib_prev.so timeout 300 python bench.py $B 2>/dev/null | line "prev-lib f16x3"
timeout 300 python bench.py $B 2>/dev/null | line "new-lib f16x3"
for p in f16mix:qstn_max f16mix:qstn_max+stn_max f16mix:qstn_max+stn_max+qstn_small+stn_small f16mix:qstn_max+stn_max+enc_max; do
  timeout 300 python bench.py $B --precision $p 2>/dev/null | line "$p"
done
done 2>&1 | tee $O/mix_bench.log
for a in f16mix:qstn_max+stn_max f16mix:qstn_max+stn_max+qstn_small+stn_small; do
  DFB_EXPERIMENT_ALIAS="f16x3=$a" timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -rf --tb=line -k "not c_client" > $O/pytest_alias_${a//[:+]/_}.log 2>&1; echo "pytest under alias $a rc $?"; tail -25 $O/pytest_alias_${a//[:+]/_}.log
done
