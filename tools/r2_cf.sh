#!/bin/bash
# round 2: "corrections first" product order in the generic GEMM kernel (FC layers): unit tests, the weight noise on the
# golden rows with and without it, the other-weight-set parity tests, throughput A/B on one box.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2C; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_network.py -m gpu -q -p no:cacheprovider -x > $O/pytest_network.log 2>&1; echo "pytest network rc $?"; tail -3 $O/pytest_network.log
DFB_DBG=11:0 python tools/mix_probe.py f16x3 bf16x3 2>&1 | cut -c1-330 | tee $O/probe_cf0.log
python tools/mix_probe.py f16x3 bf16x3 2>&1 | cut -c1-330 | tee $O/probe_cf1.log
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -p no:cacheprovider -s -k "other_weight_sets or pipeline_vs_oracle" 2>&1 | grep -E "seed|angle max|passed|failed|Error" | cut -c1-260 | tee $O/pytest_seeds.log
B="--steps 5 --no-cpu-baseline --no-library-baseline"
for rep in 1 2; do
DFB_DBG=11:0 timeout 300 python bench.py $B --allow-dbg 2>/dev/null | python -c "import sys,json; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cf=0', round(l['value']), l['kernel_ms'])"
timeout 300 python bench.py $B 2>/dev/null | python -c "import sys,json; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cf=1', round(l['value']), l['kernel_ms'])"
done 2>&1 | tee $O/bench_ab.log
