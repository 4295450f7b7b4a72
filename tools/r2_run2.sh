#!/bin/bash
# round 2, run 2: GPU tests (no -x), gate diagnostics, default bench, config4 / config2 / boxy bench lines
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2b_pytest.log
tail -12 gpurun_out/r2b_pytest.log
timeout 300 python tools/dump_gate_outputs.py > gpurun_out/r2b_dump.log 2>&1; echo "dump rc $?"
timeout 600 python bench.py > gpurun_out/r2b_bench_default.jsonl 2> gpurun_out/r2b_bench_default.err; echo "bench default rc $?"
timeout 600 python bench.py --workload boxy --steps 5 > gpurun_out/r2b_bench_boxy.jsonl 2> gpurun_out/r2b_bench_boxy.err; echo "bench boxy rc $?"
timeout 600 python bench.py --workload config2 --steps 3 > gpurun_out/r2b_bench_config2.jsonl 2> gpurun_out/r2b_bench_config2.err; echo "bench config2 rc $?"
timeout 900 python bench.py --workload config4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2b_bench_config4.jsonl 2> gpurun_out/r2b_bench_config4.err; echo "bench config4 rc $?"
for f in default boxy config2 config4; do cut -c1-400 gpurun_out/r2b_bench_$f.jsonl; tail -3 gpurun_out/r2b_bench_$f.err; done
