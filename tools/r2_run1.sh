#!/bin/bash
# round 2, run 1: full GPU test suite of the new tree, then the default bench in both parity precisions
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench_bf16x3.jsonl 2> gpurun_out/r2a_bench_bf16x3.err; echo "bench bf16x3 rc $?"
timeout 300 python bench.py --steps 5 --warmup 3 --precision f16f8 > gpurun_out/r2a_bench_f16f8.jsonl 2> gpurun_out/r2a_bench_f16f8.err; echo "bench f16f8 rc $?"
cut -c1-600 gpurun_out/r2a_bench_bf16x3.jsonl; cut -c1-600 gpurun_out/r2a_bench_f16f8.jsonl
