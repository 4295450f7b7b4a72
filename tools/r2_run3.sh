#!/bin/bash
# round 2, run 3: the new parity precision f16x3 -- GPU tests, gate dump, default bench in each precision, timelines
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2c_pytest.log
tail -12 gpurun_out/r2c_pytest.log
timeout 300 python tools/dump_gate_outputs.py > gpurun_out/r2c_dump.log 2>&1; echo "dump rc $?"
for p in f16x3 bf16x3 f16f8; do
timeout 600 python bench.py --precision $p --steps 5 > gpurun_out/r2c_bench_$p.jsonl 2> gpurun_out/r2c_bench_$p.err; echo "bench $p rc $?"
cut -c1-300 gpurun_out/r2c_bench_$p.jsonl; tail -3 gpurun_out/r2c_bench_$p.err
done
for p in f16x3 f16f8; do timeout 300 python tools/timeline.py --precision $p --batch 16384 --cta 2000 > gpurun_out/r2c_timeline_$p.log 2>&1; echo "timeline $p rc $?"; done
