#!/bin/bash
# round 2, run 6: Morton-ordered pipeline + WLS v2: full GPU tests, then the classic / DeepFit bench lines
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2h; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -5 $O/pytest_gpu.log
for o in 3 4; do timeout 600 python bench.py --workload config2 --order $o --steps 3 --no-cpu-baseline > $O/bench_config2_o$o.jsonl 2> $O/bench_config2_o$o.err; echo "bench config2 o$o rc $?"; cut -c1-200 $O/bench_config2_o$o.jsonl; tail -3 $O/bench_config2_o$o.err; done
timeout 600 python bench.py --steps 5 --no-cpu-baseline > $O/bench_default.jsonl 2> $O/bench_default.err; echo "bench default rc $?"; cut -c1-200 $O/bench_default.jsonl; tail -3 $O/bench_default.err
timeout 600 python bench.py --workload boxy --steps 5 --no-cpu-baseline > $O/bench_boxy.jsonl 2> $O/bench_boxy.err; echo "bench boxy rc $?"; cut -c1-200 $O/bench_boxy.jsonl
