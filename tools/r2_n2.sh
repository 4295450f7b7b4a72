#!/bin/bash
# round 2, final tree on TWO GPUs of one box: sharded + all-gathered == whole-cloud bit equality, the default bench line at
# N = 2 exactly as the driver launches it.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29561 tools/multi_gpu_check.py > $O/multi_gpu_check_n2.log 2>&1; echo "check rc $?"; tail -4 $O/multi_gpu_check_n2.log
timeout 400 $TR --nproc-per-node 2 --master-port 29562 bench.py --gpus 2 --steps 5 --warmup 3 --no-library-baseline > $O/bench_weak_n2.jsonl 2> $O/bench_weak_n2.err; echo "bench n2 rc $?"; grep '^{' $O/bench_weak_n2.jsonl | cut -c1-300


