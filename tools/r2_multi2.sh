#!/bin/bash
# round 2, second 8-GPU session (after the kNN sub-grids): config 5 (100 M points, variable density) at N = 8, the
# sharded == whole-cloud bit-equality check at N = 8, classic config 2 strong-scaled at N = 8
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2n; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29551 bench.py --gpus 8 --workload config5 --steps 1 --warmup 1 --no-cpu-baseline --no-library-baseline > $O/bench_config5_n8.jsonl 2> $O/bench_config5_n8.err; echo "config5 rc $?"; grep '^{' $O/bench_config5_n8.jsonl | cut -c1-260; tail -2 $O/bench_config5_n8.err
timeout 600 $TR --nproc-per-node 8 --master-port 29552 tools/multi_gpu_check.py > $O/multi_gpu_check_n8.log 2>&1; echo "check rc $?"; tail -4 $O/multi_gpu_check_n8.log
timeout 600 $TR --nproc-per-node 8 --master-port 29553 bench.py --gpus 8 --workload config2 --scaling strong --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_config2_strong_n8.jsonl 2> $O/bench_config2_strong_n8.err; echo "config2 rc $?"; grep '^{' $O/bench_config2_strong_n8.jsonl | cut -c1-260
