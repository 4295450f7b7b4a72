#!/bin/bash
# 2-GPU rehearsal of tools/r2_multi.sh at reduced sizes
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2ms; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-library-baseline > $O/weak_n2.jsonl 2> $O/weak_n2.err; echo "weak rc $?"; grep '^{' $O/weak_n2.jsonl | cut -c1-300; tail -3 $O/weak_n2.err
timeout 600 $TR --nproc-per-node 2 --master-port 29542 tools/multi_gpu_check.py > $O/check_n2.log 2>&1; echo "check rc $?"; tail -4 $O/check_n2.log
timeout 600 $TR --nproc-per-node 2 --master-port 29543 bench.py --gpus 2 --workload config5 --points 6000000 --steps 1 --warmup 1 --no-cpu-baseline --no-library-baseline > $O/config5_small_n2.jsonl 2> $O/config5_small_n2.err; echo "config5 rc $?"; grep '^{' $O/config5_small_n2.jsonl | cut -c1-300; tail -3 $O/config5_small_n2.err
timeout 600 $TR --nproc-per-node 2 --master-port 29544 bench.py --gpus 2 --workload config4 --points 1000000 --scaling strong --steps 1 --warmup 1 --no-cpu-baseline --no-library-baseline > $O/config4_small_n2.jsonl 2> $O/config4_small_n2.err; echo "config4 rc $?"; grep '^{' $O/config4_small_n2.jsonl | cut -c1-300; tail -3 $O/config4_small_n2.err
