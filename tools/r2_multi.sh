#!/bin/bash
# round 2, multi-GPU session on ONE 8-GPU box: weak scaling of the default workload at N = 1, 2, 4, 8; strong scaling of
# BASELINE config 4 (10 M points, k = 512, order 4, DeepFit) at N = 2, 4, 8; config 5 (100 M points, variable density) at
# N = 8; bit-equality of the sharded + gathered result with the whole-cloud result at N = 8.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2m; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29540
run() {  # name N args...
  name=$1; n=$2; shift 2; port=$((port+1))
  if [ "$n" == "1" ]; then timeout 900 python bench.py --gpus 1 "$@" > $O/$name.jsonl 2> $O/$name.err
  else timeout 900 $TR --nproc-per-node $n --master-port $port bench.py --gpus $n "$@" > $O/$name.jsonl 2> $O/$name.err; fi
  echo "$name rc $?"; grep '^{' $O/$name.jsonl | cut -c1-260; tail -2 $O/$name.err
}
nvidia-smi -L | head -8
for n in 1 2 4 8; do run bench_weak_n$n $n --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline; done
timeout 600 $TR --nproc-per-node 8 --master-port 29590 tools/multi_gpu_check.py > $O/multi_gpu_check_n8.log 2>&1; echo "check rc $?"; tail -4 $O/multi_gpu_check_n8.log
for n in 8 4 2; do run bench_config4_strong_n$n $n --workload config4 --scaling strong --steps 1 --warmup 1 --no-cpu-baseline --no-library-baseline; done
run bench_config5_n8 8 --workload config5 --steps 1 --warmup 1 --no-cpu-baseline --no-library-baseline
run bench_config2_strong_n8 8 --workload config2 --scaling strong --steps 5 --warmup 3 --no-cpu-baseline
