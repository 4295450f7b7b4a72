#!/bin/bash
# round 2, last capture of the final tree on one B200 (a shortened tools/r2_final.sh: the GPU budget left is ~20 min):
# GPU tests, smoke, default bench line (with CPU + library baselines), ncu launch list + --set full of the network kernels,
# reference arm, config 1 (boxy), config 2 orders 1-4, ncu --set full of kNN / PCA / WLS, timeline, config 4 on one GPU.
# Every ncu run repeats a command that has just exited 0 without ncu.  Most important first: the call may be cut short.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2L; mkdir -p $O
B="--no-cpu-baseline --no-library-baseline"
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -12 $O/pytest_gpu.log
python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc $?"; tail -1 $O/smoke.log
timeout 600 python bench.py --steps 10 > $O/bench_default.jsonl 2> $O/bench_default.err; echo "bench default rc $?"
timeout 300 python bench.py --steps 10 --precision f16mix $B > $O/bench_f16mix.jsonl 2> $O/bench_f16mix.err; echo "bench f16mix rc $?"
python bench.py --steps 1 --warmup 1 --points 32768 $B > $O/bench_small.jsonl 2>&1; echo "small rc $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 1 --points 32768 $B > $O/ncu1.log 2>&1; echo "ncu list rc $?"
ncu --set full --clock-control none --import-source on -k regex:"chain_max|head_fused" -s 8 -c 8 -o $O/full -f python bench.py --steps 1 --warmup 1 --points 32768 $B > $O/ncu2.log 2>&1; echo "ncu full rc $?"
ncu -i $O/full.ncu-rep --page raw --csv > $O/full_raw.csv 2>/dev/null; echo "export rc $?"
rm -f $O/full.ncu-rep
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.jsonl 2> $O/bench_reference_arm.err; echo "ref rc $?"
for o in 3 1 2 4; do timeout 300 python bench.py --workload config2 --order $o --steps 3 --no-cpu-baseline > $O/bench_config2_o$o.jsonl 2> $O/bench_config2_o$o.err; echo "bench config2 o$o rc $?"; done
timeout 300 python bench.py --workload boxy --steps 5 $B > $O/bench_boxy.jsonl 2> $O/bench_boxy.err; echo "bench boxy rc $?"
python bench.py --workload config2 --steps 1 --warmup 1 --points 262144 --no-cpu-baseline > $O/bench_small_classic.jsonl 2>&1; echo "small classic rc $?"
ncu --set full --clock-control none --import-source on -k regex:"knn_kernel|pca_kernel|pca_eig_kernel|wls_kernel" -s 3 -c 8 -o $O/stages -f python bench.py --workload config2 --steps 1 --warmup 1 --points 262144 --no-cpu-baseline > $O/ncu3.log 2>&1; echo "ncu stages rc $?"
ncu -i $O/stages.ncu-rep --page raw --csv > $O/stages_raw.csv 2>/dev/null; echo "export rc $?"
rm -f $O/stages.ncu-rep
timeout 200 python tools/timeline.py --precision f16x3 --batch 16384 --cta 2000 > $O/timeline_f16x3.log 2>&1; echo "timeline rc $?"
timeout 500 python bench.py --workload config4 --steps 1 --warmup 1 $B > $O/bench_config4.jsonl 2> $O/bench_config4.err; echo "bench config4 rc $?"
for f in default f16mix boxy config2_o1 config2_o2 config2_o3 config2_o4 config4; do cut -c1-200 $O/bench_$f.jsonl; tail -1 $O/bench_$f.err; done
