mkdir -p gpurun_out/prof
cd $GRAFT_REPO_ROOT
python bench.py --steps 3 --warmup 3 > gpurun_out/prof/bench.jsonl 2> gpurun_out/prof/bench.err; echo "bench exit $?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/prof/bench_ref.jsonl 2> gpurun_out/prof/bench_ref.err; echo "ref exit $?"
python bench.py --steps 1 --warmup 1 --points 32768 --no-cpu-baseline > gpurun_out/prof/bench_small.jsonl 2>&1; echo "small exit $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/prof/launches.csv python bench.py --steps 1 --warmup 1 --points 32768 --no-cpu-baseline > gpurun_out/prof/ncu1.log 2>&1; echo "ncu list exit $?"
ncu --set full --clock-control none --import-source on -k regex:"chain_max|head_fused" -s 8 -c 8 -o gpurun_out/prof/full -f python bench.py --steps 1 --warmup 1 --points 32768 --no-cpu-baseline > gpurun_out/prof/ncu2.log 2>&1; echo "ncu full exit $?"
ncu -i gpurun_out/prof/full.ncu-rep --page raw --csv > gpurun_out/prof/full_raw.csv 2>/dev/null; echo "export exit $?"
ls -la gpurun_out/prof
