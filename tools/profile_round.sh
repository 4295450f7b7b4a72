#!/bin/bash
# End-of-round capture on one B200 (run under gpurun from the repo root): the bench line, the reference arm, the ncu
# launch list and `--set full` captures of the network kernels and of the kNN / PCA / WLS kernels, the per-CTA timeline,
# the GPU tests, smoke, and the classic-path / k=512 probes.  Every ncu run repeats a command that has just exited 0
# without ncu.  Outputs land in gpurun_out/prof*/; the summaries worth keeping are copied into profiles/ by hand.
set -u
mkdir -p gpurun_out/prof gpurun_out/prof2
python bench.py --steps 3 --warmup 3 > gpurun_out/prof/bench.jsonl 2> gpurun_out/prof/bench.err; echo "bench exit $?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/prof/bench_ref.jsonl 2> gpurun_out/prof/bench_ref.err; echo "ref exit $?"
python bench.py --steps 1 --warmup 1 --points 32768 --no-cpu-baseline > gpurun_out/prof/bench_small.jsonl 2>&1; echo "small exit $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/prof/launches.csv python bench.py --steps 1 --warmup 1 --points 32768 --no-cpu-baseline > gpurun_out/prof/ncu1.log 2>&1; echo "ncu list exit $?"
ncu --set full --clock-control none --import-source on -k regex:"chain_max|head_fused" -s 8 -c 8 -o gpurun_out/prof/full -f python bench.py --steps 1 --warmup 1 --points 32768 --no-cpu-baseline > gpurun_out/prof/ncu2.log 2>&1; echo "ncu full exit $?"
ncu -i gpurun_out/prof/full.ncu-rep --page raw --csv > gpurun_out/prof/full_raw.csv 2>/dev/null; echo "export exit $?"
rm -f gpurun_out/prof/full.ncu-rep
python bench.py --steps 1 --warmup 1 --points 65536 --no-cpu-baseline > gpurun_out/prof2/bench_small.jsonl 2>&1; echo "small exit $?"
ncu --set full --clock-control none --import-source on -k regex:"knn_kernel|pca_kernel|wls_kernel" -s 3 -c 6 -o gpurun_out/prof2/stages -f python bench.py --steps 1 --warmup 1 --points 65536 --no-cpu-baseline > gpurun_out/prof2/ncu.log 2>&1; echo "ncu stages exit $?"
ncu -i gpurun_out/prof2/stages.ncu-rep --page raw --csv > gpurun_out/prof2/stages_raw.csv 2>/dev/null; echo "export exit $?"
rm -f gpurun_out/prof2/stages.ncu-rep
timeout 300 python tools/timeline.py > gpurun_out/prof/timeline.log 2>&1; echo "timeline exit $?"
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/prof/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/prof/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/prof/smoke.log 2>&1; tail -1 gpurun_out/prof/smoke.log
rm -f gpurun_out/prof/scale_probe.jsonl
for a in "--points 1000000 --cloud torus --k 256 --order 3 --queries 1000000" "--points 10000000 --cloud multi --k 256 --order 3 --queries 1000000" "--points 10000000 --cloud multi --k 512 --order 4 --queries 262144" "--points 30000000 --cloud density --k 256 --order 3 --queries 1000000"; do timeout 300 python tools/scale_probe.py $a 2>/dev/null | tail -1 >> gpurun_out/prof/scale_probe.jsonl; done
timeout 300 python tools/deepfit_probe.py 512 4 2000000 2>/dev/null | tail -1 >> gpurun_out/prof/scale_probe.jsonl
cut -c1-250 gpurun_out/prof/scale_probe.jsonl
