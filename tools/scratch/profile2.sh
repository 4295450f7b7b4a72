mkdir -p gpurun_out/prof2
python bench.py --steps 1 --warmup 1 --points 65536 --no-cpu-baseline > gpurun_out/prof2/bench_small.jsonl 2>&1; echo "small exit $?"
ncu --set full --clock-control none --import-source on -k regex:"knn_kernel|pca_kernel|wls_kernel" -s 3 -c 6 -o gpurun_out/prof2/stages -f python bench.py --steps 1 --warmup 1 --points 65536 --no-cpu-baseline > gpurun_out/prof2/ncu.log 2>&1; echo "ncu exit $?"
ncu -i gpurun_out/prof2/stages.ncu-rep --page raw --csv > gpurun_out/prof2/stages_raw.csv 2>/dev/null; echo "export exit $?"
rm -f gpurun_out/prof2/stages.ncu-rep
