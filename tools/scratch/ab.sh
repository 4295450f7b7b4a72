mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_network.py -m gpu -x -q -p no:cacheprovider --timeout 300 2>&1 | tail -3
timeout 300 python tools/timeline.py > gpurun_out/timeline10.log 2>&1; echo "exit $?" >> gpurun_out/timeline10.log
for v in "" "ab_libs/lib_prev.so" "" "ab_libs/lib_prev.so"; do DEEPFIT_B200_LIB=$v timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; l=json.loads(sys.stdin.read()); print('lib=$v', round(l['value']), l['kernel_ms'])"; done
