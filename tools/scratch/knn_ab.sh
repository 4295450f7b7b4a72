mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knn.py -m gpu -x -q -p no:cacheprovider --timeout 600 2>&1 | tail -3
for v in "" "ab_libs/lib_prev.so"; do DEEPFIT_B200_LIB=$v timeout 300 python tools/scale_probe.py --points 1000000 --cloud torus --k 256 --order 3 --queries 1000000 2>/dev/null | tail -1 | cut -c1-330; done
for v in "" "ab_libs/lib_prev.so"; do DEEPFIT_B200_LIB=$v timeout 300 python tools/scale_probe.py --points 10000000 --cloud multi --k 512 --order 4 --queries 262144 2>/dev/null | tail -1 | cut -c1-330; done
for v in "" "ab_libs/lib_prev.so"; do DEEPFIT_B200_LIB=$v timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; l=json.loads(sys.stdin.read()); print('lib=$v', round(l['value']), l['stages_one_chunk'])"; done
