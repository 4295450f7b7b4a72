#!/bin/bash
# Same-box A/B of two builds of the library (boxes differ by +-3 %, so only same-box comparisons count):
#   build the variant to compare against into ab_libs/lib_prev.so (same C ABI), then under gpurun
#   bash tools/ab_bench.sh            alternates the in-tree build and the variant twice and prints value + kernel classes
# DEEPFIT_B200_LIB selects the library the Python binding loads.
for v in "" "ab_libs/lib_prev.so" "" "ab_libs/lib_prev.so"; do
  DEEPFIT_B200_LIB=$v timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; l=json.loads(sys.stdin.read()); print('lib=$v', round(l['value']), l['kernel_ms'])"
done
