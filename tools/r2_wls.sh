#!/bin/bash
# WLS A/B on one B200: parity tests, then the stage alone for each library variant
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2g; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_pca_wls.py -q -x -p no:cacheprovider > $O/pytest_wls.log 2>&1; echo "pytest wls rc $?"; tail -3 $O/pytest_wls.log
for v in "" $VARIANTS; do
  echo "== lib=$v"
  DEEPFIT_B200_LIB=$v timeout 300 python tools/wls_probe.py --n 65536 262144 --orders 3 4 2>&1 | tail -8
done 2>&1 | tee $O/wls_ab.log
