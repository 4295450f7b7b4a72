#!/bin/bash
# WLS: TMA-staged input ring (DFB_WLS_STAGED=1) vs register loads: parity tests under the staged kernel, then the stage alone
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2i; mkdir -p $O
DFB_WLS_STAGED=1 timeout 900 python -m pytest tests/test_gpu_pca_wls.py -q -x -p no:cacheprovider > $O/pytest_wls_staged.log 2>&1; echo "pytest wls staged rc $?"; tail -3 $O/pytest_wls_staged.log
for s in 0 1; do
  echo "== DFB_WLS_STAGED=$s"
  DFB_WLS_STAGED=$s timeout 300 python tools/wls_probe.py --n 16384 65536 262144 --orders 1 2 3 4 2>&1 | tail -24
done 2>&1 | tee $O/wls_staged_ab.log
