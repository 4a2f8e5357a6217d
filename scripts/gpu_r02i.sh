#!/usr/bin/env bash
# r02i: fused correlation backward: correlation tests, per-call timings fused vs two launches on every workload, ncu capture
set -u
TAG=${1:-r02i}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -p no:cacheprovider -k "corr or fresh or dependent" > "$OUT/pytest_corr.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest_corr.log"; tail -n 8 "$OUT/pytest_corr.log"
timeout 600 python scripts/kbench.py corr_bwd all > "$OUT/kbench_corr_bwd.log" 2>&1; cat "$OUT/kbench_corr_bwd.log"
DSSMD_CORR_DEBUG=1 timeout 100 python scripts/one_kernel.py corr_bwd sceneflow_576x960 1 2>&1 | grep cfg | head -3
bash scripts/gpu_ncu.sh $TAG corr_bwd_fused corr_bwd
