#!/usr/bin/env bash
set -u
TAG=${1:-r02c}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -p no:cacheprovider -k "warp or photometric or dependent or lrwarp" > "$OUT/pytest_warp.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest_warp.log"; tail -n 8 "$OUT/pytest_warp.log"
timeout 600 python scripts/kbench.py warp_bwd all DSSMD_WARP_BWD_VARIANT=6,7 > "$OUT/kbench_warp_bwd.log" 2>&1; cat "$OUT/kbench_warp_bwd.log"
bash scripts/gpu_ncu.sh $TAG band2 warp_bwd
