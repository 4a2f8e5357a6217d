#!/usr/bin/env bash
set -u
TAG=${1:-r04g}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -k "warp or lrwarp or photometric or dispatch" > "$OUT/pytest_warp.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest_warp.log"; tail -n 4 "$OUT/pytest_warp.log" | cut -c1-200
timeout 300 python scripts/kbench.py warp_bwd train_320x640,corr_micro DSSMD_WARP_BWD_BAND_ROWS=- 2>&1 | tee -a "$OUT/kbench_train_warp_bwd.log"
timeout 300 python scripts/kbench.py warp_bwd sceneflow_576x960 DSSMD_WARP_BWD_BAND_OCC=-,3 2>&1 | tee -a "$OUT/kbench_train_warp_bwd.log"
