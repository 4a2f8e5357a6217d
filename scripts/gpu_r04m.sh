#!/usr/bin/env bash
set -u
TAG=${1:-r04m}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -k "warp or lrwarp or photometric or dispatch or back_to_back" > "$OUT/pytest_warp.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest_warp.log"; tail -n 8 "$OUT/pytest_warp.log" | cut -c1-250
timeout 300 python scripts/kbench.py warp_bwd all DSSMD_WARP_BWD_BAND_ROWS=- 2>&1 | tee -a "$OUT/kbench_warp_bwd.log"
