#!/usr/bin/env bash
set -u
TAG=${1:-r02g}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 600 python scripts/kbench.py warp_bwd all DSSMD_WARP_BWD_VARIANT=6,7 > "$OUT/kbench_warp_bwd.log" 2>&1; cat "$OUT/kbench_warp_bwd.log"
timeout 600 python scripts/kbench.py warp_bwd sceneflow_576x960,kitti_384x1280 DSSMD_WARP_BWD_BAND_ROWS=4,6,8,9,12,16,24 > "$OUT/kbench_rows.log" 2>&1; cat "$OUT/kbench_rows.log"
bash scripts/gpu_ncu.sh $TAG band2 warp_bwd
