#!/usr/bin/env bash
set -u
TAG=${1:-r02d}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 600 python scripts/kbench.py warp_bwd all DSSMD_WARP_BWD_VARIANT=6,7 > "$OUT/kbench_warp_bwd.log" 2>&1; cat "$OUT/kbench_warp_bwd.log"
bash scripts/gpu_ncu.sh $TAG band2 warp_bwd
