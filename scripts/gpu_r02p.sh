#!/usr/bin/env bash
# r02p: tcgen05 (TMA + UMMA 3xTF32) vs mma.sync 3xTF32 correlation forward at cfg2 / KITTI: timings and one full ncu capture each
set -u
TAG=${1:-r02p}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 300 python scripts/umma_time.py > "$OUT/umma_time.log" 2>&1; cat "$OUT/umma_time.log"
timeout 300 python scripts/kbench.py corr_fwd corr_micro,kitti_384x1280 DSSMD_CORR_FWD_VARIANT=-,3,2,1 > "$OUT/kbench_fwd.log" 2>&1; cat "$OUT/kbench_fwd.log"
export DSSMD_TUNING=1
bash scripts/gpu_ncu.sh ${TAG}_mma32 mma32 corr_fwd corr_micro
DSSMD_CORR_FWD_VARIANT=2 DSSMD_UMMA_SKIP=16 bash scripts/gpu_ncu.sh ${TAG}_umma umma corr_fwd corr_micro
