#!/usr/bin/env bash
set -u
TAG=${1:-r04e}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
for WL in corr_micro kitti_384x1280 train_320x640; do
for k in corr_fwd corr_bwd; do
  timeout 300 python scripts/kbench.py $k $WL DSSMD_PDL_PREFETCH=0,- 2>&1 | tee -a "$OUT/kbench_prefetch.log"
done
done
