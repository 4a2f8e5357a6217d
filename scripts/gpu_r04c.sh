#!/usr/bin/env bash
set -u
TAG=${1:-r04c}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
WL=${2:-sceneflow_576x960}
for t in 0 1; do
export DSSMD_PDL_PREFETCH_TLB=$t
echo "TLB=$t" | tee -a "$OUT/kbench_prefetch_tlb.log"
for k in corr_fwd corr_bwd; do
  timeout 300 python scripts/kbench.py $k $WL DSSMD_PDL_PREFETCH=0,1,-1 2>&1 | tee -a "$OUT/kbench_prefetch_tlb.log"
done
done
