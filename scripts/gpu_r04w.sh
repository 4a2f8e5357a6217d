#!/usr/bin/env bash
# corr_fwd_mma32 with the cross terms on bf16 MMAs: parity tests, timing against the 3xTF32 build
set -u
TAG=${1:-r04w}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -k "corr or cost_volume or dispatch" > "$OUT/pytest_corr.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest_corr.log"; tail -n 12 "$OUT/pytest_corr.log" | cut -c1-300
for w in corr_micro kitti_384x1280 train_320x640; do
  for k in ${2:-corr_fwd}; do
  timeout 300 python scripts/kbench.py $k $w DSSMD_MMA32_BF16X=0,1 2>&1 | tee -a "$OUT/kbench_bf16x.log"
  done
done
