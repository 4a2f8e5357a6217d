#!/usr/bin/env bash
set -u
TAG=${1:-r04k}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -k "corr or dispatch or cost_volume or back_to_back" -x > "$OUT/pytest_corr.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest_corr.log"; tail -n 4 "$OUT/pytest_corr.log" | cut -c1-200
for k in corr_fwd corr_bwd; do
timeout 300 python scripts/kbench.py $k sceneflow_576x960,train_320x640 DSSMD_PDL_PREFETCH=-,0,2 2>&1 | tee -a "$OUT/kbench.log"
done
