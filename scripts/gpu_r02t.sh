#!/usr/bin/env bash
# corr_bwd_wide.cu: parity tests, then the per-kernel timing against the fused kernel on the model shapes
set -u
TAG=${1:-r02t}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 -p no:cacheprovider -x -k "corr_bwd_wide" > "$OUT/pytest_wide.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest_wide.log"; tail -n 12 "$OUT/pytest_wide.log" | cut -c1-300
timeout 300 python scripts/kbench.py corr_bwd sceneflow_576x960,train_320x640,kitti_384x1280 DSSMD_CORR_BWD_WIDE=0,1,2 > "$OUT/kbench_bwd.log" 2>&1; cat "$OUT/kbench_bwd.log"
for st in 2 3 4; do DSSMD_CORR_BWD_WIDE_STAGES=$st timeout 200 python scripts/kbench.py corr_bwd sceneflow_576x960 DSSMD_CORR_BWD_WIDE_CPW=8,4 2>&1 | grep "CPW" | sed "s/^/stages=$st /"; done | tee "$OUT/kbench_bwd_stages.log"
DSSMD_CORR_BWD_WIDE_STAGES=8 timeout 200 python scripts/kbench.py corr_bwd sceneflow_576x960 DSSMD_CORR_BWD_WIDE_CPW=4 2>&1 | grep "CPW" | sed "s/^/stages=8 /" | tee -a "$OUT/kbench_bwd_stages.log"
