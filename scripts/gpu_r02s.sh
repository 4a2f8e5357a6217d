#!/usr/bin/env bash
set -u
TAG=${1:-r02s}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 -p no:cacheprovider -x -k "tcgen05_row_tiles" 2>&1 | tail -2
export DSSMD_CORR_FWD_VARIANT=4
for dbg in 0 1 2 4 8 3 6 7 14 15; do DSSMD_UMMA2_DBG=$dbg timeout 200 python scripts/kbench.py corr_fwd corr_micro,sceneflow_576x960 DSSMD_UMMA2_N32=0 2>&1 | grep "N32" | tr '\n' ' ' | sed "s/^/dbg=$dbg /"; echo; done | tee "$OUT/umma2_stage_ablation.log"
