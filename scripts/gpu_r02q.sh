#!/usr/bin/env bash
set -u
TAG=${1:-r02q}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 -p no:cacheprovider -x -k "tcgen05_row_tiles" > "$OUT/pytest_umma2.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest_umma2.log"; tail -n 12 "$OUT/pytest_umma2.log" | cut -c1-300
timeout 300 python scripts/kbench.py corr_fwd all DSSMD_CORR_FWD_VARIANT=-,4 > "$OUT/kbench_fwd.log" 2>&1; cat "$OUT/kbench_fwd.log"
DSSMD_UMMA2_N32=1 timeout 300 python scripts/kbench.py corr_fwd corr_micro,kitti_384x1280 DSSMD_CORR_FWD_VARIANT=4 > "$OUT/kbench_fwd_n32.log" 2>&1; cat "$OUT/kbench_fwd_n32.log"
timeout 300 python scripts/kbench.py corr_fwd corr_micro DSSMD_UMMA2_STAGES=3,4,5,6,7 > /dev/null 2>&1
for st in 3 5 7; do DSSMD_CORR_FWD_VARIANT=4 DSSMD_UMMA2_STAGES=$st timeout 200 python scripts/kbench.py corr_fwd corr_micro DSSMD_UMMA2_N32=0 2>&1 | grep "N32" | sed "s/^/stages=$st /"; done
