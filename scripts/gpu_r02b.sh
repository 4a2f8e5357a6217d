#!/usr/bin/env bash
# r02b: the rewritten band backward (warp_bwd_band2.cu): warp tests, then per-call timings of variants 6 / 7 / 7 with 3 blocks per SM
set -u
TAG=${1:-r02b}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -p no:cacheprovider -k "warp or photometric or dependent or lrwarp" > "$OUT/pytest_warp.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest_warp.log"; tail -n 15 "$OUT/pytest_warp.log"
timeout 600 python scripts/kbench.py warp_bwd all DSSMD_WARP_BWD_VARIANT=6,7 > "$OUT/kbench_warp_bwd.log" 2>&1; cat "$OUT/kbench_warp_bwd.log"
DSSMD_WARP_BWD_BAND_OCC=3 timeout 600 python scripts/kbench.py warp_bwd all DSSMD_WARP_BWD_VARIANT=7 > "$OUT/kbench_warp_bwd_occ3.log" 2>&1; cat "$OUT/kbench_warp_bwd_occ3.log"
