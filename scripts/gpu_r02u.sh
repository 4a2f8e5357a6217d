#!/usr/bin/env bash
set -u
TAG=${1:-r02u}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -p no:cacheprovider -x -k "corr_bwd_wide or default_dispatch or fresh_threads or corr_vs_oracle_f32 or cat" 2>&1 | tail -3
timeout 300 python scripts/kbench.py corr_bwd all DSSMD_CORR_BWD_WIDE=1,0 2>&1 | tee "$OUT/wide_final.log"
export DSSMD_TUNING=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:corr_fwd -s 12 -c 1 -o "$OUT/fwd_reg" -f ./scripts/micro/abi_loop 8 > "$OUT/ncu_fwd.log" 2>&1; tail -1 "$OUT/ncu_fwd.log"
