#!/usr/bin/env bash
set -u
TAG=${1:-r02y}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -p no:cacheprovider -k "attention or sttm" 2>&1 | grep -v "^E  " | tail -12
python scripts/attn_err.py | tee "$OUT/attn_err.log"
DSSMD_TUNING=1 DSSMD_ATTN_FWD_MMA=0 DSSMD_ATTN_BWD_MMA=0 python scripts/attn_err.py | sed 's/^/FFMA kernels: /' | tee -a "$OUT/attn_err.log"
