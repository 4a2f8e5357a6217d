#!/usr/bin/env bash
set -u
TAG=${1:-r02y}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -p no:cacheprovider -k "attention or sttm" 2>&1 | grep -v "^E  " | tail -8
timeout 600 python scripts/attn_bench.py 2>&1 | grep fused | tee "$OUT/attn_bench_umma.log"
DSSMD_TUNING=1 DSSMD_ATTN_FWD_UMMA=0 timeout 600 python scripts/attn_bench.py 2>&1 | grep fused | sed 's/^/mma.sync: /' | tee -a "$OUT/attn_bench_umma.log"
