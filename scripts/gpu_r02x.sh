#!/usr/bin/env bash
set -u
TAG=${1:-r02x}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -p no:cacheprovider -x -k "asm or recovered or loss" 2>&1 | tail -3
timeout 200 python scripts/asm_host_cost.py 2>&1 | tail -1 | tee "$OUT/asm_host_cost.json"
