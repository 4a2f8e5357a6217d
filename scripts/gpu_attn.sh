#!/usr/bin/env bash
set -u
TAG=${1:-attn}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -k "attention or sttm or fresh_threads" > "$OUT/pytest_attn.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest_attn.log"; tail -n 40 "$OUT/pytest_attn.log" | cut -c1-300
timeout 600 python scripts/attn_bench.py > "$OUT/attn_bench.log" 2>&1; cat "$OUT/attn_bench.log"
