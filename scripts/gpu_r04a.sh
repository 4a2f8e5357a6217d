#!/usr/bin/env bash
# attention WIP check: parity tests of the attention paths + forward / backward timing
set -u
TAG=${1:-r04a}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -k "attention or sttm or fresh_threads" > "$OUT/pytest_attn.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest_attn.log"; tail -n 15 "$OUT/pytest_attn.log" | cut -c1-300
timeout 600 python scripts/attn_bench.py > "$OUT/attn_bench.log" 2>&1; cat "$OUT/attn_bench.log"
