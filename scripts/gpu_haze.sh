#!/usr/bin/env bash
set -u
TAG=${1:-haze}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > "$OUT/pytest.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest.log"; tail -n 15 "$OUT/pytest.log"
timeout 300 python scripts/haze_bench.py > "$OUT/haze_bench.log" 2>&1; cat "$OUT/haze_bench.log"
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > "$OUT/bench.json" 2> "$OUT/bench.err"; cat "$OUT/bench.json"
