#!/usr/bin/env bash
set -u
TAG=${1:-r02l}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
for i in 1 2 3; do
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -p no:cacheprovider -k "warp_bwd or many_taps" > "$OUT/pytest_$i.log" 2>&1; tail -n 3 "$OUT/pytest_$i.log"; grep "^E  " "$OUT/pytest_$i.log" | head -5
done
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > "$OUT/pytest_full.log" 2>&1; tail -n 4 "$OUT/pytest_full.log"; grep "^E  " "$OUT/pytest_full.log" | head -8
