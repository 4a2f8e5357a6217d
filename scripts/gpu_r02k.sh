#!/usr/bin/env bash
set -u
TAG=${1:-r02k}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 600 python bench.py --workload train_ddp --steps 20 --warmup 5 > "$OUT/bench_train_ddp_n1.json" 2> "$OUT/bench_train_ddp_n1.err"; echo "train_ddp exit $?"; tail -5 "$OUT/bench_train_ddp_n1.err"; cat "$OUT/bench_train_ddp_n1.json" | cut -c1-1500
timeout 300 python scripts/band_occ_debug.py > "$OUT/band_occ_debug.log" 2>&1; cat "$OUT/band_occ_debug.log"
