#!/usr/bin/env bash
# ncu --set full capture of selected kernels:  gpu_prof.sh TAG WORKLOAD KERNEL_REGEX [ENV=VAL ...]
set -u
TAG=$1; W=$2; K=$3; shift 3
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
for kv in "$@"; do export "$kv"; done
CMD="python bench.py --workload $W --steps 4 --warmup 3 --no-graph --no-cpu-baseline"
timeout 300 $CMD > "$OUT/plain.log" 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 10 -c 2 -o "$OUT/prof" $CMD > "$OUT/ncu.log" 2>&1
echo "ncu exit $?"; tail -3 "$OUT/ncu.log"; ls -la "$OUT"
