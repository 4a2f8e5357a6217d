#!/usr/bin/env bash
# one ncu --set full capture of one kernel: scripts/gpu_ncu.sh TAG KERNEL_REGEX CALL [WORKLOAD]   (env vars pass through)
set -u
TAG=$1; PAT=$2; CALL=$3; WL=${4:-sceneflow_576x960}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
CMD="python scripts/one_kernel.py $CALL $WL 4"
timeout 300 $CMD > "$OUT/plain.log" 2>&1 && tail -1 "$OUT/plain.log" &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$PAT" -s 2 -c 1 -f -o "$OUT/prof" $CMD > "$OUT/ncu.log" 2>&1
echo "ncu exit $?"; tail -3 "$OUT/ncu.log"; ls -la "$OUT"
