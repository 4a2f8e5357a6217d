#!/usr/bin/env bash
set -u
TAG=${1:-r03b}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
for wv in 1 2; do DSSMD_TUNING=1 DSSMD_CORR_BWD_WIDE=$wv timeout 300 python scripts/measure_r02.py --only kitti_b1 --out "$OUT/kitti_b1_wide$wv.json" > /dev/null 2>&1; python - <<PY
import json
d=json.load(open("$OUT/kitti_b1_wide$wv.json"))["kitti_b1"]
print("wide=$wv", {k:{kk:vv["us"] for kk,vv in v.items() if isinstance(vv,dict)} for k,v in d.items() if isinstance(v,dict) and "corr_bwd" in v})
PY
done
