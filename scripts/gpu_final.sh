#!/usr/bin/env bash
# Round-end check: GPU tests, smoke, bench (all workloads), ncu capture of the mma16 kernels.
set -u
TAG=${1:-final}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > "$OUT/pytest.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest.log"; tail -n 3 "$OUT/pytest.log"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; tail -n 1 "$OUT/smoke.log"
for w in sceneflow_576x960 corr_micro kitti_384x1280 train_320x640; do
  timeout 600 python bench.py --workload $w --steps 200 --warmup 20 > "$OUT/bench_$w.json" 2> "$OUT/bench_$w.err"
  echo "bench $w exit $?"
done
timeout 300 python scripts/dtype_bench.py corr_micro,sceneflow_576x960,kitti_384x1280,train_320x640 > "$OUT/dtype_bench.log" 2>&1; cat "$OUT/dtype_bench.log"
CMD="python scripts/mma16_step.py 3"
timeout 300 $CMD > "$OUT/mma16_plain.log" 2>&1 && tail -1 "$OUT/mma16_plain.log" &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mma16' -s 3 -c 3 -o "$OUT/mma16_prof" $CMD > "$OUT/mma16_ncu.log" 2>&1
echo "ncu exit $?"; ls -la "$OUT" | head -30
