#!/usr/bin/env bash
# One gpurun call: GPU parity tests, smoke, bench on every workload, ncu launch list
# and one full ncu capture of the step's kernels.  Everything lands in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh [tag]'
set -u
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > "$OUT/gpu.csv" 2>&1

if [ "${MODE:-launches}" = "launches" ]; then
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > "$OUT/pytest.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest.log"
tail -n 30 "$OUT/pytest.log"

timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1
echo "smoke exit $?" >> "$OUT/smoke.log"
tail -n 3 "$OUT/smoke.log"

for w in sceneflow_576x960 corr_micro kitti_384x1280 train_320x640; do
  timeout 600 python bench.py --workload $w --steps 200 --warmup 20 > "$OUT/bench_$w.json" 2> "$OUT/bench_$w.err"
  echo "bench $w exit $?"; tail -c 600 "$OUT/bench_$w.err"
done
cat "$OUT/bench_sceneflow_576x960.json"

fi

# One ncu pattern per gpurun call: MODE=launches (default) or MODE=full
CMD="python bench.py --steps 4 --warmup 3 --no-graph --no-cpu-baseline"
if [ "${MODE:-launches}" = "launches" ]; then
  timeout 300 $CMD > "$OUT/ncu_plain.log" 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file "$OUT/launches.csv" $CMD > "$OUT/ncu_launches.log" 2>&1
  echo "ncu launches exit $?"
else
  timeout 300 $CMD > "$OUT/ncu_plain.log" 2>&1 &&
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'corr_|warp_' -s 12 -c 6 \
      -o "$OUT/prof" $CMD > "$OUT/ncu_full.log" 2>&1
  echo "ncu full exit $?"
fi
ls -la "$OUT"
