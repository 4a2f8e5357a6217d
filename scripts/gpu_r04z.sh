#!/usr/bin/env bash
# r04z: full GPU test-suite, smoke, headline bench, launch list + ncu --set full of the default bench command (all four calls)
set -u
TAG=${1:-r04z}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > "$OUT/pytest.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest.log"; tail -n 6 "$OUT/pytest.log"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; tail -n 1 "$OUT/smoke.log"
timeout 600 python bench.py --steps 200 --warmup 20 > "$OUT/bench_sceneflow_576x960.json" 2> "$OUT/bench_sceneflow_576x960.err"; echo "bench exit $?"; tail -3 "$OUT/bench_sceneflow_576x960.err"
for w in corr_micro kitti_384x1280 train_320x640; do
  timeout 600 python bench.py --workload $w --steps 100 --warmup 10 --no-api > "$OUT/bench_$w.json" 2> "$OUT/bench_$w.err"; echo "bench $w exit $?"
done
CMD="python bench.py --workload sceneflow_576x960 --steps 4 --warmup 3 --no-graph --no-cpu-baseline --no-api"
timeout 300 $CMD > "$OUT/plain.log" 2>&1 && {
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" $CMD > "$OUT/ncu_launches.log" 2>&1; echo "launch list exit $?"
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'corr_fwd|corr_bwd|warp_fwd|warp_bwd' -s 12 -c 4 -f -o "$OUT/prof" $CMD > "$OUT/ncu.log" 2>&1; echo "ncu exit $?"; tail -2 "$OUT/ncu.log"
}
ls -la "$OUT" | head -30
