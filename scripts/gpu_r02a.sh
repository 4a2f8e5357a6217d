#!/usr/bin/env bash
# r02a: GPU tests, the drop-in proof on the real upstream tree (staged by scripts/stage_upstream.sh), the measurements
# BASELINE.json names beside the headline, and the baseline bench of the tree as round 2 found it.
set -u
TAG=${1:-r02a}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x > "$OUT/pytest.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest.log"; tail -n 3 "$OUT/pytest.log"
timeout 900 python scripts/upstream_dropin.py --gpu --iters 5 --out "$OUT/upstream_dropin.json" > "$OUT/upstream_dropin.log" 2>&1; echo "dropin exit $?"
timeout 900 python scripts/measure_r02.py --out "$OUT/measure.json" > "$OUT/measure.log" 2>&1; echo "measure exit $?"
timeout 600 python bench.py --steps 200 --warmup 20 > "$OUT/bench_sceneflow_576x960.json" 2> "$OUT/bench_sceneflow_576x960.err"; echo "bench exit $?"
ls -la "$OUT"
