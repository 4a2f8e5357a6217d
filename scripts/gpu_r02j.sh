#!/usr/bin/env bash
# r02j: full GPU test-suite, headline bench (with api / e2e ceiling), train_ddp at N = 1, measurement document
set -u
TAG=${1:-r02j}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > "$OUT/pytest.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest.log"; tail -n 6 "$OUT/pytest.log"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; tail -n 1 "$OUT/smoke.log"
timeout 600 python bench.py --steps 200 --warmup 20 > "$OUT/bench_sceneflow_576x960.json" 2> "$OUT/bench_sceneflow_576x960.err"; echo "bench exit $?"; tail -3 "$OUT/bench_sceneflow_576x960.err"
timeout 600 python bench.py --workload train_ddp --steps 20 --warmup 5 > "$OUT/bench_train_ddp_n1.json" 2> "$OUT/bench_train_ddp_n1.err"; echo "train_ddp exit $?"; tail -5 "$OUT/bench_train_ddp_n1.err"
timeout 300 python scripts/band_occ_debug.py > "$OUT/band_occ_debug.log" 2>&1; cat "$OUT/band_occ_debug.log"
