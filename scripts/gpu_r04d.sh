#!/usr/bin/env bash
# r04d: full GPU test-suite, smoke, headline bench and the three other workloads
set -u
TAG=${1:-r04d}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x > "$OUT/pytest.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest.log"; tail -n 6 "$OUT/pytest.log"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; tail -n 1 "$OUT/smoke.log"
timeout 600 python bench.py --steps 200 --warmup 20 > "$OUT/bench_sceneflow_576x960.json" 2> "$OUT/bench_sceneflow_576x960.err"; echo "bench exit $?"; tail -3 "$OUT/bench_sceneflow_576x960.err"
python - "$OUT/bench_sceneflow_576x960.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], {k:(v['us'], v['frac_of_peak']) for k,v in d['kernels'].items()}, d['e2e']['value'], d['api'])
PY
for w in corr_micro kitti_384x1280 train_320x640; do
  timeout 600 python bench.py --workload $w --steps 100 --warmup 10 --no-api --no-cpu-baseline --no-aux > "$OUT/bench_$w.json" 2> "$OUT/bench_$w.err"; echo "bench $w exit $?"
  python - "$OUT/bench_$w.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d['config']['workload'], d['ms_per_step'], {k:(v['us'], v['frac_of_peak']) for k,v in d['kernels'].items()})
PY
done
