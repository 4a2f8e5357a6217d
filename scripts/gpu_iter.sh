#!/usr/bin/env bash
# Development iteration on the GPU box: parity tests + kernel sweeps + bench.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_iter.sh TAG "kbench args;kbench args" '
set -u
TAG=${1:-iter}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x > "$OUT/pytest.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest.log"
tail -n 5 "$OUT/pytest.log"
IFS=';' read -ra SWEEPS <<< "${2:-}"
for s in "${SWEEPS[@]}"; do
  [ -z "$s" ] && continue
  name=$(echo "$s" | tr ' ,' '__')
  timeout 600 python scripts/kbench.py $s > "$OUT/kbench_$name.log" 2>&1
  echo "kbench $s exit $?"; cat "$OUT/kbench_$name.log"
done
if [ "${TORCH_REF:-0}" = "1" ]; then
  timeout 600 python scripts/torch_ref_bench.py > "$OUT/torch_ref.log" 2>&1; tail -n 6 "$OUT/torch_ref.log"
fi
for w in sceneflow_576x960 corr_micro; do
  timeout 600 python bench.py --workload $w --steps 200 --warmup 20 > "$OUT/bench_$w.json" 2> "$OUT/bench_$w.err"
  echo "bench $w exit $?"; tail -c 400 "$OUT/bench_$w.err"
  python - <<PY
import json
d=json.load(open("$OUT/bench_$w.json"))
print("$w value",round(d['value'],1),"ms/step",round(d['ms_per_step'],4),"e2e",round(d['e2e']['value'],1))
for k,v in d['kernels'].items(): print('  ',k,v['us'],'us',v['GBps'],'GB/s',v['frac_of_peak'])
PY
done
