#!/usr/bin/env bash
# GPU tests, the haze / ASM-loss bench, and ncu captures (launch list + one full capture) of scripts/aux_step.py
set -u
TAG=${1:-aux}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > "$OUT/pytest.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest.log"; tail -n 4 "$OUT/pytest.log"
timeout 300 python scripts/haze_bench.py > "$OUT/haze_bench.log" 2>&1; cat "$OUT/haze_bench.log"
CMD="python scripts/aux_step.py 3"
timeout 300 $CMD > "$OUT/aux_plain.log" 2>&1 && tail -1 "$OUT/aux_plain.log" &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/aux_launches.csv" $CMD > "$OUT/aux_ncu_launches.log" 2>&1
echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'haze_kernel|asm_recovered|corr_' -s 9 -c 9 -o "$OUT/aux_prof" $CMD > "$OUT/aux_ncu_full.log" 2>&1
echo "ncu full exit $?"; ls -la "$OUT"
