#!/usr/bin/env bash
# L2 prefetch before griddepcontrol.wait: per-kernel A/B (graph replay, inputs rotating over > L2) and the whole step
set -u
TAG=${1:-r04b}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
WL=${2:-sceneflow_576x960}
for k in corr_fwd corr_bwd warp_fwd warp_bwd; do
  timeout 300 python scripts/kbench.py $k $WL DSSMD_PDL_PREFETCH=0,-,1,2,4,16 2>&1 | tee -a "$OUT/kbench_prefetch.log"
done
for pf in 0 -; do
  if [ "$pf" = "-" ]; then unset DSSMD_PDL_PREFETCH; else export DSSMD_PDL_PREFETCH=$pf; fi
  DSSMD_TUNING=1 timeout 600 python bench.py --workload $WL --no-cpu-baseline --no-aux --no-api 2> "$OUT/bench_pf$pf.err" | tee "$OUT/bench_pf$pf.json" | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pf=$pf', d['ms_per_step'], {k:v['us'] for k,v in d['kernels'].items()})"
done
