#!/usr/bin/env bash
# multi-GPU checks on N GPUs of one box: scripts/gpu_multi.sh TAG N
set -u
TAG=${1:-r02m}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
if [ "$N" = "2" ]; then
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -p no:cacheprovider -k "data_parallel or warp_bwd_band_matches" > "$OUT/pytest_dp.log" 2>&1; tail -n 3 "$OUT/pytest_dp.log"
fi
P=29611
for n in $(if [ "$N" = "8" ]; then echo "2 4 8"; else echo "$N"; fi); do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n --steps 100 --warmup 10 --no-aux > "$OUT/bench_n$n.json" 2> "$OUT/bench_n$n.err"; echo "bench n=$n exit $?"; P=$((P+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n --workload train_ddp --steps 20 --warmup 5 > "$OUT/train_ddp_n$n.json" 2> "$OUT/train_ddp_n$n.err"; echo "train_ddp n=$n exit $?"; P=$((P+1))
done
python - <<'PY'
import json,glob,os
for f in sorted(glob.glob(os.path.join("gpurun_out", os.environ.get("TAG_","")+"*","*.json"))): pass
PY
for f in "$OUT"/bench_n*.json; do python -c "
import json,sys
b=json.load(open('$f'))
print('$f', 'value', round(b['value']), 'ms', round(b['ms_per_step'],4), 'e2e', round(b['e2e']['value']), 'e2e_ms', round(b['e2e']['ms_per_step'],2), 'copy_ms', round(b['e2e']['copy_only_ms_per_step'],2), 'ceiling GB/s', b['e2e']['host_copy_ceiling_gbs'], b['e2e']['numa'])
"; done
for f in "$OUT"/train_ddp_n*.json; do python -c "
import json
b=json.load(open('$f'))
print('$f', 'pairs/s', round(b['value']), b['ddp'])
"; done
