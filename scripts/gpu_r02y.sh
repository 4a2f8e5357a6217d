#!/usr/bin/env bash
set -u
TAG=${1:-r02y}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -p no:cacheprovider -k "attention or sttm" 2>&1 | tail -25
cat > /tmp/attn_one.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
import dssmd_b200
q, k, v = (torch.randn(8, 32, 320, 640, device="cuda") for _ in range(3))
with torch.no_grad():
    for _ in range(3): dssmd_b200.epipolar_attention(q, k, v, 32 ** -0.5)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_mma -s 2 -c 1 -o "$OUT/attn_fwd_mma" -f python /tmp/attn_one.py > "$OUT/ncu_attn.log" 2>&1; tail -1 "$OUT/ncu_attn.log"
