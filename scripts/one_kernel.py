#!/usr/bin/env python
"""Launch one C-ABI call of a bench workload a few times, eagerly (no CUDA graph): the command line ncu wraps.
    python scripts/one_kernel.py warp_bwd [workload] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

os.environ.setdefault("DSSMD_TUNING", "1")   # the sweep / debug variables are honoured only with it
import torch  # noqa: E402

import bench  # noqa: E402
from dssmd_b200 import _lib  # noqa: E402

name = sys.argv[1]
wname = sys.argv[2] if len(sys.argv) > 2 else "sceneflow_576x960"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
dev = torch.device("cuda", 0)
w = bench.WORKLOADS[wname]
sets = [bench.make_set(torch, w, dev, 1024 + i) for i in range(2)]
calls = [bench.kernel_calls(torch, _lib.lib(), _lib, w, s, split_bwd=True) for s in sets]
for i in range(reps):
    for n in name.split(","):
        calls[i % 2][n]()
torch.cuda.synchronize()
print("ok", name, wname, reps)
