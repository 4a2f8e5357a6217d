#!/usr/bin/env python
"""corr_bwd at KITTI's feature shape (C128 48x160 D24) over the batch size, wide kernel forced vs default dispatch (development tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
os.environ["DSSMD_TUNING"] = "1"
import torch
import bench, kbench
from dssmd_b200 import _lib
lib = _lib.lib(); dev = torch.device("cuda", 0)
for B in (1, 2, 3, 4, 6, 8, 16):
    w = dict(bench.WORKLOADS["kitti_384x1280"]); w["B"] = B; w["H"], w["W"] = 64, 64   # small images: only the correlation is timed
    nsets = max(4, min(16, 400 // (B * 17)))
    sets = [bench.make_set(torch, w, dev, 1 + i) for i in range(nsets)]
    calls = [bench.kernel_calls(torch, lib, _lib, w, s) for s in sets]
    res = {}
    for wide in ("0", "2"):
        os.environ["DSSMD_CORR_BWD_WIDE"] = wide
        res[wide] = kbench.time_kernel(calls, "corr_bwd", nsets)
    print(f"B={B:2d} rows={B*48:4d}: other kernels {res['0']:7.2f} us   wide forced {res['2']:7.2f} us", flush=True)
