#!/usr/bin/env python
"""warp_bwd at 384x1280 over the batch size: the 320-thread band instance (two blocks per SM) against the 512-thread one."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["DSSMD_TUNING"] = "1"
import torch
import bench
from dssmd_b200 import _lib
lib = _lib.lib(); dev = torch.device("cuda", 0)
for B in (1, 2, 3, 4, 6, 8, 16):
    w = dict(bench.WORKLOADS["kitti_384x1280"]); w["B"] = B
    nsets = max(2, min(8, int(600e6 / (B * 384 * 1280 * 4 * 11)) + 1))
    sets = [bench.make_set(torch, w, dev, 1024 + i) for i in range(nsets)]
    calls = [bench.kernel_calls(torch, lib, _lib, w, s) for s in sets]
    out = []
    for v in ("0", "1"):
        os.environ["DSSMD_WARP_BWD_BAND_W320"] = v
        def rnd():
            for c in calls: c["warp_bwd"]()
        rnd(); torch.cuda.synchronize()
        g = bench.capture(torch, rnd)
        for _ in range(3): g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): g.replay()
        e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) * 1e3 / (20 * nsets))
    print(f"B={B:2d} rows={B*384:5d}: 512-thread instance {out[0]:7.2f} us   320-thread instance {out[1]:7.2f} us")
