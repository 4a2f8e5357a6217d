#!/usr/bin/env python
"""corr_bwd (both gradients) over the batch size at the KITTI and 320x640 feature shapes: the wide FFMA kernel (one launch) against the
mma.sync kernels issued as one launch (DSSMD_MMA32_BOTH=2 with the FFMA one-launch kernels off)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["DSSMD_TUNING"] = "1"
import torch
import bench
from dssmd_b200 import _lib
lib = _lib.lib(); dev = torch.device("cuda", 0)
def t(calls, nsets):
    def rnd():
        for c in calls: c["corr_bwd"]()
    rnd(); torch.cuda.synchronize()
    g = bench.capture(torch, rnd)
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (20 * nsets)
for wname, Bs in (("kitti_384x1280", (1, 2, 3, 4, 6, 8, 16)), ("train_320x640", (1, 2, 4, 8, 16)), ("sceneflow_576x960", (1, 2, 4, 8))):
    for B in Bs:
        w = dict(bench.WORKLOADS[wname]); w["B"] = B; w["H"], w["W"] = 32, 64   # small images: only the correlation is timed
        nsets = 6
        sets = [bench.make_set(torch, w, dev, 1024 + i) for i in range(nsets)]
        calls = [bench.kernel_calls(torch, lib, _lib, w, s) for s in sets]
        res = []
        for env in ({"DSSMD_MMA32_BOTH": "0"}, {"DSSMD_MMA32_BOTH": "2", "DSSMD_CORR_BWD_WIDE": "0", "DSSMD_CORR_BWD_FUSED": "0"}):
            for k in ("DSSMD_MMA32_BOTH", "DSSMD_CORR_BWD_WIDE", "DSSMD_CORR_BWD_FUSED"): os.environ.pop(k, None)
            os.environ.update(env)
            res.append(t(calls, nsets))
        print(f"{wname} feature shape B={B:2d}: FFMA one-launch (wide / fused) or two mma.sync launches {res[0]:7.2f} us   mma.sync one launch {res[1]:7.2f} us")
