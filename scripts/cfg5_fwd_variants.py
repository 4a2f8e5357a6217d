#!/usr/bin/env python
"""corr_fwd at the 576x960 model's feature shape [4,128,72,120] over D = 5 / 12 / 24 (max_disp 40 / 96 / 192): the register-blocked FFMA
kernel against the mma.sync kernel (DSSMD_CORR_FWD_VARIANT=3)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["DSSMD_TUNING"] = "1"
import torch
import bench
from dssmd_b200 import _lib
lib = _lib.lib(); dev = torch.device("cuda", 0)
for wname, Ds in (("sceneflow_576x960", (5, 12, 24)), ("train_320x640", (24,)), ("kitti_384x1280", (24,))):
  for D in Ds:
    w = dict(bench.WORKLOADS[wname]); w["D"] = D
    if wname == "kitti_384x1280": w["B"] = 1
    nsets = 6
    sets = [bench.make_set(torch, w, dev, 1024 + i) for i in range(nsets)]
    calls = [bench.kernel_calls(torch, lib, _lib, w, s) for s in sets]
    res = []
    for v in ("", "1", "3"):
        if v: os.environ["DSSMD_CORR_FWD_VARIANT"] = v
        else: os.environ.pop("DSSMD_CORR_FWD_VARIANT", None)
        def rnd():
            for c in calls: c["corr_fwd"]()
        try:
            rnd(); torch.cuda.synchronize()
            g = bench.capture(torch, rnd)
            for _ in range(3): g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): g.replay()
            e1.record(); torch.cuda.synchronize()
            res.append(round(e0.elapsed_time(e1) * 1e3 / (20 * nsets), 2))
        except Exception as e:
            res.append(str(e)[:40])
    print(f"{wname} B={w['B']} D={D:3d}: default {res[0]}  reg(1) {res[1]}  mma32(3) {res[2]}")
