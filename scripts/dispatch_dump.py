#!/usr/bin/env python
"""Which kernel (and launch configuration) every bench workload's four calls take by default: one launch each with DSSMD_CORR_DEBUG=1."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["DSSMD_TUNING"] = "1"
os.environ["DSSMD_CORR_DEBUG"] = "1"
import torch  # noqa: E402

import bench  # noqa: E402
from dssmd_b200 import _lib  # noqa: E402

lib = _lib.lib()
dev = torch.device("cuda", 0)
for wname in ("sceneflow_576x960", "corr_micro", "kitti_384x1280", "train_320x640"):
    w = bench.WORKLOADS[wname]
    s = bench.make_set(torch, w, dev, 1024)
    calls = bench.kernel_calls(torch, lib, _lib, w, s, split_bwd=False)
    for name, fn in calls.items():
        print(f"== {wname} {name}", file=sys.stderr, flush=True)
        fn()
        torch.cuda.synchronize()
