#!/usr/bin/env python
"""Correctness + timing of the tcgen05 correlation forward (DSSMD_CORR_FWD_VARIANT=2) against the
fp64 torch restatement and the FFMA path (development tool; run under `timeout`)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import dssmd_b200  # noqa: E402
import ref_torch  # noqa: E402

os.environ["DSSMD_CORR_DEBUG"] = "1"


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


shapes = [(1, 16, 2, 64, 5), (1, 32, 3, 80, 24), (2, 128, 4, 120, 24), (1, 256, 4, 160, 41), (1, 20, 3, 44, 7),
          (4, 128, 72, 120, 24), (8, 128, 40, 80, 24), (16, 128, 48, 160, 24), (8, 256, 80, 160, 41)]
dt = torch.bfloat16 if len(sys.argv) > 1 and sys.argv[1] == "bf16" else torch.float32
for (B, C, H, W, D) in shapes:
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + W)
    for relu in (True, False):
        L = torch.randn((B, C, H, W), generator=g, device="cuda")
        R = torch.randn((B, C, H, W), generator=g, device="cuda")
        if relu:
            L, R = L.relu(), R.relu()
        L, R = L.to(dt), R.to(dt)
        ref = ref_torch.corr_ref(L.double(), R.double(), D)
        os.environ["DSSMD_CORR_FWD_VARIANT"] = "2"
        out = dssmd_b200.build_corr(L, R, D)
        torch.cuda.synchronize()
        err = float((out.double() - ref).abs().max() / ref.abs().max())
        os.environ["DSSMD_CORR_FWD_VARIANT"] = "1"
        out1 = dssmd_b200.build_corr(L, R, D)
        err1 = float((out1.double() - ref).abs().max() / ref.abs().max())
        print(f"shape {B}x{C}x{H}x{W} D={D} relu={relu}: umma rel err {err:.3e}   ffma rel err {err1:.3e}", flush=True)
    os.environ.pop("DSSMD_CORR_DEBUG", None)
    os.environ["DSSMD_CORR_FWD_VARIANT"] = "2"
    t2 = timeit(lambda: dssmd_b200.build_corr(L, R, D))
    os.environ["DSSMD_CORR_FWD_VARIANT"] = "1"
    t1 = timeit(lambda: dssmd_b200.build_corr(L, R, D))
    nbytes = (2 * B * C * H * W + B * D * H * W) * L.element_size()
    print(f"   time (python shim, hot L2 for small shapes): umma {t2:8.1f} us ({nbytes/t2/1e3:7.0f} GB/s)   ffma {t1:8.1f} us", flush=True)
    os.environ["DSSMD_CORR_DEBUG"] = "1"
