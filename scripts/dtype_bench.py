#!/usr/bin/env python
"""Correlation forward / backward timing per element type (development tool).

    python scripts/dtype_bench.py [workload,...]     # default: corr_micro,sceneflow_576x960
CUDA-event time per C-ABI call inside a CUDA graph that rotates over input sets larger than L2;
GB/s on the algorithmic bytes of that element type."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from dssmd_b200 import _lib  # noqa: E402

CODES = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


def main():
    names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["corr_micro", "sceneflow_576x960"]
    lib = _lib.lib()
    dev = torch.device("cuda", 0)
    peak, _ = bench.measured_peak()
    for wname in names:
        w = bench.WORKLOADS[wname]
        B, C, H, W, D = (w[k] for k in ("B", "C", "Hf", "Wf", "D"))
        for dt in (torch.float32, torch.bfloat16, torch.float16):
            es = torch.empty((), dtype=dt).element_size()
            nsets = max(3, int(3 * 126e6 / ((4 * B * C * H * W + 2 * B * D * H * W) * es)) + 1)
            sets = []
            for i in range(nsets):
                g = torch.Generator(device=dev).manual_seed(1024 + i)
                L = torch.randn((B, C, H, W), generator=g, device=dev).relu_().to(dt)
                R = torch.randn((B, C, H, W), generator=g, device=dev).relu_().to(dt)
                gv = torch.randn((B, D, H, W), generator=g, device=dev).to(dt)
                sets.append((L, R, gv, torch.empty_like(gv), torch.empty_like(L), torch.empty_like(R)))
            st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            p = lambda t: ctypes.c_void_p(t.data_ptr())
            null = ctypes.c_void_p(0)
            calls = {
                "corr_fwd": lambda s: _lib.check(lib.dssmd_corr_fwd(p(s[0]), p(s[1]), p(s[3]), B, C, H, W, D, CODES[dt], st())),
                "corr_bwd_gL": lambda s: _lib.check(lib.dssmd_corr_bwd(p(s[2]), p(s[0]), p(s[1]), p(s[4]), null, B, C, H, W, D, CODES[dt], st())),
                "corr_bwd_gR": lambda s: _lib.check(lib.dssmd_corr_bwd(p(s[2]), p(s[0]), p(s[1]), null, p(s[5]), B, C, H, W, D, CODES[dt], st())),
            }
            ab = (2 * B * C * H * W + B * D * H * W) * es
            for name, fn in calls.items():
                def one_round():
                    for s in sets:
                        fn(s)
                one_round()
                torch.cuda.synchronize()
                gr = bench.capture(torch, one_round)
                for _ in range(3):
                    gr.replay()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                rounds = 20
                e0.record()
                for _ in range(rounds):
                    gr.replay()
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 1e3 / (rounds * nsets)
                print(f"{wname:20s} {str(dt):15s} {name:12s} {us:9.2f} us  {ab/us/1e3:8.1f} GB/s  frac {ab/us/1e3/peak:.3f}", flush=True)
            del sets
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
