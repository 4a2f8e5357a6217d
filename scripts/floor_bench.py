#!/usr/bin/env python
"""What a kernel that only MOVES a step's bytes achieves in the same harness (CUDA-graph replay, sets rotating through more
than L2, CUDA events): torch's vectorised elementwise kernel over T/2 bytes (`add`: T/2 read + T/2 written) for the algorithmic byte counts T
of the hot-path calls.  MEASURED_PEAKS.json's 6.5 TB/s comes from a 4 GiB copy; a 40-100 MB kernel lives in the launch ramp
and the tail of that curve.  (development tool)

    python scripts/floor_bench.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    peak, _ = bench.measured_peak()
    cases = [("corr_fwd / corr_bwd_gX @cfg5", 38.7e6), ("warp_fwd @cfg5", 88.5e6), ("warp_bwd @cfg5", 97.3e6),
             ("corr_fwd @cfg2", 226.5e6), ("1 GB", 1e9), ("4 GB", 4e9)]
    for name, T in cases:
        n = int(T / 2 / 4)
        nsets = max(2, min(8, int(600e6 // T) + 1))
        srcs = [torch.rand(n, device="cuda") for _ in range(nsets)]
        dsts = [torch.empty(n, device="cuda") for _ in range(nsets)]

        def one_round():
            for s, d in zip(srcs, dsts):
                torch.add(s, 1.0, out=d)    # an SM kernel (a graph turns copy_ into a copy-engine memcpy node: 3.0 TB/s)
        one_round()
        torch.cuda.synchronize()
        g = bench.capture(torch, one_round)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        rounds = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(rounds):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (rounds * nsets)
        print(f"copy of {T / 1e6:7.1f} MB total ({name:30s}): {us:8.2f} us  {T / us / 1e3:7.0f} GB/s  {T / us / 1e3 / peak:.2f} of the measured peak", flush=True)
        del srcs, dsts, g
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
