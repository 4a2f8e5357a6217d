#!/usr/bin/env python
"""Times the fused haze-synthesis launch (dssmd_haze_fwd) against upstream's expression chain run with torch ops on
the same GPU (development tool).  CUDA events, inputs rotating over sets larger than L2.

    python scripts/haze_bench.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
import dssmd_b200  # noqa: E402


def upstream_chain(clear, disp, baseline, focal, beta, A):
    """DeBug/inference.py:57-64, 126-134, 95-112 as torch ops (what upstream runs per batch and view)."""
    b = clear.shape[0]
    depth = focal.view(b, 1, 1, 1) * baseline.view(b, 1, 1, 1) / disp
    trans = torch.exp(-depth * beta.view(b, 1, 1, 1))
    t3 = torch.exp(-depth.repeat(1, 3, 1, 1) * beta.view(b, 1, 1, 1))
    haze = (clear * t3) + A.view(b, 1, 1, 1) * (1 - t3)
    return haze, trans, depth


def timed(fn, sets, rounds=20):
    for s in sets:
        fn(*s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rounds):
        for s in sets:
            fn(*s)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (rounds * len(sets))


def main():
    peak, _ = bench.measured_peak()
    for (B, H, W) in ((4, 576, 960), (8, 320, 640), (16, 384, 1280)):
        for sdt in (torch.float32, torch.float64):
            nsets = max(4, int(300e6 // (B * H * W * 4 * 9)) + 1)
            sets = []
            for i in range(nsets):
                g = torch.Generator(device="cuda").manual_seed(1024 + i)
                clear = torch.rand(B, 3, H, W, generator=g, device="cuda")
                disp = 0.5 + torch.rand(B, 1, H, W, generator=g, device="cuda") * 191.5
                sc = [torch.full((B,), v, dtype=sdt, device="cuda") for v in (0.1, 1050.0, 0.1, 0.9)]
                sets.append((clear, disp, *sc))
            es = 4 if sdt == torch.float32 else 8
            ab = B * H * W * (4 * 4 + 5 * es)
            us = timed(dssmd_b200.synthesize_haze, sets)
            ref = timed(upstream_chain, sets)
            print(f"haze B{B} {H}x{W} scalars {str(sdt)[6:]}: fused {us:8.1f} us  {ab / us / 1e3:7.0f} GB/s "
                  f"({ab / us / 1e3 / peak:.2f} of peak, {ab / 1e6:.1f} MB)   torch chain {ref:8.1f} us  x{ref / us:.1f}", flush=True)
            del sets


if __name__ == "__main__":
    main()
