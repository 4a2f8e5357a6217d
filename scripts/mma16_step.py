#!/usr/bin/env python
"""bf16 correlation forward + backward at BASELINE config 2 (B8, C256, 80x160, D41) through the public API, for profiling the
mma.sync kernels:  ncu ... python scripts/mma16_step.py   (development tool)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import dssmd_b200  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    torch.manual_seed(1024)
    L = torch.randn(8, 256, 80, 160, device="cuda").relu().bfloat16().requires_grad_(True)
    R = torch.randn(8, 256, 80, 160, device="cuda").relu().bfloat16().requires_grad_(True)
    g = torch.randn(8, 41, 80, 160, device="cuda").bfloat16()
    for _ in range(steps):
        vol = dssmd_b200.build_corr(L, R, 41)
        vol.backward(g)
        L.grad = R.grad = None
    torch.cuda.synchronize()
    print("mma16 step ok", float(vol.float().abs().mean()))


if __name__ == "__main__":
    main()
