#!/usr/bin/env python
"""One pass over the input-side and loss-side kernels of the training step at the headline
shape, for profiling:  haze synthesis of both views, correlation into the concat buffer (forward + backward), and the
ASM reconstruction loss over the 4-level transmission pyramid.  `ncu ... python scripts/aux_step.py` (development tool).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import dssmd_b200  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    B, H, W = 4, 576, 960
    dev = "cuda"
    torch.manual_seed(1024)
    clear_l, clear_r = torch.rand(B, 3, H, W, device=dev), torch.rand(B, 3, H, W, device=dev)
    disp_l = 0.5 + 191.5 * torch.rand(B, 1, H, W, device=dev)
    disp_r = 0.5 + 191.5 * torch.rand(B, 1, H, W, device=dev)
    sc64 = [torch.full((B,), v, dtype=torch.float64, device=dev) for v in (0.1, 1050.0, 0.1, 0.9)]
    sc32 = [t.float() for t in sc64]
    L = torch.randn(B, 128, H // 8, W // 8, device=dev).relu().requires_grad_(True)
    R = torch.randn(B, 128, H // 8, W // 8, device=dev).relu().requires_grad_(True)
    head = torch.randn(B, 32, H // 8, W // 8, device=dev)
    g_cat = torch.randn(B, 56, H // 8, W // 8, device=dev)
    pyr = [(0.1 + 0.9 * torch.rand(B, 1, H // s, W // s, device=dev)).requires_grad_(True) for s in (8, 4, 2, 1)]
    A = (0.8 + 0.2 * torch.rand(B, 1, device=dev)).requires_grad_(True)
    loss_fn = dssmd_b200.RecoveredCleanImagesLoss()
    for _ in range(steps):
        # input side (trainfiles/dssmd_traner_new.py:250-256): float64 scalars as the loaders hand them over, then float32
        haze_l, trans_l, _ = dssmd_b200.synthesize_haze(clear_l, disp_l, *sc64)
        haze_r, _, _ = dssmd_b200.synthesize_haze(clear_r, disp_r, *sc64)
        haze_l32, _, _ = dssmd_b200.synthesize_haze(clear_l, disp_l, *sc32)
        # correlation into the 56-channel concat buffer (models/DSSMD.py:129-131)
        cat = dssmd_b200.build_corr_cat(head, L, R, 24)
        cat.backward(g_cat)
        # loss side (trainfiles/dssmd_traner_new.py:298-302)
        total = sum(loss_fn(t, A, haze_l32, clear_l) * wt for t, wt in zip(pyr, (0.6, 0.8, 1.0, 1.0)))
        total.backward()
    torch.cuda.synchronize()
    print("aux step ok", float(total))


if __name__ == "__main__":
    main()
