#!/usr/bin/env python
"""First-light check of the tcgen05 attention forward against float64 (development tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DSSMD_TUNING"] = "1"; os.environ["DSSMD_CORR_DEBUG"] = "1"
import torch
import dssmd_b200
torch.manual_seed(0)
for (B, C, H, W) in ((1, 32, 1, 128), (1, 32, 2, 256), (2, 32, 3, 640), (1, 32, 2, 328), (2, 32, 5, 1280)):
    q, k, v = (torch.randn(B, C, H, W, device="cuda") for _ in range(3))
    scale = C ** -0.5
    out = dssmd_b200.epipolar_attention(q, k, v, scale)
    torch.cuda.synchronize()
    q2, k2, v2 = (t.double().permute(0, 2, 3, 1).reshape(B * H, W, C) for t in (q, k, v))
    ref = torch.matmul(torch.softmax(torch.matmul(q2, k2.transpose(-1, -2)) * scale, dim=-1), v2).view(B, H, W, C).permute(0, 3, 1, 2)
    err = float((out.double() - ref).abs().max() / ref.abs().max())
    print(f"B{B} C{C} {H}x{W}: rel err {err:.3e}", flush=True)
