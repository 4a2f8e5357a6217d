#!/usr/bin/env python
"""A few forward calls of the epipolar attention at BidNet's shape: the command line ncu wraps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dssmd_b200
q, k, v = (torch.randn(8, 32, 320, 640, device="cuda") for _ in range(3))
with torch.no_grad():
    for _ in range(3):
        dssmd_b200.epipolar_attention(q, k, v, 32 ** -0.5)
torch.cuda.synchronize()
print("ok")
