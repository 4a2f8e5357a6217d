#!/usr/bin/env python
"""Error of the fp32 tensor-core correlation against float64, with the cross terms on tf32 (3xTF32) and on bf16 (DSSMD_MMA32_BF16X=1):
max |err| / max |ref| for relu'd and for zero-mean features at the cfg2 and KITTI shapes.  Run once per setting (the variable is read per call)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["DSSMD_TUNING"] = "1"
import torch  # noqa: E402

import dssmd_b200  # noqa: E402


def ref64(L, R, D):
    L, R = L.double(), R.double()
    B, C, H, W = L.shape
    out = torch.zeros(B, D, H, W, dtype=torch.float64, device=L.device)
    for d in range(D):
        out[:, d, :, d:] = (L[:, :, :, d:] * R[:, :, :, :W - d]).mean(1)
    return out


def ref64_bwd(g, L, R):
    L = L.double().requires_grad_(True); R = R.double().requires_grad_(True)
    ref64(L, R, g.shape[1]).backward(g.double())
    return L.grad, R.grad


for name, (B, C, H, W, D) in {"cfg2": (2, 256, 80, 160, 41), "kitti": (2, 128, 48, 160, 24)}.items():
    for kind in ("relu", "zero-mean", "wide-range"):
        torch.manual_seed(7)
        L, R = torch.randn(B, C, H, W, device="cuda"), torch.randn(B, C, H, W, device="cuda")
        if kind == "relu":
            L, R = L.relu(), R.relu()
        if kind == "wide-range":   # channels of very different scale
            sc = torch.logspace(-3, 3, C, device="cuda").view(1, C, 1, 1)
            L, R = L * sc, R / sc
        g = torch.randn(B, D, H, W, device="cuda")
        r = ref64(L, R, D)
        gl, gr = ref64_bwd(g, L, R)
        for bx in ("0", "1"):
            os.environ["DSSMD_MMA32_BF16X"] = bx
            Lq, Rq = L.clone().requires_grad_(True), R.clone().requires_grad_(True)
            o = dssmd_b200.build_corr(Lq, Rq, D)
            o.backward(g)
            e = lambda a, b: float((a.double() - b).abs().max() / b.abs().max())  # noqa: E731
            print(f"{name:6s} {kind:10s} BF16X={bx}: fwd {e(o, r):.2e}  gL {e(Lq.grad, gl):.2e}  gR {e(Rq.grad, gr):.2e}")
