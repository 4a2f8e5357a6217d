#!/usr/bin/env python
"""cProfile of the eager public-API step (build_corr fwd + bwd, disp_warp fwd + bwd on device tensors) at the headline shape:
where the Python-side time per call goes (development tool)."""
import cProfile, pstats, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dssmd_b200
dev = torch.device("cuda", 0)
L = torch.randn(4, 128, 72, 120, device=dev).relu_().requires_grad_(True)
R = torch.randn(4, 128, 72, 120, device=dev).relu_().requires_grad_(True)
g = torch.randn(4, 24, 72, 120, device=dev)
img = torch.rand(4, 3, 576, 960, device=dev, requires_grad=True)
disp = (torch.rand(4, 1, 576, 960, device=dev) * 20).requires_grad_(True)
gi = torch.randn(4, 3, 576, 960, device=dev)
dssmd_b200.set_check_negative_disparity("deferred")

def step():
    dssmd_b200.build_corr(L, R, 24).backward(g)
    dssmd_b200.disp_warp(img, disp)[0].backward(gi)
    L.grad = R.grad = img.grad = disp.grad = None

for _ in range(20): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): step()
torch.cuda.synchronize()
print("eager step: %.1f us" % ((time.perf_counter() - t0) / 200 * 1e6))
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
pr.disable(); torch.cuda.synchronize()
ps = pstats.Stats(pr); ps.sort_stats("tottime").print_stats(22)
