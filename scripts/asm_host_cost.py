#!/usr/bin/env python
"""Host cost of one level of the ASM reconstruction loss through the public API (forward only and forward + backward), us per
call: wall clock over many eager calls at the 1/8 level of cfg3 (B8, 40x80 transmission map, 320x640 images), where the
kernel itself takes ~5 us -- what remains is Python -> ctypes -> launch.  VERDICT r1 item 5: <= 30 us per level."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dssmd_b200
from dssmd_b200.losses import recovered_clean_l1

dev = torch.device("cuda", 0)
B = 8
haze, clean = torch.rand(B, 3, 320, 640, device=dev), torch.rand(B, 3, 320, 640, device=dev)
res = {}
for name, (h, w) in {"1/8": (40, 80), "1/1": (320, 640)}.items():
    t = (0.2 + 0.7 * torch.rand(B, 1, h, w, device=dev)).requires_grad_(True)
    a = (0.7 + 0.3 * torch.rand(B, 1, device=dev)).requires_grad_(True)
    for mode in ("fwd", "fwd+bwd"):
        def step():
            loss = recovered_clean_l1(t, a, haze, clean)
            if mode == "fwd+bwd":
                loss.backward()
                t.grad = None; a.grad = None
        for _ in range(50): step()
        torch.cuda.synchronize()
        n = 2000
        t0 = time.perf_counter()
        for _ in range(n): step()
        torch.cuda.synchronize()
        res[f"{name} {mode}"] = round((time.perf_counter() - t0) / n * 1e6, 1)
print(json.dumps({"asm_loss_us_per_call_eager": res}))
