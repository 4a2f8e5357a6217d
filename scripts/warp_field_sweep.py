#!/usr/bin/env python
"""warp_fwd / warp_bwd at the headline shape (B4 3x576x960) over disparity fields of different roughness (development tool and
the data-sensitivity measurement of DESIGN 5.4): the bench's own field (bilinear upsample of a 19 x 31 grid of uniform values in
[0.5, 192): slopes up to 6 px/px, the ordering constraint violated almost everywhere), the same grid at lower amplitudes, and a
scene-like field (ground-plane ramp + planar objects with disparity jumps at their edges, slopes <= 0.2 px/px inside a surface)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
os.environ.setdefault("DSSMD_TUNING", "1")
import torch
import bench, kbench
from dssmd_b200 import _lib
lib = _lib.lib(); dev = torch.device("cuda", 0)
w = bench.WORKLOADS["sceneflow_576x960"]
B, H, W = w["B"], w["H"], w["W"]
peak, _ = bench.measured_peak()


def scene_like(seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    y = torch.linspace(0, 1, H, device=dev).view(1, 1, H, 1)
    x = torch.linspace(0, 1, W, device=dev).view(1, 1, 1, W)
    d = (5.0 + 60.0 * y + 4.0 * x).expand(B, 1, H, W).clone()                 # ground plane: closer towards the bottom
    for b in range(B):
        for _ in range(12):                                                  # planar objects in front of it
            r = torch.rand(7, generator=g, device=dev).tolist()
            y0, x0 = int(r[0] * H * 0.8), int(r[1] * W * 0.8)
            hh, ww = int(40 + r[2] * 200), int(60 + r[3] * 300)
            yy = torch.arange(hh, device=dev).view(hh, 1).float(); xx = torch.arange(ww, device=dev).view(1, ww).float()
            obj = 30.0 + 150.0 * r[4] + (r[5] - 0.5) * 0.4 * xx + (r[6] - 0.5) * 0.2 * yy
            reg = d[b, 0, y0:y0 + hh, x0:x0 + ww]
            obj = obj[:reg.shape[0], :reg.shape[1]]
            d[b, 0, y0:y0 + hh, x0:x0 + ww] = torch.maximum(reg, obj)
    return d.clamp_(0.5, 191.5).contiguous()


def grid_field(seed, amp):
    g = torch.Generator(device=dev).manual_seed(seed)
    base = torch.rand((B, 1, H // 32 + 1, W // 32 + 1), generator=g, device=dev)
    disp = torch.nn.functional.interpolate(base, size=(H, W), mode="bilinear", align_corners=True)
    return (0.5 + disp * (amp - 0.5)).contiguous()


fields = {"bench (grid, amplitude 192: slopes <= 6 px/px)": lambda s: grid_field(s, 192.0),
          "grid, amplitude 32 (slopes <= 1 px/px)": lambda s: grid_field(s, 32.0),
          "grid, amplitude 8": lambda s: grid_field(s, 8.0),
          "constant 20 px": lambda s: torch.full((B, 1, H, W), 20.0, device=dev),
          "scene-like (ground plane + planar objects)": scene_like}
res = {}
nsets = 4
for name, mk in fields.items():
    sets = [bench.make_set(torch, w, dev, 1024 + i) for i in range(nsets)]
    for i, s in enumerate(sets):
        s["disp"] = mk(7 + i)
    calls = [bench.kernel_calls(torch, lib, _lib, w, s) for s in sets]
    row = {}
    for k in ("warp_fwd", "warp_bwd"):
        us = kbench.time_kernel(calls, k, nsets)
        ab = bench.algorithmic_bytes(w)[k]
        row[k] = {"us": round(us, 2), "frac_of_peak": round(ab / us / 1e3 / peak, 3)}
    res[name] = row
    print(f"{name:48s} warp_fwd {row['warp_fwd']['us']:6.2f} us ({row['warp_fwd']['frac_of_peak']:.3f})   warp_bwd {row['warp_bwd']['us']:6.2f} us ({row['warp_bwd']['frac_of_peak']:.3f})", flush=True)
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
