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


def graph_timed(fn, sets, rounds=20):
    """The same over a CUDA graph of one round (kernel time without the Python -> ctypes launch path)."""
    def one_round():
        for st in sets:
            fn(*st)
    one_round()
    torch.cuda.synchronize()
    g = bench.capture(torch, one_round)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rounds):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (rounds * len(sets))


def asm_upstream(t, A, haze, clean):
    """losses/DSSMDLoss.py:87-111 with torch ops, forward + backward."""
    t = t.detach().requires_grad_(True)
    A = A.detach().requires_grad_(True)
    s = haze.shape[-1] // t.shape[-1]
    if s != 1:
        haze = torch.nn.functional.interpolate(haze, scale_factor=1. / s, mode='bilinear', align_corners=False)
        clean = torch.nn.functional.interpolate(clean, scale_factor=1. / s, mode='bilinear', align_corners=False)
    a = A.unsqueeze(-1).unsqueeze(-1)
    loss = torch.nn.functional.l1_loss(torch.clamp((haze - a * (1 - t)) / t, min=0, max=1.0), clean)
    loss.backward()
    return loss


def asm_fused(t, A, haze, clean):
    t = t.detach().requires_grad_(True)
    A = A.detach().requires_grad_(True)
    loss = dssmd_b200.recovered_clean_l1(t, A, haze, clean)
    loss.backward()
    return loss


def asm_kernel_only(t, A, haze, clean):
    from dssmd_b200 import _lib
    lib = _lib.lib()
    B, C, H0, W0 = haze.shape
    h, w = t.shape[-2:]
    nblk = int(lib.dssmd_asm_recovered_l1_blocks(B, h, w))
    key = (B, h, w)
    if key not in _scratch:
        _scratch[key] = (torch.empty((B, nblk, 2), device="cuda"), torch.empty_like(t))
    partial, gt = _scratch[key]
    _lib.check(lib.dssmd_asm_recovered_l1_fwd(_lib.ptr(t), _lib.ptr(A.reshape(B)), _lib.ptr(haze), _lib.ptr(clean), _lib.ptr(gt),
                                              _lib.ptr(partial), B, C, H0, W0, h, w, 0, _lib.stream_ptr(t.device)))


_scratch = {}


def asm_main(peak):
    for (B, H, W) in ((4, 576, 960), (8, 320, 640), (16, 384, 1280)):
        for scale in (1, 2, 4, 8):
            h, w = H // scale, W // scale
            nsets = max(4, int(300e6 // (B * H * W * 4 * 7)) + 1)
            sets = []
            for i in range(nsets):
                g = torch.Generator(device="cuda").manual_seed(7 + i)
                clean = torch.rand(B, 3, H, W, generator=g, device="cuda")
                haze = torch.rand(B, 3, H, W, generator=g, device="cuda")
                t = 0.1 + 0.9 * torch.rand(B, 1, h, w, generator=g, device="cuda")
                A = 0.8 + 0.2 * torch.rand(B, 1, generator=g, device="cuda")
                sets.append((t, A, haze, clean))
            ab = B * h * w * 4 * (2 + 6 * (1 if scale == 1 else min(4, scale * scale)))
            k_us = graph_timed(asm_kernel_only, sets)
            f_us = timed(asm_fused, sets, rounds=5)
            r_us = timed(asm_upstream, sets, rounds=5)
            print(f"asm loss B{B} {H}x{W} /{scale}: kernel {k_us:7.1f} us {ab / k_us / 1e3:6.0f} GB/s ({ab / k_us / 1e3 / peak:.2f} of peak, "
                  f"{ab / 1e6:.1f} MB)  fwd+bwd via API {f_us:7.1f} us   torch fwd+bwd {r_us:8.1f} us  x{r_us / f_us:.1f}", flush=True)
            del sets


def main():
    peak, _ = bench.measured_peak()
    asm_main(peak)
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
