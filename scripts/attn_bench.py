#!/usr/bin/env python
"""Times the fused epipolar attention (dssmd_epipolar_attn_fwd / _bwd) against upstream's expression
(models/BidNet/STTM_Simple.py:43-57) with torch ops on the same GPU, fp32 matmuls (development tool; CUDA events).

    python scripts/attn_bench.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import dssmd_b200  # noqa: E402


def upstream(q, k, v, scale):
    b, c, h, w = q.shape
    q2, k2, v2 = (t.permute(0, 2, 3, 1).reshape(b * h, w, c).contiguous() for t in (q, k, v))
    attn = torch.softmax(torch.matmul(q2, k2.transpose(-1, -2)) * scale, dim=-1)
    return torch.matmul(attn, v2).view(b, h, w, c).contiguous().permute(0, 3, 1, 2)


def timed(fn, rounds):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rounds):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / rounds


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    for (B, C, H, W) in ((8, 32, 320, 640), (4, 32, 384, 1280), (8, 128, 40, 80))[:int(os.environ.get("ATTN_BENCH_SHAPES", "3"))]:
        q, k, v, go = (torch.randn(B, C, H, W, device="cuda") for _ in range(4))
        scale = 32 ** -0.5
        flop = 2 * 2.0 * B * H * W * W * C
        for name, fn in (("fused", dssmd_b200.epipolar_attention), ("torch", upstream)):
            with torch.no_grad():
                f_ms = timed(lambda: fn(q, k, v, scale), 5)
            qq, kk, vv = (t.clone().requires_grad_(True) for t in (q, k, v))

            def step():
                out = fn(qq, kk, vv, scale)
                out.backward(go)
                qq.grad = kk.grad = vv.grad = None
            fb_ms = timed(step, 3)
            peak = torch.cuda.max_memory_allocated() / 2 ** 30
            torch.cuda.reset_peak_memory_stats()
            print(f"attn B{B} C{C} {H}x{W} {name}: fwd {f_ms:8.3f} ms ({flop / f_ms / 1e9:6.1f} TFLOP/s)   fwd+bwd {fb_ms:8.3f} ms   peak mem {peak:6.2f} GiB", flush=True)
        del q, k, v, go
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
