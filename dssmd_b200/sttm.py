"""Attention along the epipolar line: the core of upstream's STTM_Single
(models/BidNet/STTM_Simple.py:7-63, used by models/BidNet/BidNet.py:115, 239-240) as one fused kernel.

`epipolar_attention(q, k, v, scale)` takes the outputs of the three 1x1 convolutions as they are -- [B,C,H,W], NCHW --
and returns `softmax(q k^T * scale) v` per image row in NCHW: no permute / reshape / contiguous copies and no
[B*H, W, W] attention matrix (two 4.2 GB fp32 intermediates at BidNet's B8 x 32 x 320 x 640).  `STTM_Single` is
upstream's module with the same constructor, parameter names (state_dicts load unchanged) and forward signature, its
attention core swapped for the kernel (C <= 32: tensor cores with fp16 hi / lo split operands at the fp32 parity bar,
csrc/epipolar_attn_mma.cu; C = 64 / 128: fp32 FFMA).  fp32, C in {16, 32, 64, 128}, W % 4 == 0; anything else raises
NotImplementedError -- there is no fallback in this package.  `install()` does not replace upstream's class: it patches
the class's `forward` with `guarded_forward`, which takes the kernel for inputs inside those limits and hands everything
else to upstream's own, saved `forward`.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from .functional import _check4

__all__ = ["epipolar_attention", "STTM_Single", "kernel_supports", "guarded_forward"]

_CHANNELS = (16, 32, 64, 128)


def kernel_supports(channels, left_feat, right_feat):
    """True when `dssmd_epipolar_attn_*` takes the attention of these features (after 1x1 convolutions to `channels`)."""
    if not (isinstance(left_feat, torch.Tensor) and isinstance(right_feat, torch.Tensor)):
        return False
    if not (left_feat.is_cuda and right_feat.is_cuda) or left_feat.device != right_feat.device:
        return False
    if left_feat.dtype != torch.float32 or right_feat.dtype != torch.float32 or torch.is_autocast_enabled():
        return False
    if left_feat.dim() != 4 or left_feat.shape != right_feat.shape:
        return False
    B, _, H, W = left_feat.shape
    return channels in _CHANNELS and W % 4 == 0 and B * H <= 65535


def guarded_forward(upstream_forward):
    """The `forward` install() puts on upstream's STTM_Single: upstream's projections and output head, the attention core
    on the kernel; inputs outside the kernel's limits run upstream's saved forward unchanged."""
    def forward(self, left_feat, right_feat):
        if not kernel_supports(self.q.out_channels, left_feat, right_feat):
            return upstream_forward(self, left_feat, right_feat)
        out = epipolar_attention(self.q(left_feat), self.k(right_feat), self.v(right_feat), self.scale)
        return self.to_out(torch.cat((left_feat, out), dim=1))
    forward.__wrapped_upstream__ = upstream_forward
    return forward


class EpipolarAttentionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, scale):
        B, C, H, W = q.shape
        Q, K, V = q.contiguous(), k.contiguous(), v.contiguous()
        out = torch.empty_like(Q)
        need = any(ctx.needs_input_grad[:3])
        lse = torch.empty((B * H, W), dtype=torch.float32, device=Q.device) if need else None
        if Q.numel() > 0:
            with _lib.device_guard(Q.device):
                _lib.check(_lib.lib().dssmd_epipolar_attn_fwd(_lib.ptr(Q), _lib.ptr(K), _lib.ptr(V), _lib.ptr(out), _lib.ptr(lse),
                                                              B, C, H, W, float(scale), _lib.dtype_code(Q), _lib.stream_ptr(Q.device)))
        if need:
            ctx.save_for_backward(Q, K, V, out, lse)
        ctx.scale = float(scale)
        return out

    @staticmethod
    def backward(ctx, g):
        Q, K, V, out, lse = ctx.saved_tensors
        B, C, H, W = Q.shape
        need_q, need_k, need_v = ctx.needs_input_grad[:3]
        g = g.contiguous()
        gq = torch.empty_like(Q) if need_q else None
        gk = torch.empty_like(K) if (need_k or need_v) else None     # dK and dV come out of one kernel
        gv = torch.empty_like(V) if (need_k or need_v) else None
        if Q.numel() > 0 and (need_q or need_k or need_v):
            delta = torch.empty(B * H * W + 4 * B * H, dtype=torch.float32, device=Q.device)   # [B*H, W] sums + per-row scales
            with _lib.device_guard(Q.device):
                _lib.check(_lib.lib().dssmd_epipolar_attn_bwd(_lib.ptr(Q), _lib.ptr(K), _lib.ptr(V), _lib.ptr(out), _lib.ptr(g), _lib.ptr(lse),
                                                              _lib.ptr(delta), _lib.ptr(gq), _lib.ptr(gk), _lib.ptr(gv), B, C, H, W, ctx.scale,
                                                              _lib.dtype_code(Q), _lib.stream_ptr(Q.device)))
        return gq, (gk if need_k else None), (gv if need_v else None), None


def epipolar_attention(q, k, v, scale):
    """out[b,:,y,x] = sum_x' softmax_x'(scale * <q[b,:,y,x], k[b,:,y,x']>) v[b,:,y,x']  -- q, k, v, out: [B,C,H,W]."""
    _lib.require_cuda("epipolar_attention", q, k, v)
    for t in (q, k, v):
        _check4("epipolar_attention", t)
    if q.shape != k.shape or q.shape != v.shape:
        raise RuntimeError(f"epipolar_attention: q {tuple(q.shape)}, k {tuple(k.shape)} and v {tuple(v.shape)} differ")
    if any(t.dtype != torch.float32 for t in (q, k, v)):
        raise NotImplementedError("epipolar_attention: only float32 is implemented")
    B, C, H, W = q.shape
    if C not in _CHANNELS or W % 4 != 0 or B * H > 65535:
        raise NotImplementedError(f"epipolar_attention: needs C in (16, 32, 64, 128), W % 4 == 0 and B*H <= 65535, got {tuple(q.shape)}")
    return EpipolarAttentionFn.apply(q, k, v, scale)


class STTM_Single(nn.Module):
    """Upstream's STTM_Single (models/BidNet/STTM_Simple.py:7-63): same constructor, same parameter names, same forward."""

    def __init__(self, dim, heads=8, dim_head=64, out_dim=256):
        super().__init__()
        inner_dim = dim_head * heads
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.attend = nn.Softmax(dim=-1)          # kept for state_dict / repr parity; the kernel does the softmax
        self.q = nn.Conv2d(dim, inner_dim, kernel_size=1, stride=1, padding=0, bias=False)
        self.k = nn.Conv2d(dim, inner_dim, kernel_size=1, stride=1, padding=0, bias=False)
        self.v = nn.Conv2d(dim, inner_dim, kernel_size=1, stride=1, padding=0, bias=False)
        self.norm = nn.LayerNorm(dim)             # unused upstream as well; holds parameters, so it stays
        self.to_out = nn.Sequential(
            nn.Conv2d(inner_dim * 2, out_dim, kernel_size=1, stride=1, padding=0, bias=False),
            nn.BatchNorm2d(out_dim),
            nn.LeakyReLU(0.2),
            nn.Conv2d(out_dim, out_dim, kernel_size=1, stride=1, padding=0, bias=False),
        )

    def forward(self, left_feat, right_feat):
        q = self.q(left_feat)
        k = self.k(right_feat)
        v = self.v(right_feat)
        out = epipolar_attention(q, k, v, self.scale)
        new_left = torch.cat((left_feat, out), dim=1)
        return self.to_out(new_left)
