"""Photometric / reconstruction loss on the disparity warp.

Upstream defines the primitive -- LRwarp_error, utils/disparity_warper.py:109-115 == models/UtilsNet/stereo_op.py:67-73 --
but no loss calls it (the intended call site is the commented-out Resample2d in models/UtilsNet/submoudles.py:169).
photometric_loss is that loss, stated with upstream's own pieces so that it can be checked against them:

    loss = sum_s  w_s * LRwarp_error(imgL, disp_s, imgR).abs().mean()

over a disparity pyramid (the models return [1/8, 1/4, 1/2, 1] -- models/DSSMD.py:236-262 -- and upstream weights its
disparity pyramid 0.6 / 0.8 / 1.0 / 1.0, losses/DSSMDLoss.py).  Each term runs as ONE kernel that produces the row sums of
|err| and d/d(disp) together (dssmd_photometric_l1_fwd): the residual is never materialised and the backward is a scale.
When an image requires a gradient, or the shape is outside the fused kernel's limits, the term is composed from
LRwarp_error (same value, upstream's autograd graph).
"""
from __future__ import annotations

import torch

from . import _lib
from .functional import LRWarpErrorFn, _check4, _new_flag, _raise_if_flag, resize_bilinear

__all__ = ["photometric_l1", "photometric_loss", "PYRAMID_WEIGHTS"]

PYRAMID_WEIGHTS = (0.6, 0.8, 1.0, 1.0)  # upstream's weights for the [1/8, 1/4, 1/2, 1] disparity pyramid


class _PhotometricL1Fn(torch.autograd.Function):
    """mean |imgR' - warp(imgL', disp)| with the gradient w.r.t. disp computed in the forward pass."""

    @staticmethod
    def forward(ctx, imgL, disp, imgR):
        B, C, H, W = imgL.shape
        Lc, Rc, Dc = imgL.contiguous(), imgR.contiguous(), disp.contiguous()
        rows = torch.empty(B * H, dtype=torch.float32, device=Lc.device)
        need_d = ctx.needs_input_grad[1]
        g_unit = torch.empty_like(Dc) if need_d else None
        flag = _new_flag()
        with torch.cuda.device(Lc.device):
            _lib.check(_lib.lib().dssmd_photometric_l1_fwd(_lib.ptr(Lc), _lib.ptr(Rc), _lib.ptr(Dc), _lib.ptr(g_unit), _lib.ptr(rows),
                                                           _lib.ptr(flag), B, C, H, W, _lib.dtype_code(Lc), _lib.stream_ptr(Lc.device)))
        _raise_if_flag(flag, Lc.device)
        n = float(B * C * H * W)
        ctx.n = n
        if need_d:
            ctx.save_for_backward(g_unit)
        return rows.sum() / n

    @staticmethod
    def backward(ctx, g):
        g_disp = None
        if ctx.needs_input_grad[1]:
            (g_unit,) = ctx.saved_tensors
            g_disp = g_unit * (g / ctx.n)
        return None, g_disp, None


def _fused_ok(imgL, disp, imgR):
    B, C = imgL.shape[:2]
    H, W = disp.shape[-2:]
    return (imgL.dtype == torch.float32 and not imgL.requires_grad and not imgR.requires_grad and 1 <= C <= 4 and W % 4 == 0
            and W >= 4 and H >= 2 and B <= 65535 and C * H * W < 2 ** 31)


def photometric_l1(imgL, disp, imgR):
    """LRwarp_error(imgL, disp, imgR).abs().mean() (argument order of upstream's LRwarp_error)."""
    _lib.require_cuda("photometric_l1", imgL, disp, imgR)
    for t in (imgL, disp, imgR):
        _check4("photometric_l1", t)
    if imgL.shape != imgR.shape:
        raise RuntimeError(f"photometric_l1: imgL {tuple(imgL.shape)} and imgR {tuple(imgR.shape)} differ")
    if imgL.dtype != disp.dtype or imgR.dtype != disp.dtype:
        raise RuntimeError("photometric_l1: imgL, disp and imgR must share a dtype")
    B, C, H0, W0 = imgL.shape
    H, W = disp.shape[-2:]
    if tuple(disp.shape) != (B, 1, H, W):
        raise RuntimeError(f"photometric_l1: disp must be [B,1,H,W], got {tuple(disp.shape)}")
    if W0 <= W and (H0, W0) != (H, W):
        raise RuntimeError(f"photometric_l1: images {H0}x{W0} are not wider than disp {H}x{W} and do not match it")
    if disp.numel() == 0 or not _fused_ok(imgL, disp, imgR):
        return LRWarpErrorFn.apply(imgL, disp, imgR).abs().mean()
    if W0 > W:  # upstream resizes both images to disp's size first (disparity_warper.py:110-113)
        imgL, imgR = resize_bilinear(imgL, (H, W)), resize_bilinear(imgR, (H, W))
    return _PhotometricL1Fn.apply(imgL, disp, imgR)


def photometric_loss(imgL, imgR, disp_pyramid, weights=None):
    """sum_s w_s * mean |LRwarp_error(imgL, disp_s, imgR)| over a list of disparity maps at any resolutions.

    `weights` defaults to upstream's pyramid weights when there are four scales, else to ones.  Disparities are in
    pixels of their own resolution, as LRwarp_error expects."""
    if isinstance(disp_pyramid, torch.Tensor):
        disp_pyramid = [disp_pyramid]
    if weights is None:
        weights = PYRAMID_WEIGHTS if len(disp_pyramid) == len(PYRAMID_WEIGHTS) else (1.0,) * len(disp_pyramid)
    if len(weights) != len(disp_pyramid):
        raise ValueError(f"photometric_loss: {len(disp_pyramid)} disparity maps but {len(weights)} weights")
    total = None
    for w, d in zip(weights, disp_pyramid):
        term = photometric_l1(imgL, d, imgR) * float(w)
        total = term if total is None else total + term
    return total
