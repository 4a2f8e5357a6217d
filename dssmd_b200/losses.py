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

__all__ = ["photometric_l1", "photometric_loss", "PYRAMID_WEIGHTS", "RecoveredCleanImagesLoss", "RecoveredCleanImagesLossV2",
           "recovered_clean_l1", "asm_kernel_supports", "guarded_asm_forward"]

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
        flag, token = _new_flag(Lc.device)
        with _lib.device_guard(Lc.device):
            _lib.check(_lib.lib().dssmd_photometric_l1_fwd(_lib.ptr(Lc), _lib.ptr(Rc), _lib.ptr(Dc), _lib.ptr(g_unit), _lib.ptr(rows),
                                                           flag, B, C, H, W, _lib.dtype_code(Lc), _lib.stream_ptr(Lc.device)))
        _raise_if_flag(token, Lc.device)
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


# ---- ASM reconstruction loss -------------------------------------------------------------------
_ASM_BLOCKS = {}   # (B, h, w, device index) -> blocks of the launch (a property of the shape and the device's SM count)


class _RecoveredCleanL1Fn(torch.autograd.Function):
    """mean |clamp((haze_s - A (1 - t)) / t, 0, 1) - clean_s| with d/dt and d/dA computed in the forward launch."""

    @staticmethod
    def forward(ctx, trans, airlight, haze, clean):
        B, C, H0, W0 = haze.shape
        h, w = trans.shape[-2:]
        T, A = trans.contiguous(), airlight.reshape(B).contiguous()
        Hz, Cl = haze.contiguous(), clean.contiguous()
        lib = _lib.lib()
        need_t = ctx.needs_input_grad[0]
        g_unit = torch.empty_like(T) if need_t else None
        with _lib.device_guard(T.device):   # the block count depends on the SM count of the tensors' device
            key = (B, h, w, T.device.index)
            nblk = _ASM_BLOCKS.get(key)
            if nblk is None:
                nblk = _ASM_BLOCKS[key] = int(lib.dssmd_asm_recovered_l1_blocks(B, h, w))
            # one buffer: [B, nblk, 2] block partials, then 1 + B results (loss, d sum|diff| / d airlight[b]) reduced on the device
            buf = torch.empty(B * nblk * 2 + 1 + B, dtype=torch.float32, device=T.device)
            out = buf[B * nblk * 2:]
            _lib.check(lib.dssmd_asm_recovered_l1_loss(_lib.ptr(T), _lib.ptr(A), _lib.ptr(Hz), _lib.ptr(Cl), _lib.ptr(g_unit),
                                                       _lib.ptr(buf), _lib.ptr(out), B, C, H0, W0, h, w, _lib.dtype_code(T),
                                                       _lib.stream_ptr(T.device)))
        ctx.n = float(B * C * h * w)
        ctx.a_shape = tuple(airlight.shape)
        ctx.save_for_backward(g_unit, out[1:])
        return out[0]

    @staticmethod
    def backward(ctx, g):
        g_unit, ga = ctx.saved_tensors
        scale = g / ctx.n
        g_t = g_unit * scale if ctx.needs_input_grad[0] else None
        g_a = (ga * scale).reshape(ctx.a_shape) if ctx.needs_input_grad[1] else None
        return g_t, g_a, None, None


def asm_kernel_supports(transmission_map, airlight, haze_image, clean_image):
    """True when `dssmd_asm_recovered_l1_fwd` takes this level: CUDA float32 tensors, no gradient into the images, image
    size an integer multiple of the transmission map's."""
    ts = (transmission_map, airlight, haze_image, clean_image)
    if not all(isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 for t in ts) or torch.is_autocast_enabled():
        return False
    if transmission_map.dim() != 4 or haze_image.dim() != 4 or clean_image.shape != haze_image.shape:
        return False
    if torch.is_grad_enabled() and (haze_image.requires_grad or clean_image.requires_grad):
        return False
    B, _, H0, W0 = haze_image.shape
    h, w = transmission_map.shape[-2:]
    if tuple(transmission_map.shape) != (B, 1, h, w) or airlight.numel() != B or airlight.shape[0] != B or w == 0 or h == 0:
        return False
    scale = W0 // w
    return scale >= 1 and W0 == w * scale and H0 == h * scale


def guarded_asm_forward(upstream_forward):
    """The `forward` install() puts on upstream's RecoveredCleanImagesLoss / V2: the fused launch for levels the kernel
    takes, upstream's saved forward for everything else (other dtypes, AMP, image gradients, non-integer scales)."""
    def forward(self, transmission_map, airlight, haze_image, clean_image):
        if self.loss_type != 'normal' or not asm_kernel_supports(transmission_map, airlight, haze_image, clean_image):
            return upstream_forward(self, transmission_map, airlight, haze_image, clean_image)
        return recovered_clean_l1(transmission_map, airlight, haze_image, clean_image)
    forward.__wrapped_upstream__ = upstream_forward
    return forward


def recovered_clean_l1(transmission_map, airlight, haze_image, clean_image):
    """One level of upstream's reconstruction loss (losses/DSSMDLoss.py:87-111) as one kernel launch: the clean image is
    recovered through the atmospheric scattering model from the predicted transmission [B,1,h,w] and airlight (B values),
    clamped to [0, 1] and compared (L1 mean) with the clean image, both images bilinearly downsampled to the
    transmission map's size first.  Gradients flow to the transmission map and the airlight."""
    _lib.require_cuda("RecoveredCleanImagesLoss", transmission_map, airlight, haze_image, clean_image)
    _check4("RecoveredCleanImagesLoss", transmission_map)
    _check4("RecoveredCleanImagesLoss", haze_image)
    _check4("RecoveredCleanImagesLoss", clean_image)
    B, C, H0, W0 = haze_image.shape
    h, w = transmission_map.shape[-2:]
    if tuple(transmission_map.shape) != (B, 1, h, w):
        raise RuntimeError(f"RecoveredCleanImagesLoss: transmission map must be [B,1,h,w], got {tuple(transmission_map.shape)}")
    if clean_image.shape != haze_image.shape:
        raise RuntimeError(f"RecoveredCleanImagesLoss: haze {tuple(haze_image.shape)} and clean {tuple(clean_image.shape)} images differ")
    if airlight.numel() != B or airlight.shape[0] != B:
        raise RuntimeError(f"RecoveredCleanImagesLoss: airlight must hold one value per sample, got {tuple(airlight.shape)}")
    if any(t.dtype != torch.float32 for t in (transmission_map, airlight, haze_image, clean_image)):
        raise NotImplementedError("RecoveredCleanImagesLoss: only float32 is implemented")
    if haze_image.requires_grad or clean_image.requires_grad:
        raise NotImplementedError("RecoveredCleanImagesLoss: gradients w.r.t. the images are not implemented "
                                  "(upstream feeds the synthesised hazy input and the ground-truth clean image)")
    scale = W0 // w if w else 0
    if scale < 1 or W0 != w * scale or H0 != h * scale:
        raise NotImplementedError(f"RecoveredCleanImagesLoss: images {H0}x{W0} must be an integer multiple of the "
                                  f"transmission map {h}x{w}")
    return _RecoveredCleanL1Fn.apply(transmission_map, airlight, haze_image, clean_image)


class RecoveredCleanImagesLoss(torch.nn.Module):
    """Same constructor and forward signature as upstream's losses/DSSMDLoss.py:81-111 (airlight: [B,1])."""

    def __init__(self, type='normal'):  # noqa: A002  (upstream's argument name)
        super().__init__()
        self.loss_type = type

    def forward(self, transmission_map, airlight, haze_image, clean_image):
        if self.loss_type != 'normal':   # upstream: `loss` is unbound for any other type
            raise UnboundLocalError("local variable 'loss' referenced before assignment")
        return recovered_clean_l1(transmission_map, airlight, haze_image, clean_image)


class RecoveredCleanImagesLossV2(RecoveredCleanImagesLoss):
    """losses/DSSMDLoss.py:115-145: as above, the airlight already broadcastable ([B,1,1,1])."""
