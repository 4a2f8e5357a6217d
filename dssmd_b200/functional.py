"""autograd wrappers around the C ABI.  Plumbing only: allocate outputs with
torch (so the caching allocator and autograd see them), pick the device/stream
of the input tensors, call the kernel, map errors."""
from __future__ import annotations

import os
import threading

import torch

from . import _lib

_CHECK_NEGATIVE = True      # True: upstream's behaviour (raise at the call); "deferred": raise at a later call; False: no check


def set_check_negative_disparity(flag) -> None:
    """Upstream's `assert disp.min() >= 0` (utils/disparity_warper.py:93) costs a device->host sync per warp.

    True (default)  same behaviour: the call waits for its kernel and raises AssertionError itself.
    "deferred"      no host wait inside the call: the kernel's verdict is read at a LATER disp_warp / LRwarp_error /
                    photometric call on the same device and thread (or by check_negative_disparity_now()), and the
                    AssertionError is raised there, naming how many calls ago the violation happened.
    False           no check (callers who guarantee non-negative disparities)."""
    global _CHECK_NEGATIVE
    if flag not in (True, False, "deferred"):
        raise ValueError("set_check_negative_disparity: expected True, False or 'deferred'")
    _CHECK_NEGATIVE = flag


def _check4(name, t):
    if t.dim() != 4:
        raise RuntimeError(f"{name}: expected a 4-D [B,C,H,W] tensor, got shape {tuple(t.shape)}")


class CorrelationFn(torch.autograd.Function):
    """out[b,d,y,x] = mean_c L[b,c,y,x] * R[b,c,y,x-d] (upstream build_corr)."""

    @staticmethod
    def forward(ctx, left, right, max_disp):
        _lib.require_cuda("build_corr", left, right)
        _check4("build_corr", left); _check4("build_corr", right)
        if left.shape != right.shape:
            raise RuntimeError(f"build_corr: left {tuple(left.shape)} and right {tuple(right.shape)} differ")
        if left.dtype != right.dtype:
            raise RuntimeError(f"build_corr: left is {left.dtype}, right is {right.dtype}")
        D = int(max_disp)
        B, C, H, W = left.shape
        L, R = left.contiguous(), right.contiguous()
        out = torch.empty((B, max(D, 0), H, W), dtype=L.dtype, device=L.device)
        if out.numel() > 0 and C > 0:
            with _lib.device_guard(L.device):
                _lib.check(_lib.lib().dssmd_corr_fwd(_lib.ptr(L), _lib.ptr(R), _lib.ptr(out), B, C, H, W, D,
                                                     _lib.dtype_code(L), _lib.stream_ptr(L.device)))
        elif out.numel() > 0:
            out.fill_(float("nan"))  # mean over zero channels
        ctx.save_for_backward(L, R)
        ctx.D = D
        return out

    @staticmethod
    def backward(ctx, g):
        L, R = ctx.saved_tensors
        B, C, H, W = L.shape
        need_l, need_r = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gL = torch.empty_like(L) if need_l else None
        gR = torch.empty_like(R) if need_r else None
        if (need_l or need_r) and L.numel() > 0:
            if ctx.D == 0:
                for t in (gL, gR):
                    if t is not None:
                        t.zero_()
            else:
                g = g.contiguous()
                with _lib.device_guard(L.device):
                    _lib.check(_lib.lib().dssmd_corr_bwd(_lib.ptr(g), _lib.ptr(L), _lib.ptr(R), _lib.ptr(gL), _lib.ptr(gR),
                                                         B, C, H, W, ctx.D, _lib.dtype_code(L), _lib.stream_ptr(L.device)))
        return gL, gR, None


class CorrelationCatFn(torch.autograd.Function):
    """cat((head, build_corr(left, right, max_disp)), dim=1) with the correlation kernel writing its channels directly
    into the concat buffer, and the backward reading them out of the concat gradient in place (no volume round trip,
    no slice copy).  Upstream pattern: models/DSSMD.py:129-131."""

    @staticmethod
    def forward(ctx, head, left, right, max_disp):
        _lib.require_cuda("build_corr_cat", head, left, right)
        for t in (head, left, right):
            _check4("build_corr_cat", t)
        if left.shape != right.shape:
            raise RuntimeError(f"build_corr_cat: left {tuple(left.shape)} and right {tuple(right.shape)} differ")
        if left.dtype != right.dtype or head.dtype != left.dtype:
            raise RuntimeError(f"build_corr_cat: dtypes differ ({head.dtype}, {left.dtype}, {right.dtype})")
        B, C, H, W = left.shape
        if head.shape[0] != B or tuple(head.shape[2:]) != (H, W):   # torch.cat's own requirement
            raise RuntimeError(f"build_corr_cat: Sizes of tensors must match except in dimension 1 "
                               f"(head {tuple(head.shape)}, features {tuple(left.shape)})")
        D = max(int(max_disp), 0)
        Ch = head.shape[1]
        L, R = left.contiguous(), right.contiguous()
        out = torch.empty((B, Ch + D, H, W), dtype=L.dtype, device=L.device)
        out[:, :Ch].copy_(head)
        if D > 0 and out.numel() > 0 and C > 0:
            with _lib.device_guard(L.device):
                _lib.check(_lib.lib().dssmd_corr_fwd_cat(_lib.ptr(L), _lib.ptr(R), _lib.ptr(out), B, C, H, W, D, Ch + D, Ch,
                                                         _lib.dtype_code(L), _lib.stream_ptr(L.device)))
        elif D > 0 and out.numel() > 0:
            out[:, Ch:].fill_(float("nan"))  # mean over zero channels
        ctx.save_for_backward(L, R)
        ctx.D, ctx.Ch = D, Ch
        return out

    @staticmethod
    def backward(ctx, g):
        L, R = ctx.saved_tensors
        B, C, H, W = L.shape
        D, Ch = ctx.D, ctx.Ch
        need_h, need_l, need_r = ctx.needs_input_grad[:3]
        g = g.contiguous()
        g_head = g[:, :Ch] if need_h else None
        gL = torch.empty_like(L) if need_l else None
        gR = torch.empty_like(R) if need_r else None
        if (need_l or need_r) and L.numel() > 0:
            if D == 0:
                for t in (gL, gR):
                    if t is not None:
                        t.zero_()
            else:
                with _lib.device_guard(L.device):
                    _lib.check(_lib.lib().dssmd_corr_bwd_cat(_lib.ptr(g), _lib.ptr(L), _lib.ptr(R), _lib.ptr(gL), _lib.ptr(gR),
                                                             B, C, H, W, D, Ch + D, Ch, _lib.dtype_code(L), _lib.stream_ptr(L.device)))
        return g_head, gL, gR, None


def _pad_code(padding_mode):
    if padding_mode not in ("zeros", "border", "reflection"):
        raise ValueError(f"nn.functional.grid_sample(): expected padding_mode to be 'zeros', 'border', or "
                         f"'reflection', but got: '{padding_mode}'")
    if padding_mode == "reflection":
        raise NotImplementedError("disp_warp: padding_mode='reflection' is not implemented by the CUDA kernel "
                                  "(upstream documents 'zeros' and 'border' only)")
    return _lib.PAD_CODE[padding_mode]


def _check_warp_args(name, img, disp):
    _lib.require_cuda(name, img, disp)
    _check4(name, img); _check4(name, disp)
    if img.dtype != disp.dtype:
        raise RuntimeError(f"{name}: expected img and disp to have the same dtype, got {img.dtype} and {disp.dtype}")
    B, C, H, W = img.shape
    if tuple(disp.shape) != (B, 1, H, W):
        raise RuntimeError(f"{name}: disp must be [B,1,H,W]={(B, 1, H, W)}, got {tuple(disp.shape)}")
    if img.dtype not in (torch.float32, torch.float64):
        raise NotImplementedError(f"{name}: dtype {img.dtype} not supported (the sampling coordinates need fp32; "
                                  "upstream's own 16-bit path is numerically meaningless at image widths)")


def _band_eligible(B, C, H, W, tensors):
    """Shapes the one-pass band kernel of dssmd_warp_bwd takes (csrc/warp_bwd_band2.cu: image-like channel counts, rows of
    whole float4s, 16-byte aligned tensors): it needs no workspace, so none is allocated."""
    if os.environ.get("DSSMD_TUNING") and int(os.environ.get("DSSMD_WARP_BWD_VARIANT") or 7) < 6:
        return False   # a development override forces one of the paths that want the workspace
    return (1 <= C <= 4 and W % 4 == 0 and 4 <= W <= 2048 and H >= 2 and B <= 65535 and C * H * W < 2 ** 31
            and all(t is None or t.data_ptr() % 16 == 0 for t in tensors))


def _warp_bwd(g, I, Dp, g_img, g_disp, pad):
    """dssmd_warp_bwd; the scratch buffer only for the shapes whose fallback (two-kernel) path wants one."""
    B, C, H, W = I.shape
    lib = _lib.lib()
    code = _lib.dtype_code(I)
    nbytes = 0
    if g_img is not None and not (I.dtype == torch.float32 and _band_eligible(B, C, H, W, (g, I, Dp, g_img, g_disp))):
        nbytes = int(lib.dssmd_warp_bwd_workspace_bytes(B, C, H, W, code))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=I.device) if nbytes else None
    with _lib.device_guard(I.device):
        _lib.check(lib.dssmd_warp_bwd(_lib.ptr(g), _lib.ptr(I), _lib.ptr(Dp), _lib.ptr(g_img), _lib.ptr(g_disp),
                                      _lib.ptr(ws), nbytes, B, C, H, W, pad, code, _lib.stream_ptr(I.device)))


# ---- the disp >= 0 check (upstream utils/disparity_warper.py:93) -------------------------------------------------------
# The kernel sets a flag in PINNED HOST memory with a plain store (unified addressing), so reading it needs an event wait on
# the launch stream, not a device->host copy that would queue behind whatever the copy engine is moving.  One flag ring and
# one event per (thread, device): nothing is allocated on the per-call path.
_RING = 64
_tls = threading.local()


class _FlagState:
    def __init__(self):
        self.flags = torch.zeros(_RING, dtype=torch.int32).pin_memory()
        self.base = self.flags.data_ptr()
        self.event = torch.cuda.Event()
        self.pending = []       # deferred mode: (slot, event, serial)
        self.serial = 0
        self.free_events = []


def _flag_state(device):
    states = getattr(_tls, "states", None)
    if states is None:
        states = _tls.states = {}
    st = states.get(device.index)
    if st is None:
        st = states[device.index] = _FlagState()
    return st


def _drain(st, wait):
    """Deferred mode: read the verdicts of the earlier calls whose kernels have finished (all of them if `wait`)."""
    bad = None
    while st.pending:
        slot, ev, serial = st.pending[0]
        if wait:
            ev.synchronize()
        elif not ev.query():
            break
        st.pending.pop(0)
        st.free_events.append(ev)
        if int(st.flags[slot]) != 0 and bad is None:
            bad = serial
    if bad is not None:
        ago = st.serial - bad
        st.pending.clear()
        raise AssertionError(f"disp.min() >= 0  (deferred check: violated by the warp call {ago} call(s) ago)")


def check_negative_disparity_now(device=None) -> None:
    """Deferred mode: wait for the outstanding warp kernels of this thread and raise if any saw a negative / NaN disparity."""
    states = getattr(_tls, "states", None) or {}
    for idx, st in states.items():
        if device is None or torch.device(device).index == idx:
            _drain(st, wait=True)


def _new_flag(device):
    """-> (device-visible address of this call's flag or None, token for _raise_if_flag)."""
    if _CHECK_NEGATIVE is False or torch.cuda.is_current_stream_capturing():
        return None, None    # a graph replay cannot be interrupted by a host-side read-back: no check while capturing
    st = _flag_state(device)
    if _CHECK_NEGATIVE is True:
        st.flags[0] = 0
        return st.base, (st, 0)
    _drain(st, wait=len(st.pending) >= _RING - 1)
    st.serial += 1
    slot = 1 + st.serial % (_RING - 1)
    st.flags[slot] = 0
    return st.base + 4 * slot, (st, slot)


def _raise_if_flag(token, device):
    if token is None:
        return
    st, slot = token
    if slot == 0:   # upstream's behaviour: wait for the kernel, raise here
        st.event.record(torch.cuda.current_stream(device))
        st.event.synchronize()
        if int(st.flags[0]) != 0:
            raise AssertionError("disp.min() >= 0")  # upstream utils/disparity_warper.py:93
        return
    ev = st.free_events.pop() if st.free_events else torch.cuda.Event()
    ev.record(torch.cuda.current_stream(device))
    st.pending.append((slot, ev, st.serial))


class DispWarpFn(torch.autograd.Function):
    """(warped, valid_mask) = disp_warp(img, disp, padding_mode)."""

    @staticmethod
    def forward(ctx, img, disp, padding_mode):
        pad = _pad_code(padding_mode)
        _check_warp_args("disp_warp", img, disp)
        B, C, H, W = img.shape
        I, Dp = img.contiguous(), disp.contiguous()
        warped = torch.empty_like(I)
        mask = torch.empty_like(I)
        if I.numel() > 0:
            flag, token = _new_flag(I.device)
            with _lib.device_guard(I.device):
                _lib.check(_lib.lib().dssmd_warp_fwd(_lib.ptr(I), _lib.ptr(Dp), _lib.ptr(warped), _lib.ptr(mask), flag,
                                                     B, C, H, W, pad, _lib.dtype_code(I), _lib.stream_ptr(I.device)))
            _raise_if_flag(token, I.device)
        ctx.save_for_backward(I, Dp)
        ctx.pad = pad
        ctx.mark_non_differentiable(mask)
        return warped, mask

    @staticmethod
    def backward(ctx, g_warped, _g_mask):
        I, Dp = ctx.saved_tensors
        B, C, H, W = I.shape
        need_i, need_d = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        g_img = torch.empty_like(I) if need_i else None
        g_disp = torch.empty_like(Dp) if need_d else None
        if (need_i or need_d) and I.numel() > 0:
            _warp_bwd(g_warped.contiguous(), I, Dp, g_img, g_disp, ctx.pad)
        return g_img, g_disp, None


def resize_bilinear(img, size):
    """F.interpolate(img, size=size, mode='bilinear', align_corners=False), forward only."""
    _lib.require_cuda("resize_bilinear", img)
    B, C, Hi, Wi = img.shape
    Ho, Wo = int(size[0]), int(size[1])
    src = img.contiguous()
    out = torch.empty((B, C, Ho, Wo), dtype=img.dtype, device=img.device)
    with _lib.device_guard(img.device):
        _lib.check(_lib.lib().dssmd_resize_bilinear_fwd(_lib.ptr(src), _lib.ptr(out), B, C, Hi, Wi, Ho, Wo,
                                                        _lib.dtype_code(src), _lib.stream_ptr(img.device)))
    return out


def _resize_bilinear_bwd(g_out, in_size):
    B, C, Ho, Wo = g_out.shape
    Hi, Wi = in_size
    g = g_out.contiguous()
    g_in = torch.empty((B, C, Hi, Wi), dtype=g.dtype, device=g.device)
    with _lib.device_guard(g.device):
        _lib.check(_lib.lib().dssmd_resize_bilinear_bwd(_lib.ptr(g), _lib.ptr(g_in), B, C, Hi, Wi, Ho, Wo,
                                                        _lib.dtype_code(g), _lib.stream_ptr(g.device)))
    return g_in


class ResizeBilinearFn(torch.autograd.Function):
    """F.interpolate(img, size=size, mode='bilinear', align_corners=False) with its backward (the two kernels of the fused
    LRwarp_error path, usable on their own)."""

    @staticmethod
    def forward(ctx, img, size):
        ctx.in_size = tuple(img.shape[-2:])
        return resize_bilinear(img, size)

    @staticmethod
    def backward(ctx, g):
        return _resize_bilinear_bwd(g, ctx.in_size), None


def lrwarp_error(imgL, disp, imgR):
    """Upstream LRwarp_error (utils/disparity_warper.py:109-115).  Each image that is WIDER than `disp` is resized to
    disp's size on its own (:110-113), so the two images need not have the same shape; when they do, one fused launch."""
    if isinstance(imgL, torch.Tensor) and isinstance(imgR, torch.Tensor) and imgL.dim() == 4 and imgR.dim() == 4 \
            and imgL.shape != imgR.shape and isinstance(disp, torch.Tensor) and disp.dim() == 4:
        _lib.require_cuda("LRwarp_error", imgL, disp, imgR)
        size = tuple(disp.shape[-2:])
        if imgL.size(-1) > disp.size(-1):
            imgL = ResizeBilinearFn.apply(imgL, size)
        if imgR.size(-1) > disp.size(-1):
            imgR = ResizeBilinearFn.apply(imgR, size)
        warped, _ = DispWarpFn.apply(imgL, disp, 'border')
        return imgR - warped
    return LRWarpErrorFn.apply(imgL, disp, imgR)


class LRWarpErrorFn(torch.autograd.Function):
    """imgR' - warp(imgL', disp): upstream LRwarp_error, one fused launch forward."""

    @staticmethod
    def forward(ctx, imgL, disp, imgR):
        _lib.require_cuda("LRwarp_error", imgL, disp, imgR)
        for t in (imgL, disp, imgR):
            _check4("LRwarp_error", t)
        if imgL.dtype != disp.dtype or imgR.dtype != disp.dtype:
            raise RuntimeError("LRwarp_error: imgL, disp and imgR must share a dtype")
        if imgL.dtype not in (torch.float32, torch.float64):
            raise NotImplementedError(f"LRwarp_error: dtype {imgL.dtype} not supported")
        if imgL.shape != imgR.shape:
            raise RuntimeError(f"LRwarp_error: imgL {tuple(imgL.shape)} and imgR {tuple(imgR.shape)} differ")
        B, C, H0, W0 = imgL.shape
        Bd, one, H, W = disp.shape
        if Bd != B or one != 1:
            raise RuntimeError(f"LRwarp_error: disp must be [B,1,H,W], got {tuple(disp.shape)}")
        if W0 <= W and (H0, W0) != (H, W):
            raise RuntimeError(f"LRwarp_error: images {H0}x{W0} are not wider than disp {H}x{W} and do not match it")
        Lc, Rc, Dc = imgL.contiguous(), imgR.contiguous(), disp.contiguous()
        out = torch.empty((B, C, H, W), dtype=Lc.dtype, device=Lc.device)
        if out.numel() > 0:
            flag, token = _new_flag(Lc.device)
            with _lib.device_guard(Lc.device):
                _lib.check(_lib.lib().dssmd_lrwarp_error_fwd(_lib.ptr(Lc), _lib.ptr(Dc), _lib.ptr(Rc), _lib.ptr(out), flag,
                                                             B, C, H0, W0, H, W, _lib.dtype_code(Lc), _lib.stream_ptr(Lc.device)))
            _raise_if_flag(token, Lc.device)
        ctx.save_for_backward(Lc, Dc)
        ctx.in_size = (H0, W0)
        return out

    @staticmethod
    def backward(ctx, g):
        Lc, Dc = ctx.saved_tensors
        B, C, H0, W0 = Lc.shape
        H, W = Dc.shape[-2:]
        resized = W0 > W
        need_l, need_d, need_r = ctx.needs_input_grad
        g = g.contiguous()
        g_imgL = g_disp = g_imgR = None
        if need_r:
            g_imgR = _resize_bilinear_bwd(g, (H0, W0)) if resized else g
        if need_l or need_d:
            Ls = resize_bilinear(Lc, (H, W)) if resized else Lc
            gw = g.neg()  # d(err)/d(warped) = -1
            g_Ls = torch.empty_like(Ls) if need_l else None
            g_disp = torch.empty_like(Dc) if need_d else None
            _warp_bwd(gw, Ls, Dc, g_Ls, g_disp, _lib.PAD_CODE["border"])
            if need_l:
                g_imgL = _resize_bilinear_bwd(g_Ls, (H0, W0)) if resized else g_Ls
        return g_imgL, g_disp, g_imgR
