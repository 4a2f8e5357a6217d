"""On-GPU haze synthesis of the training loop's input side.

Mirrors, name for name, the functions upstream's trainers import from DeBug/inference.py
(trainfiles/dssmd_traner_new.py:22) and call once per batch (:250-256):

    convert_disp_to_depth(baseline, focal_length, disp)      DeBug/inference.py:57-64
    depth2trans(depth, beta)                                 DeBug/inference.py:126-134
    recover_haze_images(clean_images, beta, A, depth)        DeBug/inference.py:95-112

each as ONE kernel launch instead of 2 / 3 / 8 ATen kernels, plus `synthesize_haze`, which does the whole chain
(disparity -> depth -> transmission -> hazy image) in one launch.  Results follow upstream's type promotion: the data
loaders hand the per-sample scalars over as float64, so float32 images give float64 results (upstream then calls
`.float()`).  Upstream's `assert x.shape[0] == b` checks are kept (AssertionError).  The kernels are forward-only
(upstream calls these functions on `requires_grad=False` data): an input that requires a gradient raises
NotImplementedError here.  `install()` binds `guarded(kernel_fn, upstream_fn)` in upstream's modules, which hands such
calls (and non-CUDA / non-fp32/fp64 tensors) to upstream's own saved function instead.
"""
from __future__ import annotations

import torch

from . import _lib

__all__ = ["convert_disp_to_depth", "depth2trans", "recover_haze_images", "synthesize_haze", "kernel_supports", "guarded"]


def _needs_grad(*ts):
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in ts)


def _no_grad_path(name, *ts):
    if _needs_grad(*ts):
        raise NotImplementedError(f"{name}: the kernel is forward-only and an input requires a gradient "
                                  "(upstream calls it on requires_grad=False data)")


def kernel_supports(*tensors):
    """True when `dssmd_haze_fwd` takes these arguments: CUDA float32 / float64 tensors that carry no gradient."""
    ts = [t for t in tensors if isinstance(t, torch.Tensor)]
    return (len(ts) == len(tensors) and all(t.is_cuda and t.dtype in (torch.float32, torch.float64) for t in ts)
            and not _needs_grad(*ts))


def guarded(kernel_fn, upstream_fn):
    """What install() binds in upstream's modules: the kernel for arguments it takes, upstream's saved function otherwise."""
    import functools
    import inspect
    sig = inspect.signature(kernel_fn)   # same parameter names as upstream's: the trainers call these with keywords

    @functools.wraps(kernel_fn)
    def fn(*args, **kwargs):
        try:
            bound = sig.bind(*args, **kwargs).args
        except TypeError:
            return upstream_fn(*args, **kwargs)   # upstream raises its own error
        if not kernel_supports(*bound):
            return upstream_fn(*args, **kwargs)
        return kernel_fn(*bound)
    fn.__wrapped_upstream__ = upstream_fn
    fn.__kernel__ = kernel_fn
    return fn


def _compute_dtype(name, images, scalars):
    dt = images[0].dtype
    for t in list(images[1:]) + list(scalars):
        dt = torch.promote_types(dt, t.dtype)
    if dt not in (torch.float32, torch.float64):
        raise NotImplementedError(f"{name}: only float32 / float64 arithmetic is implemented (promoted dtype {dt})")
    return dt


def _scalar(name, t, b, dt):
    """Per-sample scalar [B] (any shape with B elements, as upstream's .view(b,1,1,1) accepts) in the compute dtype."""
    assert t.shape[0] == b, f"{name}: expected {b} per-sample values, got shape {tuple(t.shape)}"
    if t.numel() != b:
        raise RuntimeError(f"{name}: per-sample scalars must have exactly B = {b} elements, got {tuple(t.shape)}")
    return t.reshape(b).to(dt).contiguous()


def _launch(name, clear, src, src_is_disp, baseline, focal, beta, airlight, want_haze, want_trans, want_depth):
    tensors = [t for t in (clear, src, baseline, focal, beta, airlight) if t is not None]
    _lib.require_cuda(name, *tensors)
    if src.dim() != 4:
        raise RuntimeError(f"{name}: expected a 4-D [B,C,H,W] tensor, got {tuple(src.shape)}")
    images = [t for t in (clear, src) if t is not None]
    scalars = [t for t in (baseline, focal, beta, airlight) if t is not None]
    cdt = _compute_dtype(name, images, scalars)
    idt = torch.float64 if any(t.dtype == torch.float64 for t in images) else torch.float32
    if any(t.dtype not in (torch.float32, torch.float64) for t in images):
        raise NotImplementedError(f"{name}: image tensors must be float32 or float64")
    b, cs, h, w = src.shape
    # per-sample scalars make every channel of src equivalent: fold them into the rows
    srcc = src.to(idt).contiguous().view(b, 1, cs * h, w)
    hh = cs * h
    C = 0
    clearc = None
    if clear is not None:
        if clear.dim() != 4 or clear.shape[0] != b or tuple(clear.shape[-2:]) != (h, w) or cs != 1:
            raise RuntimeError(f"{name}: clean images {tuple(clear.shape)} do not match a [B,1,H,W] depth {tuple(src.shape)}")
        clearc = clear.to(idt).contiguous()
        C = clear.shape[1]
    sc = {k: (_scalar(name, t, b, cdt) if t is not None else None)
          for k, t in (("baseline", baseline), ("focal", focal), ("beta", beta), ("airlight", airlight))}
    dev = src.device
    haze = torch.empty((b, C, h, w), dtype=cdt, device=dev) if want_haze else None
    trans = torch.empty((b, cs, h, w), dtype=cdt, device=dev) if want_trans else None
    depth = torch.empty((b, cs, h, w), dtype=cdt, device=dev) if want_depth else None
    if src.numel() > 0:
        with _lib.device_guard(dev):
            _lib.check(_lib.lib().dssmd_haze_fwd(
                _lib.ptr(clearc), _lib.ptr(srcc), 1 if src_is_disp else 0, _lib.ptr(sc["baseline"]), _lib.ptr(sc["focal"]),
                _lib.ptr(sc["beta"]), _lib.ptr(sc["airlight"]), _lib.ptr(haze), _lib.ptr(trans), _lib.ptr(depth),
                b, C, hh, w, _lib.DTYPE_CODE[idt], _lib.DTYPE_CODE[cdt], _lib.stream_ptr(dev)))
    return haze, trans, depth


def convert_disp_to_depth(baseline, focal_length, disp):
    """depth = focal_length * baseline / disp  (DeBug/inference.py:57-64)."""
    b = disp.shape[0]
    assert baseline.shape[0] == b
    assert focal_length.shape[0] == b
    _no_grad_path("convert_disp_to_depth", baseline, focal_length, disp)
    return _launch("convert_disp_to_depth", None, disp, True, baseline, focal_length, None, None, False, False, True)[2]


def depth2trans(depth, beta):
    """transmission = exp(-depth * beta)  (DeBug/inference.py:126-134)."""
    b = depth.shape[0]
    assert beta.shape[0] == b
    _no_grad_path("depth2trans", depth, beta)
    return _launch("depth2trans", None, depth, False, None, None, beta, None, False, True, False)[1]


def recover_haze_images(clean_images, beta, A, depth):
    """foggy = clean * t + A * (1 - t), t = exp(-depth * beta)  (DeBug/inference.py:95-112)."""
    b = clean_images.shape[0]
    assert beta.shape[0] == b
    assert A.shape[0] == b
    _no_grad_path("recover_haze_images", clean_images, beta, A, depth)
    return _launch("recover_haze_images", clean_images, depth, False, None, None, beta, A, True, False, False)[0]


def synthesize_haze(clean_images, disp, baseline, focal_length, beta, A):
    """The three calls of the training loop (trainfiles/dssmd_traner_new.py:250-256) in one launch:
    returns (haze, transmission, depth) for one view."""
    b = clean_images.shape[0]
    assert baseline.shape[0] == b and focal_length.shape[0] == b and beta.shape[0] == b and A.shape[0] == b
    _no_grad_path("synthesize_haze", clean_images, disp, baseline, focal_length, beta, A)
    return _launch("synthesize_haze", clean_images, disp, True, baseline, focal_length, beta, A, True, True, True)
