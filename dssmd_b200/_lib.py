"""ctypes binding of libdssmd_b200.so (the C ABI declared in include/dssmd_b200.h).

There is deliberately no fallback: if the shared library has not been built, or
a tensor is not on a CUDA device, the ops raise.  Build with
`python -m dssmd_b200.build` (or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdssmd_b200.so")

# dtype codes of include/dssmd_b200.h
DTYPE_CODE = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2, torch.float64: 3}
PAD_CODE = {"zeros": 0, "border": 1}
ERR_INVALID_ARGUMENT, ERR_UNSUPPORTED, ERR_CUDA = 1, 2, 3

_vp, _i = ctypes.c_void_p, ctypes.c_int
# name -> argument types, exactly the prototypes of include/dssmd_b200.h
SIGNATURES = {
    "dssmd_corr_fwd": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "dssmd_corr_bwd": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "dssmd_warp_fwd": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "dssmd_warp_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_size_t, _i, _i, _i, _i, _i, _i, _vp],
    "dssmd_resize_bilinear_fwd": [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "dssmd_resize_bilinear_bwd": [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "dssmd_lrwarp_error_fwd": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "dssmd_photometric_l1_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "dssmd_selftest_division": [_i, _i, _vp, _vp],
    "dssmd_corr_fwd_cat": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "dssmd_corr_bwd_cat": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "dssmd_asm_recovered_l1_blocks": [_i, _i, _i],
    "dssmd_asm_recovered_l1_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "dssmd_asm_recovered_l1_loss": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "dssmd_epipolar_attn_fwd": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, ctypes.c_float, _i, _vp],
    "dssmd_epipolar_attn_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, ctypes.c_float, _i, _vp],
    "dssmd_haze_fwd": [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
}

_lib = None
_lock = threading.Lock()


class LibraryNotBuiltError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Load (once) and return the shared library; raises if it is missing."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise LibraryNotBuiltError(
                        f"{LIB_PATH} not found: the CUDA extension is not built and dssmd_b200 has no "
                        "CPU/PyTorch fallback. Run `python -m dssmd_b200.build`.")
                handle = ctypes.CDLL(LIB_PATH)
                handle.dssmd_version.restype = ctypes.c_int
                handle.dssmd_last_error.restype = ctypes.c_char_p
                handle.dssmd_warp_bwd_workspace_bytes.argtypes = [_i, _i, _i, _i, _i]
                handle.dssmd_warp_bwd_workspace_bytes.restype = ctypes.c_size_t
                for name, argtypes in SIGNATURES.items():
                    fn = getattr(handle, name)
                    fn.argtypes = argtypes
                    fn.restype = ctypes.c_int
                _lib = handle
    return _lib


def last_error() -> str:
    msg = lib().dssmd_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int) -> None:
    """Map a non-zero return code to the Python exception the shim promises."""
    if rc == 0:
        return
    msg = last_error()
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)


def ptr(t):
    """Device address of a tensor as ctypes takes it for a void* argument (a plain int, None = NULL): no c_void_p
    object per argument on the per-call path."""
    return None if t is None else t.data_ptr()


try:   # raw handle of the current stream without building a torch.cuda.Stream object (~2 us per call)
    _raw_stream = torch._C._cuda_getCurrentRawStream
except AttributeError:  # pragma: no cover
    _raw_stream = None


def stream_ptr(device):
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if _raw_stream is not None:
        return _raw_stream(idx) or None
    return torch.cuda.current_stream(device).cuda_stream or None


class _NoGuard:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def device_guard(device):
    """`with torch.cuda.device(device)` only when the tensors' device is not the current one (the context manager costs
    ~8 us per call; one process per GPU and nn.DataParallel replica threads are already on their device)."""
    idx = device.index
    if idx is None or idx == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(idx)


def dtype_code(t: torch.Tensor) -> int:
    try:
        return DTYPE_CODE[t.dtype]
    except KeyError:
        raise RuntimeError(f"dssmd_b200: unsupported dtype {t.dtype}") from None


def require_cuda(name: str, *tensors: torch.Tensor) -> None:
    dev = None
    for t in tensors:
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
        if not t.is_cuda:
            raise RuntimeError(f"{name}: tensor is on {t.device}; dssmd_b200 runs on CUDA devices only "
                               "(there is no CPU fallback)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"{name}: tensors are on different devices ({dev} vs {t.device})")
