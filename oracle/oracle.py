"""ctypes/numpy front end of the CPU oracle (oracle/dssmd_oracle.c).

TEST INFRASTRUCTURE, NOT PRODUCT.  Importable only from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
dssmd_b200/ never imports this module.

Each function takes/returns C-contiguous numpy arrays (float32 or float64,
chosen by the dtype of the first argument) in the reference's NCHW layout and
restates one upstream function; see the C sources for the file:line citations.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libdssmd_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile).  Returns the .so path."""
    srcs = [os.path.join(_HERE, f) for f in ("dssmd_oracle.c", "dssmd_oracle_impl.h", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_num_threads.restype = ctypes.c_int
        _lib.oracle_version.restype = ctypes.c_int
    return _lib


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def set_num_threads(n: int) -> None:
    lib().oracle_set_num_threads(ctypes.c_int(int(n)))


def _suffix(a: np.ndarray) -> str:
    if a.dtype == np.float32:
        return "f32"
    if a.dtype == np.float64:
        return "f64"
    raise TypeError(f"oracle supports float32/float64, got {a.dtype}")


def _c(a: np.ndarray, dtype=None) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype if dtype is not None else a.dtype)


def _p(a: np.ndarray) -> ctypes.c_void_p:
    return ctypes.c_void_p(a.ctypes.data)


_I = ctypes.c_int


def corr_fwd(L: np.ndarray, R: np.ndarray, max_disp: int) -> np.ndarray:
    """build_corr(L, R, max_disp) -> [B, max_disp, H, W]."""
    L = _c(L); R = _c(R, L.dtype)
    B, C, H, W = L.shape
    out = np.empty((B, max_disp, H, W), dtype=L.dtype)
    getattr(lib(), "oracle_corr_fwd_" + _suffix(L))(_p(L), _p(R), _p(out), _I(B), _I(C), _I(H), _I(W), _I(max_disp))
    return out


def corr_bwd(g: np.ndarray, L: np.ndarray, R: np.ndarray):
    """Backward of build_corr: returns (gL, gR)."""
    L = _c(L); R = _c(R, L.dtype); g = _c(g, L.dtype)
    B, C, H, W = L.shape
    D = g.shape[1]
    gL = np.empty_like(L); gR = np.empty_like(R)
    getattr(lib(), "oracle_corr_bwd_" + _suffix(L))(_p(g), _p(L), _p(R), _p(gL), _p(gR),
                                                    _I(B), _I(C), _I(H), _I(W), _I(D))
    return gL, gR


_PAD = {"zeros": 0, "border": 1}


def warp_fwd(img: np.ndarray, disp: np.ndarray, padding_mode: str = "border"):
    """disp_warp(img, disp, padding_mode) -> (warped, valid_mask, negative_flag)."""
    img = _c(img); disp = _c(disp, img.dtype)
    B, C, H, W = img.shape
    assert disp.shape == (B, 1, H, W), disp.shape
    warped = np.empty_like(img); mask = np.empty_like(img)
    fn = getattr(lib(), "oracle_warp_fwd_" + _suffix(img))
    fn.restype = ctypes.c_int
    bad = fn(_p(img), _p(disp), _p(warped), _p(mask), _I(B), _I(C), _I(H), _I(W), _I(_PAD[padding_mode]))
    return warped, mask, bool(bad)


def warp_bwd(g: np.ndarray, img: np.ndarray, disp: np.ndarray, padding_mode: str = "border"):
    """Backward of disp_warp w.r.t. (img, disp) for upstream grad g of `warped`."""
    img = _c(img); disp = _c(disp, img.dtype); g = _c(g, img.dtype)
    B, C, H, W = img.shape
    g_img = np.empty_like(img); g_disp = np.empty_like(disp)
    getattr(lib(), "oracle_warp_bwd_" + _suffix(img))(_p(g), _p(img), _p(disp), _p(g_img), _p(g_disp),
                                                     _I(B), _I(C), _I(H), _I(W), _I(_PAD[padding_mode]))
    return g_img, g_disp


def resize_bilinear(img: np.ndarray, size) -> np.ndarray:
    """F.interpolate(img, size=size, mode='bilinear', align_corners=False)."""
    img = _c(img)
    B, C, Hi, Wi = img.shape
    Ho, Wo = int(size[0]), int(size[1])
    out = np.empty((B, C, Ho, Wo), dtype=img.dtype)
    getattr(lib(), "oracle_resize_bilinear_" + _suffix(img))(_p(img), _p(out), _I(B), _I(C), _I(Hi), _I(Wi), _I(Ho), _I(Wo))
    return out


def resize_bilinear_bwd(g_out: np.ndarray, in_size) -> np.ndarray:
    g_out = _c(g_out)
    B, C, Ho, Wo = g_out.shape
    Hi, Wi = int(in_size[0]), int(in_size[1])
    g_in = np.empty((B, C, Hi, Wi), dtype=g_out.dtype)
    getattr(lib(), "oracle_resize_bilinear_bwd_" + _suffix(g_out))(_p(g_out), _p(g_in), _I(B), _I(C), _I(Hi), _I(Wi), _I(Ho), _I(Wo))
    return g_in


def lrwarp_error(imgL: np.ndarray, disp: np.ndarray, imgR: np.ndarray) -> np.ndarray:
    """LRwarp_error(imgL, disp, imgR) -> imgR' - warp(imgL', disp)."""
    imgL = _c(imgL); imgR = _c(imgR, imgL.dtype); disp = _c(disp, imgL.dtype)
    B, C, H0, W0 = imgL.shape
    _, _, H, W = disp.shape
    if W0 <= W:
        assert (H0, W0) == (H, W), "images no wider than disp must already match its size"
    out = np.empty((B, C, H, W), dtype=imgL.dtype)
    scratch = np.empty((3, B, C, H, W), dtype=imgL.dtype)
    fn = getattr(lib(), "oracle_lrwarp_error_" + _suffix(imgL))
    fn.restype = ctypes.c_int
    fn(_p(imgL), _p(disp), _p(imgR), _p(out), _p(scratch), _I(B), _I(C), _I(H0), _I(W0), _I(H), _I(W))
    return out


def photometric_l1(imgL: np.ndarray, disp: np.ndarray, imgR: np.ndarray):
    """LRwarp_error(imgL, disp, imgR).abs().mean() and its gradient w.r.t. disp, composed from the restated
    pieces exactly as autograd composes upstream's: d|e|/de = sign(e), d(err)/d(warped) = -1, then the
    backward of the 'border' warp of the (resized) left image.  Returns (loss as a Python float, g_disp)."""
    err = lrwarp_error(imgL, disp, imgR)
    n = err.size
    loss = float(np.abs(err.astype(np.float64)).sum() / n)
    H, W = disp.shape[-2:]
    Ls = resize_bilinear(_c(imgL), (H, W)) if imgL.shape[-1] > W else _c(imgL)
    g = (-np.sign(err) / n).astype(err.dtype)
    _, g_disp = warp_bwd(g, Ls, _c(disp, err.dtype), "border")
    return loss, g_disp


def haze(clear, src, src_is_disp, baseline, focal, beta, airlight, want=("haze", "trans", "depth")):
    """convert_disp_to_depth -> depth2trans -> recover_haze_images (DeBug/inference.py:57-64, 126-134, 95-112) in
    upstream's operation order and type promotion: the result dtype is float64 as soon as any image or scalar is.
    `src` is the disparity (src_is_disp) or the depth, [B,1,H,W]; scalars are [B] arrays or None when unused.
    Returns a dict with the requested outputs."""
    src = _c(src)
    arrs = [a for a in (clear, src, baseline, focal, beta, airlight) if a is not None]
    cdt = np.float64 if any(np.asarray(a).dtype == np.float64 for a in arrs) else np.float32
    idt = np.float64 if any(np.asarray(a).dtype == np.float64 for a in (clear, src) if a is not None) else np.float32
    src = _c(src, idt)
    B = src.shape[0]
    HW = int(np.prod(src.shape[1:]))
    C = 0
    if clear is not None:
        clear = _c(clear, idt)
        C = clear.shape[1]
        assert src.shape[1] == 1 and clear.shape[0] == B and clear.shape[2:] == src.shape[2:]
    sc = [None if a is None else _c(np.asarray(a).reshape(B), cdt) for a in (baseline, focal, beta, airlight)]
    out = {}
    if "haze" in want:
        out["haze"] = np.empty(clear.shape, dtype=cdt)
    if "trans" in want:
        out["trans"] = np.empty(src.shape, dtype=cdt)
    if "depth" in want:
        out["depth"] = np.empty(src.shape, dtype=cdt)
    null = ctypes.c_void_p(None)
    pp = lambda a: null if a is None else _p(a)  # noqa: E731
    fn = getattr(lib(), "oracle_haze_fwd_" + ("f64" if cdt == np.float64 else "f32"))
    fn(pp(clear), _p(src), _I(1 if src_is_disp else 0), _I(1 if idt == np.float32 else 0), pp(sc[0]), pp(sc[1]), pp(sc[2]),
       pp(sc[3]), pp(out.get("haze")), pp(out.get("trans")), pp(out.get("depth")), _I(B), _I(C), ctypes.c_long(HW))
    return out


def recovered_clean_l1(trans, airlight, haze_img, clean_img):
    """RecoveredCleanImagesLoss.forward (losses/DSSMDLoss.py:87-111) and its autograd gradients:
    returns (loss, g_trans [B,1,h,w], g_airlight [B])."""
    trans = _c(trans)
    haze_img = _c(haze_img, trans.dtype); clean_img = _c(clean_img, trans.dtype)
    B, C, H0, W0 = haze_img.shape
    h, w = trans.shape[-2:]
    if W0 // w != 1:
        haze_img, clean_img = resize_bilinear(haze_img, (h, w)), resize_bilinear(clean_img, (h, w))
    A = _c(np.asarray(airlight).reshape(B), trans.dtype)
    g_t = np.empty_like(trans); g_A = np.empty_like(A)
    fn = getattr(lib(), "oracle_recovered_l1_" + _suffix(trans))
    fn.restype = ctypes.c_double
    loss = fn(_p(trans), _p(A), _p(haze_img), _p(clean_img), _p(g_t), _p(g_A), _I(B), _I(C), ctypes.c_long(h * w))
    return float(loss), g_t, g_A


def epipolar_attention(q, k, v, scale, g=None):
    """Core of STTM_Single.forward (models/BidNet/STTM_Simple.py:43-57) on NCHW tensors: out, and with an upstream
    gradient g also (g_q, g_k, g_v)."""
    q = _c(q); k = _c(k, q.dtype); v = _c(v, q.dtype)
    B, C, H, W = q.shape
    out = np.empty_like(q)
    null = ctypes.c_void_p(None)
    fn = getattr(lib(), "oracle_epipolar_attn_" + _suffix(q))
    sc = ctypes.c_float(scale) if q.dtype == np.float32 else ctypes.c_double(scale)
    if g is None:
        fn(_p(q), _p(k), _p(v), _p(out), null, null, null, null, _I(B), _I(C), _I(H), _I(W), sc)
        return out
    g = _c(g, q.dtype)
    gq, gk, gv = np.empty_like(q), np.empty_like(q), np.empty_like(q)
    fn(_p(q), _p(k), _p(v), _p(out), _p(g), _p(gq), _p(gk), _p(gv), _I(B), _I(C), _I(H), _I(W), sc)
    return out, gq, gk, gv


# ---- bf16 helpers (the CUDA path takes bf16 I/O with fp32 accumulation) ----
def bf16_round(a: np.ndarray) -> np.ndarray:
    """Round float32 values to the nearest bfloat16 (ties to even), kept as float32."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    out = rounded.astype(np.uint32).view(np.float32).copy()
    nan = np.isnan(a)
    out[nan] = np.nan
    return out.reshape(a.shape)
