"""Drop-in for the correlation entry point of upstream models/UtilsNet/submoudles.py.

Only `build_corr` (submoudles.py:301-311) is on the hot path; the conv/deconv
helpers and ResBlocks of that file are dense cuDNN convolutions and stay
upstream's (dssmd_b200.install() rebinds just this one name in place).
"""
from __future__ import annotations

from .functional import CorrelationFn

__all__ = ["build_corr"]


def build_corr(img_left, img_right, max_disp=40):
    """1-D horizontal correlation cost volume (same signature as upstream).

    Args:
        img_left, img_right: [B, C, H, W] CUDA tensors (fp32 / bf16 / fp16 / fp64), any strides.
        max_disp: number of disparity channels D; d = 0 .. max_disp-1.
    Returns:
        contiguous [B, max_disp, H, W] tensor, same dtype/device:
        out[:, d, :, x] = mean_c(left[:, c, :, x] * right[:, c, :, x-d]) for x >= d, else 0.
    """
    return CorrelationFn.apply(img_left, img_right, max_disp)
