"""Drop-in for the correlation entry point of upstream models/UtilsNet/submoudles.py.

Only `build_corr` (submoudles.py:301-311) is on the hot path; the conv/deconv
helpers and ResBlocks of that file are dense cuDNN convolutions and stay
upstream's (dssmd_b200.install() rebinds just this one name in place).
"""
from __future__ import annotations

from .functional import CorrelationCatFn, CorrelationFn

__all__ = ["build_corr", "build_corr_cat"]


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


def build_corr_cat(head, img_left, img_right, max_disp=40):
    """`torch.cat((head, build_corr(img_left, img_right, max_disp)), dim=1)` in one pass over the volume
    (upstream models/DSSMD.py:129-131 with head = conv_redir(conv3_l)).

    The correlation kernel writes its `max_disp` channels directly behind the `head` channels of the concat buffer, and
    the backward reads them out of the concat gradient in place: the volume is never written on its own, re-read and
    re-written by `cat`, nor sliced into a contiguous copy on the way back.  Results are bit-identical to the two-step
    form.  Returns a contiguous [B, head_channels + max_disp, H, W] tensor.
    """
    return CorrelationCatFn.apply(head, img_left, img_right, max_disp)
