"""Drop-in for upstream utils/disparity_warper.py (and, re-exported by
dssmd_b200.stereo_op, for models/UtilsNet/stereo_op.py, its byte-identical twin).

disp_warp and LRwarp_error run the CUDA kernels; normalize_coords, meshgrid and
read_pfm are kept importable with upstream's semantics (they are host-side /
pure-tensor helpers that the kernels make unnecessary on the hot path).
"""
from __future__ import annotations

import re

import numpy as np
import torch

from .functional import DispWarpFn, lrwarp_error

__all__ = ["read_pfm", "normalize_coords", "meshgrid", "disp_warp", "LRwarp_error"]


def read_pfm(file):
    """Read a PFM image -> (data flipped to top-down, scale).  upstream :11-46."""
    with open(file, 'rb') as f:
        header = f.readline().rstrip().decode("ascii")
        if header == 'PF':
            color = True
        elif header == 'Pf':
            color = False
        else:
            raise Exception('Not a PFM file.')
        dim_match = re.match(r'^(\d+)\s(\d+)\s$', f.readline().decode("ascii"))
        if not dim_match:
            raise Exception('Malformed PFM header.')
        width, height = map(int, dim_match.groups())
        scale = float(f.readline().decode("ascii").rstrip())
        endian = '<' if scale < 0 else '>'
        scale = abs(scale)
        data = np.fromfile(f, endian + 'f')
    shape = (height, width, 3) if color else (height, width)
    return np.flipud(np.reshape(data, shape)), scale


def normalize_coords(grid):
    """Normalize coordinates of image scale to [-1, 1].  grid: [B, 2, H, W] -> [B, H, W, 2].
    Like upstream (:48-58) this writes the normalised values into `grid` in place."""
    assert grid.size(1) == 2
    h, w = grid.size()[2:]
    grid[:, 0, :, :] = 2 * (grid[:, 0, :, :].clone() / (w - 1)) - 1
    grid[:, 1, :, :] = 2 * (grid[:, 1, :, :].clone() / (h - 1)) - 1
    return grid.permute((0, 2, 3, 1))


def meshgrid(img, homogeneous=False):
    """Pixel-coordinate grid of img [B,_,H,W] -> [B,2,H,W] (x then y); [B,3,H,W] if homogeneous.  upstream :60-80."""
    b, _, h, w = img.size()
    xs = torch.arange(0, w, device=img.device).view(1, 1, w).expand(1, h, w).type_as(img)
    ys = torch.arange(0, h, device=img.device).view(1, h, 1).expand(1, h, w).type_as(img)
    grid = torch.cat((xs, ys), dim=0).unsqueeze(0).expand(b, 2, h, w)
    if homogeneous:
        ones = torch.ones_like(xs).unsqueeze(0).expand(b, 1, h, w)
        grid = torch.cat((grid, ones), dim=1)
        assert grid.size(1) == 3
    return grid


def disp_warp(img, disp, padding_mode='border'):
    """Warping by disparity (same signature as upstream :83-106).

    Args:
        img: [B, C, H, W] CUDA tensor (fp32 / fp64)
        disp: [B, 1, H, W], non-negative (AssertionError otherwise, as upstream)
        padding_mode: 'zeros' or 'border'
    Returns:
        warped_img: [B, C, H, W]
        valid_mask: [B, C, H, W] of 0/1
    """
    return DispWarpFn.apply(img, disp, padding_mode)


def LRwarp_error(imgL, disp, imgR):
    """imgR - warp(imgL, disp), both images first resized to disp's size when wider (upstream :109-115)."""
    return lrwarp_error(imgL, disp, imgR)
