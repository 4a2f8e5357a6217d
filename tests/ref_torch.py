"""PyTorch restatement of the upstream ops, used by the GPU tests as a second,
full-size checker next to the C oracle (the upstream tree itself does not exist
on the GPU box).  Semantics follow upstream models/UtilsNet/submoudles.py:301-311
and utils/disparity_warper.py:48-115; written independently, runs on any device."""
import torch
import torch.nn.functional as F


def corr_ref(left, right, max_disp):
    B, C, H, W = left.shape
    vol = left.new_zeros(B, max_disp, H, W)
    for d in range(min(max_disp, W)):
        vol[:, d, :, d:] = (left[..., d:] * right[..., :W - d]).mean(dim=1)
    return vol


def warp_ref(img, disp, padding_mode="border"):
    B, C, H, W = img.shape
    xs = torch.arange(W, device=img.device, dtype=img.dtype).view(1, 1, W).expand(B, H, W)
    ys = torch.arange(H, device=img.device, dtype=img.dtype).view(1, H, 1).expand(B, H, W)
    gx = 2 * ((xs - disp[:, 0]) / (W - 1)) - 1
    gy = 2 * (ys / (H - 1)) - 1
    grid = torch.stack((gx, gy), dim=-1)
    warped = F.grid_sample(img, grid, mode="bilinear", padding_mode=padding_mode, align_corners=False)
    m = F.grid_sample(torch.ones_like(img), grid, mode="bilinear", padding_mode="zeros", align_corners=False)
    mask = (m >= 0.9999).to(img.dtype)
    return warped, mask


def lrwarp_ref(imgL, disp, imgR):
    size = disp.shape[-2:]
    if imgL.size(-1) > disp.size(-1):
        imgL = F.interpolate(imgL, size=size, mode="bilinear", align_corners=False)
    if imgR.size(-1) > disp.size(-1):
        imgR = F.interpolate(imgR, size=size, mode="bilinear", align_corners=False)
    return imgR - warp_ref(imgL, disp)[0]
