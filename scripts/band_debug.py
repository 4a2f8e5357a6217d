"""Development tool: compare the one-pass warp backward (variant 6) with the two-kernel path (variant 5)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dssmd_b200 as ops

def images(B, C, H, W, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    img = torch.rand((B, C, H, W), generator=g, device="cuda")
    base = torch.rand((B, 1, H // 8 + 1, W // 8 + 1), generator=g, device="cuda")
    disp = torch.nn.functional.interpolate(base, size=(H, W), mode="bilinear", align_corners=True)
    return img, (0.5 + disp * (192.0 * W / 960.0 - 0.5)).contiguous()

for shape in [(1, 4, 5, 256), (1, 3, 5, 256), (1, 2, 5, 256), (1, 1, 5, 256), (1, 4, 9, 96), (1, 4, 5, 512)]:
    for pad in ("border", "zeros"):
        B, C, H, W = shape
        img, disp = images(B, C, H, W, 5)
        g = torch.randn((B, C, H, W), device="cuda")
        res = []
        for v in ("6", "5"):
            os.environ["DSSMD_WARP_BWD_VARIANT"] = v
            i2 = img.clone().requires_grad_(True); d2 = disp.clone().requires_grad_(True)
            w, _ = ops.disp_warp(i2, d2, padding_mode=pad)
            w.backward(g)
            res.append((i2.grad.clone(), d2.grad.clone()))
        di = (res[0][0] - res[1][0]).abs()
        idx = (di > 1e-4).nonzero()
        print(shape, pad, "g_img maxdiff", float(di.max()), "n bad", idx.shape[0], "first", idx[:6].tolist(),
              "g_disp maxdiff", float((res[0][1] - res[1][1]).abs().max()), flush=True)
        if idx.shape[0]:
            b, c, y, x = idx[0].tolist()
            print("   got", float(res[0][0][b, c, y, x]), "ref", float(res[1][0][b, c, y, x]))
