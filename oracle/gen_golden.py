"""Generate tests/golden/*.npz by running the UPSTREAM functions themselves.

Run in the build container only (needs /root/reference; the GPU box has no
copy):   python oracle/gen_golden.py

The upstream repo has no tests or fixtures for this path, so the committed
fixtures are outputs of its own code on seeded inputs (CPU, torch of this
image).  They pin the oracle (tests/test_oracle_golden.py) and, on the GPU, the
CUDA kernels (tests/test_gpu_parity.py).  Seeds follow scripts/train.sh:33
(1024) plus a per-case offset.
"""
import os
import sys
import warnings

import numpy as np
import torch

REF = os.environ.get("DSSMD_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

from models.UtilsNet.submoudles import build_corr  # noqa: E402
from models.UtilsNet.stereo_op import disp_warp as disp_warp_stereo_op  # noqa: E402
from utils.disparity_warper import LRwarp_error, disp_warp  # noqa: E402

# DeBug/inference.py imports skimage / matplotlib (absent here) for dataset I/O and plotting only: empty stubs let
# the module import unmodified
import types  # noqa: E402
for _m in ("skimage", "skimage.io", "skimage.transform", "matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(_m, types.ModuleType(_m))
from losses.DSSMDLoss import RecoveredCleanImagesLoss, RecoveredCleanImagesLossV2  # noqa: E402
from models.BidNet.STTM_Simple import STTM_Single  # noqa: E402
from DeBug.inference import convert_disp_to_depth, depth2trans, recover_haze_images  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
SEED = 1024


def _np(t):
    return t.detach().to(torch.float32).contiguous().numpy()


def corr_cases():
    # name, B, C, H, W, D, relu, dtype
    return [
        ("tiny", 2, 8, 5, 12, 4, True, torch.float32),
        ("d_gt_w", 1, 16, 3, 6, 9, False, torch.float32),
        ("model_like", 1, 128, 4, 24, 24, True, torch.float32),
        ("ragged", 2, 20, 3, 37, 11, False, torch.float32),
        ("d41", 1, 64, 2, 56, 41, True, torch.float32),
        ("bf16", 1, 64, 3, 24, 8, True, torch.bfloat16),
    ]


def gen_corr():
    out = {}
    for i, (name, B, C, H, W, D, relu, dt) in enumerate(corr_cases()):
        torch.manual_seed(SEED + i)
        L = torch.randn(B, C, H, W)
        R = torch.randn(B, C, H, W)
        if relu:
            L, R = L.relu(), R.relu()
        g = torch.randn(B, D, H, W)
        L, R, g = L.to(dt), R.to(dt), g.to(dt)
        Lr, Rr = L.clone().requires_grad_(True), R.clone().requires_grad_(True)
        vol = build_corr(Lr, Rr, D)
        vol.backward(g)
        assert vol.is_contiguous() and vol.shape == (B, D, H, W) and vol.dtype == dt
        out.update({f"{name}.L": _np(L), f"{name}.R": _np(R), f"{name}.g": _np(g),
                    f"{name}.D": np.int32(D), f"{name}.out": _np(vol),
                    f"{name}.gL": _np(Lr.grad), f"{name}.gR": _np(Rr.grad),
                    f"{name}.bf16": np.bool_(dt == torch.bfloat16)})
    np.savez_compressed(os.path.join(OUT, "corr.npz"), **out)
    print("corr.npz:", len(corr_cases()), "cases")


def warp_cases():
    # name, B, C, H, W, max disparity, padding_mode, smooth image?
    return [
        ("border_small", 2, 3, 9, 16, 6.0, "border", False),
        ("zeros_small", 2, 3, 9, 16, 6.0, "zeros", False),
        ("border_wide", 1, 3, 12, 40, 17.5, "border", True),
        ("zeros_wide", 1, 3, 12, 40, 17.5, "zeros", True),
        ("c1", 1, 1, 5, 7, 3.0, "border", False),
        ("c5_zero_disp", 1, 5, 6, 10, 0.0, "border", False),
        ("big_disp", 1, 3, 4, 12, 30.0, "border", False),
    ]


def gen_warp():
    out = {}
    for i, (name, B, C, H, W, dmax, pad, smooth) in enumerate(warp_cases()):
        torch.manual_seed(SEED + 100 + i)
        img = torch.rand(B, C, H, W)
        if smooth:
            img = torch.nn.functional.avg_pool2d(img, 3, 1, 1)
        disp = torch.rand(B, 1, H, W) * dmax
        g = torch.randn(B, C, H, W)
        ir, dr = img.clone().requires_grad_(True), disp.clone().requires_grad_(True)
        warped, mask = disp_warp(ir, dr, padding_mode=pad)
        w2, m2 = disp_warp_stereo_op(img, disp, padding_mode=pad)
        assert torch.equal(w2, warped.detach()) and torch.equal(m2, mask.detach())
        warped.backward(g)
        out.update({f"{name}.img": _np(img), f"{name}.disp": _np(disp), f"{name}.g": _np(g),
                    f"{name}.pad": np.str_(pad), f"{name}.warped": _np(warped), f"{name}.mask": _np(mask),
                    f"{name}.g_img": _np(ir.grad), f"{name}.g_disp": _np(dr.grad)})
    np.savez_compressed(os.path.join(OUT, "warp.npz"), **out)
    print("warp.npz:", len(warp_cases()), "cases")


def lrwarp_cases():
    # name, B, C, H0, W0, H, W
    return [
        ("half", 1, 3, 12, 20, 6, 10),
        ("same", 2, 3, 8, 14, 8, 14),
        ("quarter", 1, 3, 16, 32, 4, 8),
        ("eighth", 1, 3, 16, 48, 2, 6),
        ("non_integer", 1, 3, 10, 21, 7, 9),
    ]


def gen_lrwarp():
    out = {}
    for i, (name, B, C, H0, W0, H, W) in enumerate(lrwarp_cases()):
        torch.manual_seed(SEED + 200 + i)
        imgL, imgR = torch.rand(B, C, H0, W0), torch.rand(B, C, H0, W0)
        disp = torch.rand(B, 1, H, W) * (W / 4.0)
        g = torch.randn(B, C, H, W)
        lr, rr, dr = (t.clone().requires_grad_(True) for t in (imgL, imgR, disp))
        err = LRwarp_error(lr, dr, rr)
        err.backward(g)
        out.update({f"{name}.imgL": _np(imgL), f"{name}.imgR": _np(imgR), f"{name}.disp": _np(disp),
                    f"{name}.g": _np(g), f"{name}.err": _np(err), f"{name}.g_imgL": _np(lr.grad),
                    f"{name}.g_imgR": _np(rr.grad), f"{name}.g_disp": _np(dr.grad)})
    np.savez_compressed(os.path.join(OUT, "lrwarp.npz"), **out)
    print("lrwarp.npz:", len(lrwarp_cases()), "cases")


def photo_cases():
    # name, B, C, H0, W0, [(H, W) per scale], weights
    return [
        ("pyramid4", 2, 3, 16, 32, [(2, 4), (4, 8), (8, 16), (16, 32)], (0.6, 0.8, 1.0, 1.0)),
        ("single", 1, 3, 12, 24, [(12, 24)], (1.0,)),
        ("c1_half", 1, 1, 8, 40, [(4, 20)], (0.5,)),
    ]


def gen_photo():
    """The photometric loss has no upstream caller; its definition in upstream terms is
    sum_s w_s * LRwarp_error(imgL, disp_s, imgR).abs().mean()  (dssmd_b200/losses.py)."""
    out = {}
    for i, (name, B, C, H0, W0, sizes, weights) in enumerate(photo_cases()):
        torch.manual_seed(SEED + 300 + i)
        imgL, imgR = torch.rand(B, C, H0, W0), torch.rand(B, C, H0, W0)
        disps = [(torch.rand(B, 1, h, w) * (w / 4.0)).requires_grad_(True) for h, w in sizes]
        terms = [LRwarp_error(imgL, d, imgR).abs().mean() for d in disps]
        total = sum(wt * t for wt, t in zip(weights, terms))
        total.backward()
        out.update({f"{name}.imgL": _np(imgL), f"{name}.imgR": _np(imgR), f"{name}.n": np.int32(len(sizes)),
                    f"{name}.weights": np.asarray(weights, dtype=np.float64), f"{name}.total": np.float64(total.item()),
                    f"{name}.terms": np.asarray([t.item() for t in terms], dtype=np.float64)})
        for k, d in enumerate(disps):
            out[f"{name}.disp{k}"] = _np(d)
            out[f"{name}.g_disp{k}"] = _np(d.grad)
    np.savez_compressed(os.path.join(OUT, "photo.npz"), **out)
    print("photo.npz:", len(photo_cases()), "cases")


def haze_cases():
    # name, B, H, W, image dtype, scalar dtype, (baseline, focal), beta range, airlight range
    return [
        ("sceneflow_f64_scalars", 2, 12, 20, torch.float32, torch.float64, (0.1, 1050.0), (0.02, 0.30), (0.8, 1.0)),
        ("kitti_f64_scalars", 2, 6, 31, torch.float32, torch.float64, (0.54, 721.277), (0.008, 0.06), (0.85, 1.0)),
        ("all_f32", 3, 5, 16, torch.float32, torch.float32, (0.1, 1050.0), (0.02, 0.30), (0.8, 1.0)),
        ("all_f64", 1, 4, 9, torch.float64, torch.float64, (0.54, 721.277), (0.008, 0.06), (0.85, 1.0)),
    ]


def gen_haze():
    """The per-batch chain of trainfiles/dssmd_traner_new.py:250-256 on the dtypes the data loaders hand over."""
    out = {}
    for i, (name, B, H, W, idt, sdt, (bl, fl), brange, arange) in enumerate(haze_cases()):
        torch.manual_seed(SEED + 400 + i)
        clear = torch.rand(B, 3, H, W).to(idt)
        disp = (0.5 + torch.rand(B, 1, H, W) * 191.5 * W / 960.0 * 8).to(idt)
        baseline = torch.full((B,), bl, dtype=sdt)
        focal = torch.full((B,), fl, dtype=sdt)
        beta = (brange[0] + torch.rand(B, dtype=torch.float64) * (brange[1] - brange[0])).to(sdt)
        A = (arange[0] + torch.rand(B, dtype=torch.float64) * (arange[1] - arange[0])).to(sdt)
        depth = convert_disp_to_depth(baseline=baseline, focal_length=focal, disp=disp)
        trans = depth2trans(depth, beta=beta)
        haze = recover_haze_images(clean_images=clear, beta=beta, A=A, depth=depth)
        keep = lambda t: t.detach().contiguous().numpy()  # noqa: E731  (dtype preserved: float64 results stay float64)
        out.update({f"{name}.clear": keep(clear), f"{name}.disp": keep(disp), f"{name}.baseline": keep(baseline),
                    f"{name}.focal": keep(focal), f"{name}.beta": keep(beta), f"{name}.airlight": keep(A),
                    f"{name}.depth": keep(depth), f"{name}.trans": keep(trans), f"{name}.haze": keep(haze)})
    np.savez_compressed(os.path.join(OUT, "haze.npz"), **out)
    print("haze.npz:", len(haze_cases()), "cases")


def asm_cases():
    # name, B, C, H0, W0, scale, V2?
    return [
        ("full", 2, 3, 8, 16, 1, False),
        ("half", 2, 3, 8, 16, 2, False),
        ("quarter_v2", 1, 3, 16, 32, 4, True),
        ("eighth", 2, 3, 16, 32, 8, False),
        ("ragged_full", 1, 3, 5, 7, 1, False),
    ]


def gen_asm():
    """RecoveredCleanImagesLoss (losses/DSSMDLoss.py:81-145) on a synthesised hazy image, a perturbed transmission map and a
    perturbed airlight, so that the clamp is active on both sides for part of the pixels."""
    out = {}
    for i, (name, B, C, H0, W0, scale, v2) in enumerate(asm_cases()):
        torch.manual_seed(SEED + 500 + i)
        h, w = H0 // scale, W0 // scale
        clean = torch.rand(B, C, H0, W0)
        t_true = 0.15 + 0.8 * torch.rand(B, 1, H0, W0)
        A_true = 0.8 + 0.2 * torch.rand(B, 1)
        haze = clean * t_true + A_true.view(B, 1, 1, 1) * (1 - t_true)
        t_small = torch.nn.functional.interpolate(t_true, size=(h, w), mode="bilinear", align_corners=False) if scale > 1 else t_true
        trans = (t_small * (0.7 + 0.6 * torch.rand(B, 1, h, w))).clamp(0.05, 1.0).requires_grad_(True)
        A = (A_true + 0.2 * (torch.rand(B, 1) - 0.5)).requires_grad_(True)
        mod = (RecoveredCleanImagesLossV2 if v2 else RecoveredCleanImagesLoss)(type="normal")
        loss = mod(trans, A.view(B, 1, 1, 1) if v2 else A, haze, clean)
        loss.backward()
        out.update({f"{name}.trans": _np(trans), f"{name}.airlight": _np(A), f"{name}.haze": _np(haze), f"{name}.clean": _np(clean),
                    f"{name}.v2": np.bool_(v2), f"{name}.loss": np.float64(loss.item()),
                    f"{name}.g_trans": _np(trans.grad), f"{name}.g_airlight": _np(A.grad)})
    np.savez_compressed(os.path.join(OUT, "asm.npz"), **out)
    print("asm.npz:", len(asm_cases()), "cases")


def attn_cases():
    # name, B, dim, heads, dim_head, H, W
    return [
        ("bidnet_like", 1, 32, 1, 32, 2, 72),        # BidNet's configuration (models/BidNet/BidNet.py:115), W not a multiple of 64
        ("two_tiles", 1, 16, 1, 16, 2, 128),
        ("short_row", 2, 16, 2, 8, 2, 20),
        ("c64", 1, 64, 2, 32, 2, 40),                # to_out takes 2 * inner_dim channels: dim == heads * dim_head
        ("c128", 1, 128, 4, 32, 1, 68),               # upstream's own example shape class (STTM_Simple.py:73)
    ]


def gen_attn():
    """STTM_Single (models/BidNet/STTM_Simple.py) run unmodified: the fixtures hold the outputs of its three 1x1 convolutions
    (forward hooks), the attention output it concatenates to the left features (input of to_out), the gradient arriving
    there and the gradients autograd returns to q, k, v -- plus the module's weights, inputs and final output for the
    drop-in module test."""
    out = {}
    for i, (name, B, dim, heads, dh, H, W) in enumerate(attn_cases()):
        torch.manual_seed(SEED + 600 + i)
        mod = STTM_Single(dim=dim, heads=heads, dim_head=dh, out_dim=dim).train()
        left, right = 2 * torch.randn(B, dim, H, W), 2 * torch.randn(B, dim, H, W)
        kept = {}

        def keep(key):
            def hook(_m, _inp, o):
                o.retain_grad()
                kept[key] = o
            return hook
        hs = [mod.q.register_forward_hook(keep("q")), mod.k.register_forward_hook(keep("k")), mod.v.register_forward_hook(keep("v"))]

        def pre(_m, inp):
            inp[0].retain_grad()
            kept["cat"] = inp[0]
        hs.append(mod.to_out[0].register_forward_pre_hook(pre))
        left_r, right_r = left.clone().requires_grad_(True), right.clone().requires_grad_(True)
        y = mod(left_r, right_r)
        gy = torch.randn_like(y)
        y.backward(gy)
        for h in hs:
            h.remove()
        inner = heads * dh
        out.update({f"{name}.q": _np(kept["q"]), f"{name}.k": _np(kept["k"]), f"{name}.v": _np(kept["v"]),
                    f"{name}.scale": np.float64(mod.scale), f"{name}.out": _np(kept["cat"][:, dim:]),
                    f"{name}.g_out": _np(kept["cat"].grad[:, dim:]), f"{name}.g_q": _np(kept["q"].grad),
                    f"{name}.g_k": _np(kept["k"].grad), f"{name}.g_v": _np(kept["v"].grad),
                    f"{name}.cfg": np.asarray([dim, heads, dh, dim], dtype=np.int32)})
        assert kept["q"].shape == (B, inner, H, W)
        if name in ("bidnet_like", "short_row"):   # the whole module, for the drop-in test (kept to two cases: fixture size)
            out.update({f"{name}.left": _np(left), f"{name}.right": _np(right), f"{name}.y": _np(y), f"{name}.g_y": _np(gy),
                        f"{name}.g_left": _np(left_r.grad), f"{name}.g_right": _np(right_r.grad)})
            for k_, t in mod.state_dict().items():
                out[f"{name}.w:{k_}"] = t.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, "attn.npz"), **out)
    print("attn.npz:", len(attn_cases()), "cases")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)  # fixed reduction order for the fixtures
    print("torch", torch.__version__, "reference at", REF)
    gen_corr()
    gen_warp()
    gen_lrwarp()
    gen_photo()
    gen_haze()
    gen_asm()
    gen_attn()
