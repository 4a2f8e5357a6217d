"""Drop-in proof on the REAL upstream tree (models/DSSMD.py:6-9,129-131, SSMDNet.py:176,
DSSMD_new.py:193, SDNet.py:137): the unmodified upstream models run on the kernels after `dssmd_b200.install()`.

    python scripts/upstream_dropin.py --binding                 # CPU: rebinding + signature equality (tests/test_upstream_dropin.py)
    python scripts/upstream_dropin.py --gpu [--out FILE.json]   # GPU: pyramids / gradients patched vs unpatched, step times,
                                                                #      one reference-style training step (dssmd_traner_new.py:239-360)

The upstream tree is never part of this repository.  It is looked up in $DSSMD_REFERENCE, /root/reference (build
container) or <repo>/_upstream (a git-ignored copy `scripts/stage_upstream.sh` makes for one gpurun call and removes
afterwards).  Prints one JSON document.
"""
from __future__ import annotations

import argparse
import inspect
import json
import os
import sys
import types
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True   # /root/reference is read-only
warnings.filterwarnings("ignore")


def find_upstream():
    for cand in (os.environ.get("DSSMD_REFERENCE"), "/root/reference", os.path.join(ROOT, "_upstream")):
        if cand and os.path.isfile(os.path.join(cand, "models", "UtilsNet", "submoudles.py")):
            return cand
    return None


def import_upstream(path):
    """Upstream on sys.path; empty stubs for the packages it only uses for dataset I/O and plotting."""
    sys.path.insert(0, path)
    for m in ("skimage", "skimage.io", "skimage.transform", "matplotlib", "matplotlib.pyplot"):
        if m not in sys.modules:
            try:
                __import__(m)
            except Exception:
                sys.modules[m] = types.ModuleType(m)


MODEL_MODULES = ("models.SSMDNet", "models.DSSMD", "models.DSSMD_new", "models.SDNet")
HOT_NAMES = ("build_corr", "disp_warp", "CostVolume")


def check_binding():
    """install() on the real modules: every name the models resolve is the replacement, signatures are upstream's."""
    import importlib

    import dssmd_b200
    from dssmd_b200 import cost_volume, disparity_warper, submoudles

    res = {"checks": [], "ok": True}

    def check(name, cond, detail=""):
        res["checks"].append({"name": name, "ok": bool(cond), "detail": detail})
        res["ok"] = res["ok"] and bool(cond)

    mods = {m: importlib.import_module(m) for m in MODEL_MODULES}
    ups = {
        "build_corr": importlib.import_module("models.UtilsNet.submoudles").build_corr,
        "disp_warp": importlib.import_module("utils.disparity_warper").disp_warp,
        "LRwarp_error": importlib.import_module("utils.disparity_warper").LRwarp_error,
        "stereo_op.disp_warp": importlib.import_module("models.UtilsNet.stereo_op").disp_warp,
        "stereo_op.LRwarp_error": importlib.import_module("models.UtilsNet.stereo_op").LRwarp_error,
        "meshgrid": importlib.import_module("utils.disparity_warper").meshgrid,
        "normalize_coords": importlib.import_module("utils.disparity_warper").normalize_coords,
        "CostVolume.__init__": importlib.import_module("models.UtilsNet.cost_volume").CostVolume.__init__,
        "CostVolume.forward": importlib.import_module("models.UtilsNet.cost_volume").CostVolume.forward,
    }
    ours = {
        "build_corr": submoudles.build_corr, "disp_warp": disparity_warper.disp_warp, "LRwarp_error": disparity_warper.LRwarp_error,
        "stereo_op.disp_warp": dssmd_b200.stereo_op.disp_warp, "stereo_op.LRwarp_error": dssmd_b200.stereo_op.LRwarp_error,
        "meshgrid": disparity_warper.meshgrid, "normalize_coords": disparity_warper.normalize_coords,
        "CostVolume.__init__": cost_volume.CostVolume.__init__, "CostVolume.forward": cost_volume.CostVolume.forward,
    }
    for k, up in ups.items():
        su, so = str(inspect.signature(up)), str(inspect.signature(ours[k]))
        check(f"signature:{k}", su == so, f"upstream {su} | ours {so}")

    # the trainers and the loss / attention modules (guarded patches)
    tr = importlib.import_module("trainfiles.dssmd_traner_new")
    dl = importlib.import_module("losses.DSSMDLoss")
    st = importlib.import_module("models.BidNet.STTM_Simple")
    inf = importlib.import_module("DeBug.inference")
    up_loss_cls, up_sttm_cls = dl.RecoveredCleanImagesLoss, st.STTM_Single
    up_loss_fwd, up_sttm_fwd = up_loss_cls.forward, up_sttm_cls.forward
    up_haze = {n: getattr(inf, n) for n in ("convert_disp_to_depth", "depth2trans", "recover_haze_images")}
    for n, f in up_haze.items():
        check(f"signature:haze.{n}", str(inspect.signature(f)) == str(inspect.signature(getattr(dssmd_b200.haze, n))))
    check("signature:RecoveredCleanImagesLoss.forward",
          str(inspect.signature(up_loss_fwd)) == str(inspect.signature(dssmd_b200.RecoveredCleanImagesLoss.forward)))
    check("signature:STTM_Single.forward", str(inspect.signature(up_sttm_fwd)) == str(inspect.signature(dssmd_b200.STTM_Single.forward)))
    check("signature:STTM_Single.__init__", str(inspect.signature(up_sttm_cls.__init__)) == str(inspect.signature(dssmd_b200.STTM_Single.__init__)))

    patched = dssmd_b200.install()
    res["patched"] = [list(p) for p in patched]
    for m, mod in mods.items():
        check(f"rebound:{m}.build_corr", getattr(mod, "build_corr", None) is dssmd_b200.build_corr)
        check(f"rebound:{m}.disp_warp", getattr(mod, "disp_warp", None) is dssmd_b200.disp_warp)
        check(f"rebound:{m}.CostVolume", getattr(mod, "CostVolume", None) is dssmd_b200.CostVolume)
    check("rebound:utils.disparity_warper.LRwarp_error", sys.modules["utils.disparity_warper"].LRwarp_error is dssmd_b200.LRwarp_error)
    check("rebound:stereo_op.LRwarp_error", sys.modules["models.UtilsNet.stereo_op"].LRwarp_error is dssmd_b200.LRwarp_error)
    # guarded: the class objects are still upstream's (isinstance / state_dict / pickling unchanged), forward is patched
    check("guarded:STTM_Single is upstream's class", st.STTM_Single is up_sttm_cls and hasattr(st.STTM_Single.forward, "__wrapped_upstream__"))
    check("guarded:RecoveredCleanImagesLoss is upstream's class", tr.RecoveredCleanImagesLoss is up_loss_cls
          and tr.RecoveredCleanImagesLoss.forward.__wrapped_upstream__ is up_loss_fwd)
    for n, f in up_haze.items():
        g = getattr(tr, n)
        check(f"guarded:trainer.{n}", getattr(g, "__wrapped_upstream__", None) is f and g.__kernel__ is getattr(dssmd_b200.haze, n))

    # a patched model really calls the kernel entry point: on CPU tensors it must raise the no-fallback error
    import torch
    torch.manual_seed(1024)
    net = mods["models.SSMDNet"].SSMDNet(in_channels=3, dehaze_switch=True).eval()
    x = torch.rand(1, 3, 64, 128)
    try:
        with torch.no_grad():
            net(x, x)
        check("patched SSMDNet on CPU raises", False, "ran without raising")
    except RuntimeError as e:
        check("patched SSMDNet on CPU raises", "no CPU fallback" in str(e), str(e)[:120])
    # guarded patches delegate to upstream for inputs outside the kernel limits (here: CPU tensors)
    t = torch.rand(1, 1, 8, 16) * 0.8 + 0.1
    loss_mod = tr.RecoveredCleanImagesLoss()
    a = loss_mod(t, torch.rand(1, 1), torch.rand(1, 3, 8, 16), torch.rand(1, 3, 8, 16))
    check("guarded ASM loss delegates on CPU", torch.is_tensor(a) and a.dim() == 0)
    sm = st.STTM_Single(dim=8, heads=1, dim_head=8, out_dim=8).eval()
    with torch.no_grad():
        o = sm(torch.randn(1, 8, 2, 6), torch.randn(1, 8, 2, 6))
    check("guarded STTM_Single delegates on CPU", tuple(o.shape) == (1, 8, 2, 6))
    d = tr.convert_disp_to_depth(baseline=torch.ones(2), focal_length=torch.ones(2), disp=torch.ones(2, 1, 2, 2))
    check("guarded haze fn delegates on CPU (keyword call)", tuple(d.shape) == (2, 1, 2, 2))

    dssmd_b200.uninstall()
    check("uninstall restores build_corr", mods["models.DSSMD"].build_corr is ups["build_corr"])
    check("uninstall restores forward", st.STTM_Single.forward is up_sttm_fwd and dl.RecoveredCleanImagesLoss.forward is up_loss_fwd)
    with torch.no_grad():
        pyr = net(x, x)[0]
    check("unpatched SSMDNet runs on CPU again", len(pyr) == 4)
    return res


def _rel(a, b):
    import torch
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def _time(fn, iters, torch):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def run_gpu(shapes, iters):
    import importlib

    import torch

    import dssmd_b200
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda", 0)
    res = {"device": torch.cuda.get_device_name(0), "torch": torch.__version__, "models": [], "tolerance": 1e-5}
    specs = [("models.SSMDNet", "SSMDNet", False), ("models.DSSMD", "DSSMMD", False),
             ("models.DSSMD_new", "DSSMMD", True), ("models.SDNet", "SDNet", False)]
    for modname, cls, needs_cam in specs:
        mod = importlib.import_module(modname)
        for (B, H, W) in shapes:
            torch.manual_seed(1024)
            net = getattr(mod, cls)(in_channels=3, dehaze_switch=True).to(dev).train()
            left, right = torch.rand(B, 3, H, W, device=dev), torch.rand(B, 3, H, W, device=dev)
            cam = (torch.full((B, 1, 1, 1), 0.1, device=dev), torch.full((B, 1, 1, 1), 1050.0, device=dev)) if needs_cam else ()
            p0 = next(net.parameters())

            def step(record=None):
                net.zero_grad(set_to_none=True)
                out = net(left, right, *cam)
                pyr = out[0]
                loss = sum(p.float().mean() for p in pyr)
                loss.backward()
                if record is not None:
                    record["pyr"] = [p.detach().clone() for p in pyr]
                    record["grad"] = p0.grad.detach().clone()
                    record["other"] = [o.detach().clone() for o in out[1:] if torch.is_tensor(o)]

            ref, new = {}, {}
            dssmd_b200.uninstall()
            step(ref)
            t_ref = _time(step, iters, torch)
            dssmd_b200.install()
            assert mod.build_corr is dssmd_b200.build_corr
            step(new)
            t_new = _time(step, iters, torch)
            dssmd_b200.uninstall()
            entry = {
                "model": f"{modname}.{cls}", "batch": B, "H": H, "W": W,
                "pyramid_shapes": [list(p.shape) for p in new["pyr"]],
                "pyramid_rel_err": [_rel(a, b) for a, b in zip(new["pyr"], ref["pyr"])],
                "first_conv_grad_rel_err": _rel(new["grad"], ref["grad"]),
                "other_outputs_rel_err": [_rel(a, b) for a, b in zip(new["other"], ref["other"])],
                "step_ms_upstream_ops": round(t_ref, 3), "step_ms_dssmd_b200": round(t_new, 3),
            }
            entry["ok"] = max(entry["pyramid_rel_err"]) <= 1e-5
            res["models"].append(entry)
            del net
            torch.cuda.empty_cache()
    res["train_step"] = train_step(dev, iters)
    res["ok"] = all(m["ok"] for m in res["models"]) and res["train_step"]["ok"]
    return res


def train_step(dev, iters, B=2, H=320, W=640):
    """One step of upstream's live training loop (trainfiles/dssmd_traner_new.py:239-360) on synthetic tensors, with the
    names the trainer module itself holds: haze synthesis -> DSSMD_new.DSSMMD -> the four DSSMDLoss terms -> backward ->
    Adam(amsgrad).  Run unpatched and patched from the same seed; the first step's losses must agree."""
    import importlib

    import torch

    import dssmd_b200
    tr = importlib.import_module("trainfiles.dssmd_traner_new")

    def run(patched):
        (dssmd_b200.install if patched else dssmd_b200.uninstall)()
        torch.manual_seed(1024)
        net = tr.DSSMMD(dehaze_switch=True, in_channels=3).to(dev).train()
        opt = torch.optim.Adam(net.parameters(), 1e-4, betas=(0.9, 0.999), amsgrad=True)
        disp_loss = tr.Disparity_Loss(type='smooth_l1', weights=[0.6, 0.8, 1.0, 1.0])
        trans_loss = tr.Transmission_Loss(weights=[0.6, 0.8, 1.0, 1.0])
        air_loss = tr.Airlight_Loss(type='l1_loss')
        rec_loss = tr.RecoveredCleanImagesLoss(type='normal')
        g = torch.Generator(device=dev).manual_seed(7)
        clear_l = torch.nn.functional.avg_pool2d(torch.rand(B, 3, H, W, device=dev, generator=g), 9, 1, 4)
        clear_r = torch.nn.functional.avg_pool2d(torch.rand(B, 3, H, W, device=dev, generator=g), 9, 1, 4)
        base = torch.rand(B, 1, H // 32 + 1, W // 32 + 1, device=dev, generator=g)
        d_l = 0.5 + 127.5 * torch.nn.functional.interpolate(base, size=(H, W), mode="bilinear", align_corners=True)
        d_r = d_l.flip(-1).contiguous()
        # the loaders hand the per-sample scalars over as float64
        focal = torch.full((B,), 1050.0, dtype=torch.float64, device=dev)
        baseline = torch.full((B,), 0.1, dtype=torch.float64, device=dev)
        beta = torch.full((B,), 0.1, dtype=torch.float64, device=dev)
        airlight = torch.full((B,), 0.9, dtype=torch.float64, device=dev)
        log = {}

        def step():
            left_depth = tr.convert_disp_to_depth(baseline=baseline, focal_length=focal, disp=d_l)
            right_depth = tr.convert_disp_to_depth(baseline=baseline, focal_length=focal, disp=d_r)
            left_trans = tr.depth2trans(left_depth, beta=beta)
            haze_l = tr.recover_haze_images(clean_images=clear_l, beta=beta, A=airlight, depth=left_depth).float()
            haze_r = tr.recover_haze_images(clean_images=clear_r, beta=beta, A=airlight, depth=right_depth).float()
            opt.zero_grad()
            pyr, tpyr, pred_air = net(haze_l, haze_r, baseline.float().view(B, 1, 1, 1), focal.float().view(B, 1, 1, 1))
            l_d = disp_loss(pyr, d_l.float())
            l_t = trans_loss(tpyr, left_trans.float())
            l_a = air_loss(pred_air, airlight.float().view(B, 1))
            l_r = 0
            for i, pt in enumerate(tpyr):
                l_r = l_r + rec_loss(pt, pred_air, haze_l, clear_l) * [0.6, 0.8, 1.0, 1.0][i]
            total = l_d * 2.0 + l_t + l_a + l_r
            total.backward()
            opt.step()
            log.setdefault("first", [float(total), float(l_d), float(l_t), float(l_a), float(l_r)])
            log["last"] = float(total)

        step()
        ms = _time(step, iters, torch)
        return log, ms

    ref, t_ref = run(False)
    new, t_new = run(True)
    dssmd_b200.uninstall()
    rel = [abs(a - b) / max(abs(b), 1e-30) for a, b in zip(new["first"], ref["first"])]
    return {"loop": "trainfiles/dssmd_traner_new.py:239-360 on synthetic tensors", "model": "models.DSSMD_new.DSSMMD", "batch": B, "H": H, "W": W,
            "first_step_losses_upstream_ops": ref["first"], "first_step_losses_dssmd_b200": new["first"],
            "first_step_rel_err": rel, "step_ms_upstream_ops": round(t_ref, 3), "step_ms_dssmd_b200": round(t_new, 3),
            "ok": max(rel) <= 1e-4}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--binding", action="store_true")
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    up = find_upstream()
    if up is None:
        print(json.dumps({"skipped": "no upstream tree (DSSMD_REFERENCE, /root/reference, _upstream)"}))
        return 0
    import_upstream(up)
    res = {"upstream": up}
    if a.binding:
        res["binding"] = check_binding()
    if a.gpu:
        res["gpu"] = run_gpu([(2, 320, 640), (2, 576, 960)], a.iters)   # B >= 2: the airlight head has BatchNorm on a 1x1 map
    res["ok"] = all(v.get("ok", True) for v in res.values() if isinstance(v, dict))
    doc = json.dumps(res, indent=1)
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        with open(a.out, "w") as f:
            f.write(doc + "\n")
    print(doc)
    return 0 if res["ok"] else 1


if __name__ == "__main__":
    sys.exit(main())
