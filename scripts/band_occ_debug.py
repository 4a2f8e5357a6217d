"""Where do the 2- and 3-blocks-per-SM builds of warp_bwd_band2 differ? (development)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["DSSMD_TUNING"] = "1"
import torch
import dssmd_b200
from test_gpu_parity import images, feat
img, disp = images(4, 3, 576, 960, 71)
g = feat((4, 3, 576, 960), 72, False)
outs = {}
for occ, rows in (("2", ""), ("3", ""), ("2", "6"), ("3", "8"), ("2", "8")):
    os.environ["DSSMD_WARP_BWD_BAND_OCC"] = occ
    if rows: os.environ["DSSMD_WARP_BWD_BAND_ROWS"] = rows
    else: os.environ.pop("DSSMD_WARP_BWD_BAND_ROWS", None)
    i2 = img.clone().requires_grad_(True); d2 = disp.clone().requires_grad_(True)
    w, _ = dssmd_b200.disp_warp(i2, d2); w.backward(g)
    outs[(occ, rows)] = (i2.grad.clone(), d2.grad.clone())
ref = outs[("2", "")]
for k, (gi, gd) in outs.items():
    di = (gi - ref[0]).abs(); dd = (gd - ref[1]).abs()
    rows_i = di.amax(dim=(0, 1, 3)).nonzero().flatten().tolist()
    print(k, "g_img max diff", float(di.max()), "n", int((di > 0).sum()), "rows", rows_i[:20], "| g_disp max diff", float(dd.max()), "n", int((dd > 0).sum()),
          "rows", dd.amax(dim=(0, 1, 3)).nonzero().flatten().tolist()[:20])
