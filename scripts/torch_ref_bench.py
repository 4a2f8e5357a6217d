#!/usr/bin/env python
"""Times the torch restatement of the upstream ops ON THE GPU ("what users run today"):
the same op sequence upstream's build_corr / disp_warp issue, with autograd backward.
Context numbers for the bench, not part of the product path."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import bench  # noqa: E402
import ref_torch  # noqa: E402


def timeit(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def main():
    dev = torch.device("cuda", 0)
    out = {}
    for wname, w in bench.WORKLOADS.items():
        s = bench.make_set(torch, w, dev, 1024)
        L, R = s["L"].clone().requires_grad_(True), s["R"].clone().requires_grad_(True)
        img, disp = s["img"].clone().requires_grad_(True), s["disp"].clone().requires_grad_(True)

        def corr_f():
            with torch.no_grad():
                ref_torch.corr_ref(L, R, w["D"])

        def corr_fb():
            L.grad = R.grad = None
            ref_torch.corr_ref(L, R, w["D"]).backward(s["g_vol"])

        def warp_f():
            with torch.no_grad():
                ref_torch.warp_ref(img, disp)

        def warp_fb():
            img.grad = disp.grad = None
            ref_torch.warp_ref(img, disp)[0].backward(s["g_img"])

        r = {"corr_fwd_us": timeit(corr_f), "corr_fwd_bwd_us": timeit(corr_fb),
             "warp_fwd_us": timeit(warp_f), "warp_fwd_bwd_us": timeit(warp_fb)}
        r["step_us"] = r["corr_fwd_bwd_us"] + r["warp_fwd_bwd_us"]
        r["pairs_per_s"] = w["B"] / (r["step_us"] * 1e-6)
        out[wname] = {k: round(v, 1) for k, v in r.items()}
        print(wname, out[wname], flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
