#!/usr/bin/env python
"""Per-kernel timing sweeps over the tuning environment overrides (development tool).

    python scripts/kbench.py corr_fwd           # sweep a few DSSMD_CORR_FWD_* settings on every workload
Times each C-ABI launch with CUDA events inside a CUDA graph that rotates through
input sets larger than L2; prints us, GB/s (algorithmic bytes) and fraction of the measured HBM peak.
"""
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

os.environ.setdefault("DSSMD_TUNING", "1")   # the sweep / debug variables are honoured only with it
import torch  # noqa: E402

import bench  # noqa: E402
from dssmd_b200 import _lib  # noqa: E402


def time_kernel(calls, name, nsets, rounds=20):
    def one_round():
        for i in range(nsets):
            calls[i][name]()
    one_round()
    torch.cuda.synchronize()
    g = bench.capture(torch, one_round)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rounds):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (rounds * nsets)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "corr_fwd"
    workloads = sys.argv[2].split(",") if len(sys.argv) > 2 and sys.argv[2] != "all" else ["sceneflow_576x960", "corr_micro", "kitti_384x1280", "train_320x640"]
    lib = _lib.lib()
    dev = torch.device("cuda", 0)
    peak, _ = bench.measured_peak()
    sweeps = {
        "corr_fwd": [("DSSMD_CORR_FWD_XT", ["", "1", "2", "3", "4"]), ("DSSMD_CORR_FWD_CS", ["1", "2", "4"]),
                     ("DSSMD_CORR_FWD_CK", ["8", "16"])],
        "corr_bwd": [("DSSMD_CORR_BWD_FUSED", ["1", "0"])],
        "corr_bwd_gL": [("DSSMD_CORR_BWD_CG", ["", "1", "2", "4", "8"]), ("DSSMD_CORR_BWD_ROWS", ["", "1", "2", "4"]),
                        ("DSSMD_CORR_BWD_SMEM_KB", ["", "40", "100"])],
        "corr_bwd_gR": [("DSSMD_CORR_BWD_CG", ["", "1", "2", "4", "8"]), ("DSSMD_CORR_BWD_ROWS", ["", "1", "2", "4"]),
                        ("DSSMD_CORR_BWD_SMEM_KB", ["", "40", "100"])],
        "warp_fwd": [("DSSMD_WARP_FWD_X", [""])],
        "warp_bwd": [("DSSMD_WARP_BWD_FLAGS", ["", "0", "1", "2", "3"]), ("DSSMD_WARP_BWD_PXT", ["2", "4"])],
    }[which]
    if len(sys.argv) > 3:   # custom sweep: VAR=v1,v2,...  ('-' = unset)
        var, vals = sys.argv[3].split("=")
        sweeps = [(var, ["" if v == "-" else v for v in vals.split(",")])]
    for wname in workloads:
        w = bench.WORKLOADS[wname]
        nsets = 4
        sets = [bench.make_set(torch, w, dev, 1024 + i) for i in range(nsets)]
        calls = [bench.kernel_calls(torch, lib, _lib, w, s, split_bwd=True) for s in sets]
        ab = bench.algorithmic_bytes(w)[which]
        print(f"== {wname} {which}: algorithmic {ab/1e6:.1f} MB, floor {ab/peak/1e3:.1f} us @ {peak:.0f} GB/s", flush=True)
        # one-at-a-time sweep around the defaults
        for var, values in sweeps:
            for v in values:
                for k, _ in sweeps:
                    os.environ.pop(k, None)
                if v:
                    os.environ[var] = v
                try:
                    us = time_kernel(calls, which, nsets)
                    print(f"   {var}={v or 'default':8s} {us:9.2f} us  {ab/us/1e3:8.1f} GB/s  frac {ab/us/1e3/peak:.3f}", flush=True)
                except Exception as e:  # noqa: BLE001
                    print(f"   {var}={v}: {type(e).__name__}: {e}", flush=True)
        for k, _ in sweeps:
            os.environ.pop(k, None)
        del sets, calls
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
