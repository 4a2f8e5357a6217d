#!/usr/bin/env python
"""What a kernel boundary costs inside the step: every ordered pair (A, B) of the four bench calls replayed as A B A B ... in a
CUDA graph (inputs rotating over > L2) against the sum of the two calls timed alone (A A A ..., B B B ...).  Development tool."""
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("DSSMD_TUNING", "1")
import torch  # noqa: E402

import bench  # noqa: E402
from dssmd_b200 import _lib  # noqa: E402


def time_seq(calls, names, nsets, rounds=20):
    def one_round():
        for i in range(nsets):
            for n in names:
                calls[i][n]()
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
    wname = sys.argv[1] if len(sys.argv) > 1 else "sceneflow_576x960"
    lib = _lib.lib()
    dev = torch.device("cuda", 0)
    w = bench.WORKLOADS[wname]
    nsets = 4
    sets = [bench.make_set(torch, w, dev, 1024 + i) for i in range(nsets)]
    calls = [bench.kernel_calls(torch, lib, _lib, w, s, split_bwd=False) for s in sets]
    names = ["corr_fwd", "corr_bwd", "warp_fwd", "warp_bwd"]
    alone = {n: time_seq(calls, [n], nsets) for n in names}
    print("alone:", {n: round(t, 2) for n, t in alone.items()}, "sum", round(sum(alone.values()), 2))
    for a, b in itertools.permutations(names, 2):
        t = time_seq(calls, [a, b], nsets)
        print(f"{a:9s} -> {b:9s}: {t:6.2f} us  (alone {alone[a] + alone[b]:6.2f}, boundary cost {t - alone[a] - alone[b]:+5.2f} for the two boundaries)")
    for order in (names, ["corr_fwd", "warp_fwd", "warp_bwd", "corr_bwd"], ["warp_fwd", "corr_fwd", "corr_bwd", "warp_bwd"]):
        print("step", order, round(time_seq(calls, order, nsets), 2))


if __name__ == "__main__":
    main()
