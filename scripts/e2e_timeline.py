"""Event timeline of bench.py's end-to-end step (development tool): where do the streams wait?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench, dssmd_b200
w = bench.WORKLOADS["sceneflow_576x960"]; dev = torch.device("cuda", 0)
s0 = bench.make_set(torch, w, dev, 1)
h = {k: s0[k].cpu().pin_memory() for k in ("L", "R", "g_vol", "img", "disp", "g_img")}
out_keys = ("vol", "gL", "gR", "warped", "mask", "d_img", "d_disp")
hout = [{k: torch.empty_like(s0[k], device="cpu").pin_memory() for k in out_keys} for _ in range(2)]
s_in, s_cmp, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
marks = []
def mark(name, stream):
    ev = torch.cuda.Event(enable_timing=True); ev.record(stream); marks.append((name, ev, time.perf_counter()))
def upload(keys):
    with torch.cuda.stream(s_in):
        d = {k: h[k].to(dev, non_blocking=True) for k in keys}
        for t in d.values(): t.record_stream(s_cmp)
        ev = torch.cuda.Event(); ev.record(s_in)
    return d, ev
def download(out, items, after):
    with torch.cuda.stream(s_out):
        s_out.wait_event(after)
        for k, t in items:
            t.record_stream(s_out)
            out[k].copy_(t, non_blocking=True)
done = [None, None]
def step(i):
    slot = i % 2
    if done[slot] is not None: done[slot].synchronize()
    out = hout[slot]
    mark(f"{i} h2d start", s_in)
    dw, ev_w_in = upload(("img", "disp", "g_img"))
    mark(f"{i} W h2d end", s_in)
    dc, ev_c_in = upload(("L", "R", "g_vol"))
    mark(f"{i} C h2d end", s_in)
    with torch.cuda.stream(s_cmp):
        s_cmp.wait_event(ev_w_in)
        img, disp = dw["img"].requires_grad_(True), dw["disp"].requires_grad_(True)
        warped, mask = dssmd_b200.disp_warp(img, disp)
        warped.backward(dw["g_img"])
        ev = torch.cuda.Event(); ev.record(s_cmp)
        mark(f"{i} W compute end", s_cmp)
    download(out, (("warped", warped.detach()), ("mask", mask), ("d_img", img.grad), ("d_disp", disp.grad)), ev)
    mark(f"{i} W d2h end", s_out)
    with torch.cuda.stream(s_cmp):
        s_cmp.wait_event(ev_c_in)
        L, R = dc["L"].requires_grad_(True), dc["R"].requires_grad_(True)
        vol = dssmd_b200.build_corr(L, R, w["D"])
        vol.backward(dc["g_vol"])
        ev = torch.cuda.Event(); ev.record(s_cmp)
        mark(f"{i} C compute end", s_cmp)
    download(out, (("vol", vol.detach()), ("gL", L.grad), ("gR", R.grad)), ev)
    mark(f"{i} C d2h end", s_out)
    e = torch.cuda.Event(); e.record(s_out); done[slot] = e
for i in range(4): step(i)
torch.cuda.synchronize(); marks.clear()
base = torch.cuda.Event(enable_timing=True); base.record(); torch.cuda.synchronize()
t0 = time.perf_counter()
N = 12
for i in range(4, 4 + N): step(i)
torch.cuda.synchronize()
print("wall per step %.3f ms" % ((time.perf_counter() - t0) * 1e3 / N))
for name, ev, th in marks[:40]:
    print("%-22s gpu %8.3f ms   host-issue %8.3f ms" % (name, base.elapsed_time(ev), (th - t0) * 1e3))
ok = all(torch.equal(hout[(4 + N - 1) % 2][k], s0[k].cpu()) for k in ()) 
