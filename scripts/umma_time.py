import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("DSSMD_TUNING", "1")
import torch, dssmd_b200
sys.path.insert(0, os.path.join(ROOT, "tests"))
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
os.environ["DSSMD_CORR_FWD_VARIANT"] = "2"
for (B, C, H, W, D) in [(8, 256, 80, 160, 41), (16, 128, 48, 160, 24)]:
    L = torch.randn(B, C, H, W, device="cuda").relu(); R = torch.randn(B, C, H, W, device="cuda").relu()
    ref = torch.stack([torch.cat([torch.zeros(B, H, d, device="cuda", dtype=torch.float64), (L[..., d:].double() * R[..., :W - d].double()).mean(1)], -1) if d else (L.double() * R.double()).mean(1) for d in range(D)], 1)
    for env in [{}, {"DSSMD_UMMA_SKIP": "16"}]:
        os.environ.pop("DSSMD_UMMA_SKIP", None)
        os.environ.update(env)
        out = dssmd_b200.build_corr(L, R, D)
        err = float((out.double() - ref).abs().max() / ref.abs().max())
        print((B, C, H, W, D), env, round(timeit(lambda: dssmd_b200.build_corr(L, R, D)), 1), "us  rel err %.3e" % err, flush=True)
