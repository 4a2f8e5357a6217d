import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, dssmd_b200
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
os.environ["DSSMD_CORR_FWD_VARIANT"] = "2"
for (B, C, H, W, D) in [(8, 256, 80, 160, 41)]:
    L = torch.randn(B, C, H, W, device="cuda").relu(); R = torch.randn(B, C, H, W, device="cuda").relu()
    for env in [{}, {"DSSMD_UMMA_SKIP": "1"}, {"DSSMD_UMMA_SKIP": "2"}, {"DSSMD_UMMA_SKIP": "4"}, {"DSSMD_UMMA_SKIP": "8"}, {"DSSMD_UMMA_SKIP": "15"}, {"DSSMD_UMMA_SKIP": "11"}]:
        for k in ("DSSMD_UMMA_PASSES", "DSSMD_UMMA_N256", "DSSMD_UMMA_SKIP"): os.environ.pop(k, None)
        os.environ.update(env)
        print((B, C, H, W, D), env, round(timeit(lambda: dssmd_b200.build_corr(L, R, D)), 1), "us", flush=True)
