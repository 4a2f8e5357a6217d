import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, dssmd_b200, ref_torch
torch.set_printoptions(precision=3, linewidth=200, sci_mode=False)
os.environ["DSSMD_CORR_FWD_VARIANT"] = "2"; os.environ["DSSMD_CORR_DEBUG"] = "1"
B, C, H, W, D = 1, 16, 2, 64, 5
# structured inputs: L[c][x] = 1 for c==0 only; R[c][x] = x+1 for c==0  => out[d][x] = (x-d+1)/C
L = torch.zeros(B, C, H, W, device="cuda"); R = torch.zeros(B, C, H, W, device="cuda")
L[:, 0] = 1.0
R[:, 0] = torch.arange(1, W + 1, device="cuda").float().view(1, 1, W) + 100 * torch.arange(H, device="cuda").float().view(1, H, 1)
out = dssmd_b200.build_corr(L, R, D); torch.cuda.synchronize()
ref = ref_torch.corr_ref(L, R, D)
print("out*C row0 d=0:", (out[0, 0, 0] * C)[:40])
print("ref*C row0 d=0:", (ref[0, 0, 0] * C)[:40])
print("out*C row1 d=2:", (out[0, 2, 1] * C)[:40])
print("ref*C row1 d=2:", (ref[0, 2, 1] * C)[:40])
print("nonzero frac", float((out != 0).float().mean()), "max", float(out.abs().max()), "ref max", float(ref.abs().max()))
# channel test: L = 1 everywhere channel c; R = delta at channel c value (c+1)
L = torch.ones(B, C, H, W, device="cuda"); R = torch.zeros(B, C, H, W, device="cuda")
for c in range(C): R[:, c] = float(c + 1)
out = dssmd_b200.build_corr(L, R, D); torch.cuda.synchronize()
print("sum test out*C:", (out[0, 0, 0] * C)[:8], "expect", sum(range(1, C + 1)))
