import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, dssmd_b200, ref_torch
B, C, H, W, D = map(int, sys.argv[1:6])
os.environ["DSSMD_CORR_FWD_VARIANT"] = "2"; os.environ["DSSMD_CORR_DEBUG"] = "1"
g = torch.Generator(device="cuda").manual_seed(1)
L = torch.randn((B, C, H, W), generator=g, device="cuda"); R = torch.randn((B, C, H, W), generator=g, device="cuda")
ref = ref_torch.corr_ref(L.double(), R.double(), D)
out = dssmd_b200.build_corr(L, R, D); torch.cuda.synchronize()
print((B, C, H, W, D), "rel err", float((out.double() - ref).abs().max() / ref.abs().max()), flush=True)
