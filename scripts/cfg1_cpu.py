"""BASELINE config 1: upstream SSMDNet (BMVC 2018) forward, batch 1, one synthetic hazy 320x640 stereo pair, repo-default
max_disp, fp32 on CPU -- the reference's own CPU path, timed where the upstream tree exists (the build container).
    python scripts/cfg1_cpu.py > profiles/r02_cfg1_cpu.json"""
import json, os, sys, time, warnings
warnings.filterwarnings("ignore")
sys.dont_write_bytecode = True
REF = os.environ.get("DSSMD_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
import torch
from models.SSMDNet import SSMDNet
from models.UtilsNet.submoudles import build_corr
torch.manual_seed(1024)
net = SSMDNet(in_channels=3, max_disp=192, dehaze_switch=True).eval()
l, r = torch.rand(1, 3, 320, 640), torch.rand(1, 3, 320, 640)
threads = torch.get_num_threads()
with torch.no_grad():
    net(l, r)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); out = net(l, r); ts.append(time.perf_counter() - t0)
    L, R = torch.rand(1, 128, 40, 80).relu(), torch.rand(1, 128, 40, 80).relu()
    build_corr(L, R, 24)
    tc = []
    for _ in range(10):
        t0 = time.perf_counter(); build_corr(L, R, 24); tc.append(time.perf_counter() - t0)
cpu = [l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")]
print(json.dumps({"config": "cfg1: models/SSMDNet.py:275-301, SSMDNet(in_channels=3, max_disp=192, dehaze_switch=True).eval() forward, B1 3x320x640 fp32, CPU",
                  "where": "build container (the only place the upstream tree exists)", "torch": torch.__version__, "threads": threads,
                  "cpu": cpu[0] if cpu else None, "logical_cpus": os.cpu_count(),
                  "forward_s_best_of_5": round(min(ts), 4), "forward_s_median": round(sorted(ts)[2], 4), "pairs_per_s": round(1 / min(ts), 2),
                  "pyramid": [list(p.shape) for p in out[0]],
                  "build_corr_ms_at_model_shape_B1_C128_40x80_D24": round(min(tc) * 1e3, 3)}, indent=1))
