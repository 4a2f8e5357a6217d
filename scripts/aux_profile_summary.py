#!/usr/bin/env python
"""profiles/<NAME>_aux_summary.md from one gpurun directory of scripts/gpu_aux.sh:
    python scripts/aux_profile_summary.py gpurun_out/TAG NAME
Launch list of scripts/aux_step.py (our kernels only), the ncu --set full metrics per kernel, and the haze / ASM-loss
bench log (CUDA events, not under ncu)."""
import collections, csv, io, os, re, shutil, subprocess, sys

src, name = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
prof = os.path.join(root, "profiles")
OURS = ("haze_kernel", "asm_recovered", "corr_fwd", "corr_bwd", "warp_")


def short(k):
    k = re.sub(r"\(.*", "", k).replace("void ", "")
    return k.replace("dssmd::<unnamed>::", "").replace("<unnamed>::", "").replace("unnamed>::", "")


out = [f"# {name} — input-side and loss-side kernels (`python scripts/aux_step.py 3`: haze synthesis of both views, correlation "
       "into the concat buffer fwd + bwd, ASM reconstruction loss over the 4-level pyramid; B4 576x960)", ""]
rows = [r for r in csv.reader(open(os.path.join(src, "aux_launches.csv"))) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ix = {h: i for i, h in enumerate(hdr)}
dur = collections.defaultdict(list)
other = 0.0
for r in rows:
    if r is hdr or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", "")) / 1e3
    k = r[ix["Kernel Name"]]
    if any(o in k for o in OURS):
        dur[short(k)].append(v)
    else:
        other += v
tot = sum(sum(v) for v in dur.values())
out += ["## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised)", "",
        "| kernel | launches | mean us | share of our kernels' time |", "|---|---|---|---|"]
for k, v in sorted(dur.items(), key=lambda kv: -sum(kv[1])):
    out.append(f"| `{k}` | {len(v)} | {sum(v) / len(v):.1f} | {sum(v) / tot:.3f} |")
out += ["", f"Other kernels in the same process (torch: input generation, partial-sum reductions, autograd glue): {other:.0f} us "
        f"in total against {tot:.0f} us of ours.", ""]
shutil.copy(os.path.join(src, "aux_launches.csv"), os.path.join(prof, f"{name}_aux_launches.csv"))

rep = os.path.join(src, "aux_prof.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    h = {k: i for i, k in enumerate(rr[0])}
    out += ["## `ncu --set full --clock-control none --import-source on` (one launch per kernel)", "",
            "| kernel | us | DRAM read MB | DRAM write MB | regs | issue-active % | warps-active % | DRAM % of peak | top stalls |", "|---|---|---|---|---|---|---|---|---|"]
    seen = set()
    for r in rr[2:]:
        k = short(r[h["Kernel Name"]])
        if k in seen:
            continue
        seen.add(k)
        st = {c[len("smsp__pcsamp_warps_issue_stalled_"):]: float(r[i] or 0) for c, i in h.items()
              if c.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in c}
        ts = sum(st.values()) or 1
        top = ", ".join(f"{a} {100 * b / ts:.0f}%" for a, b in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        g = lambda m: float(r[h[m]].replace(",", ""))  # noqa: E731
        out.append(f"| `{k}` | {g('gpu__time_duration.sum'):.1f} | {g('dram__bytes_read.sum'):.1f} | {g('dram__bytes_write.sum'):.1f} | "
                   f"{int(g('launch__registers_per_thread'))} | {g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f} | "
                   f"{g('sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} | {g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | {top} |")
    out.append("")
log = os.path.join(src, "haze_bench.log")
if os.path.exists(log):
    out += ["## `python scripts/haze_bench.py` on the same box (CUDA events, inputs rotating over sets larger than L2; "
            "`kernel` = CUDA-graph replay of the C-ABI launch, the other columns include the Python launch path)", "", "```"]
    out += [l.rstrip() for l in open(log)]
    out += ["```", ""]
open(os.path.join(prof, f"{name}_aux_summary.md"), "w").write("\n".join(out))
print("\n".join(out))
