#!/usr/bin/env python
"""Turn one gpurun profiling directory into the committed summaries under profiles/.

    python scripts/profile_summary.py gpurun_out/TAG ROUND_NAME [WORKLOAD]
Reads TAG/launches.csv (ncu --metrics gpu__time_duration.sum launch list of the default bench
command), TAG/prof.ncu-rep (ncu --set full of the same command) and TAG/bench_<workload>.json;
writes profiles/<ROUND_NAME>_summary.md, copies the launch list and updates profiles/traffic.json
(DRAM bytes per bench-level call, which bench.py reports as roofline.traffic)."""
import collections, csv, io, json, os, re, shutil, subprocess, sys

src, name = sys.argv[1], sys.argv[2]
workload = sys.argv[3] if len(sys.argv) > 3 else "sceneflow_576x960"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
prof = os.path.join(root, "profiles")

# bench-level call -> kernel-name substrings launched by it
CALLS = {"corr_fwd": ["corr_fwd"], "corr_bwd": ["corr_bwd_wide_kernel", "corr_bwd_fused_kernel"],
         "corr_bwd_gL": ["corr_bwd_reg_kernel<0", "corr_bwd_reg_kernel<float, 0", "corr_bwd_kernel<float, 0", "corr_bwd_mma32_kernel<0"],
         "corr_bwd_gR": ["corr_bwd_reg_kernel<1", "corr_bwd_reg_kernel<float, 1", "corr_bwd_kernel<float, 1", "corr_bwd_mma32_kernel<1"], "warp_fwd": ["warp_fwd"],
         "warp_bwd": ["warp_bwd"]}


def call_of(kname):
    for call, subs in CALLS.items():
        if any(s in kname for s in subs):
            return call
    return None


out = [f"# {name} — ncu summary of `python bench.py --workload {workload} --steps 4 --warmup 3 --no-graph --no-cpu-baseline`", ""]
launch_csv = os.path.join(src, "launches.csv")
shares = {}
if os.path.exists(launch_csv):
    rows = [r for r in csv.reader(open(launch_csv)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ix = {h: i for i, h in enumerate(hdr)}
    dur = collections.defaultdict(list)
    for r in rows:
        if r is hdr or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        dur[r[ix["Kernel Name"]]].append(v)
    ours = {k: v for k, v in dur.items() if "dssmd" in k or call_of(k)}
    tot = sum(sum(v) for v in ours.values()) or 1.0
    out += ["## Launch list (`--metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised)", "",
            "| kernel | launches | mean us | share of our kernels' time |", "|---|---|---|---|"]
    for k, v in sorted(ours.items(), key=lambda kv: -sum(kv[1])):
        short = re.sub(r"\(.*", "", k).replace("void ", "").replace("dssmd::<unnamed>::", "").replace("<unnamed>::", "")
        out.append(f"| `{short[:70]}` | {len(v)} | {sum(v)/len(v):.1f} | {sum(v)/tot:.3f} |")
        c = call_of(k)
        if c:
            shares[c] = shares.get(c, 0.0) + sum(v) / tot
    shutil.copy(launch_csv, os.path.join(prof, f"{name}_launches_{workload}.csv"))
    out += ["", "Share per bench-level call (ncu): " + ", ".join(f"{k} {v:.3f}" for k, v in sorted(shares.items(), key=lambda kv: -kv[1])), ""]

bj = os.path.join(src, f"bench_{workload}.json")
bench = json.load(open(bj)) if os.path.exists(bj) else None
if bench:
    tot = sum(k["us"] for k in bench["kernels"].values())
    out += ["## bench.py on the same box (CUDA events, graph replay, L2-rotating inputs)", "",
            "| call | us | algorithmic MB | GB/s | frac of measured HBM peak | share of step |", "|---|---|---|---|---|---|"]
    for k, v in bench["kernels"].items():
        out.append(f"| {k} | {v['us']:.1f} | {v['algorithmic_bytes']/1e6:.1f} | {v['GBps']:.0f} | {v['frac_of_peak']:.3f} | {v['us']/tot:.3f} |")
    out += ["", f"value {bench['value']:.0f} {bench['unit']}, {bench['ms_per_step']*1e3:.1f} us/step; e2e {bench['e2e']['value']:.0f} {bench['unit']} "
            f"({bench['e2e']['ms_per_step']:.2f} ms/step, {bench['e2e']['h2d_bytes_per_step']/1e6:.0f} MB up, {bench['e2e']['d2h_bytes_per_step']/1e6:.0f} MB down); "
            f"cpu_baseline {bench['cpu_baseline']['value']:.1f} {bench['unit']} on {bench['cpu_baseline']['cores']} cores." if bench.get("cpu_baseline") else "", ""]

rep = os.path.join(src, "prof.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    ix = {h: i for i, h in enumerate(rows[0])}
    per = collections.OrderedDict()
    for r in rows[2:]:
        k = r[ix["Kernel Name"]]
        if k in per:
            continue
        g = lambda m: float((r[ix[m]] or "0").replace(",", "")) if m in ix else 0.0
        unit = lambda m: rows[1][ix[m]] if m in ix else ""
        def bytes_of(m):
            v, u = g(m), unit(m)
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        per[k] = {"us": g("gpu__time_duration.sum") / (1e3 if unit("gpu__time_duration.sum") in ("ns", "nsecond") else 1),
                  "dram": bytes_of("dram__bytes_read.sum") + bytes_of("dram__bytes_write.sum"),
                  "dram_r": bytes_of("dram__bytes_read.sum"), "dram_w": bytes_of("dram__bytes_write.sum"),
                  "regs": g("launch__registers_per_thread"), "issue": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                  "occ": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
                  "dram_pct": g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")}
    out += ["## `ncu --set full --clock-control none --import-source on` (one launch per kernel)", "",
            "| kernel | us | DRAM read MB | DRAM write MB | regs | issue-active % | warps-active % | DRAM % of peak |", "|---|---|---|---|---|---|---|---|"]
    traffic, t_r, t_w = collections.defaultdict(float), collections.defaultdict(float), collections.defaultdict(float)
    for k, v in per.items():
        short = re.sub(r"\(.*", "", k).replace("void ", "").replace("<unnamed>::", "")
        out.append(f"| `{short[:70]}` | {v['us']:.1f} | {v['dram_r']/1e6:.1f} | {v['dram_w']/1e6:.1f} | {v['regs']:.0f} | {v['issue']:.0f} | {v['occ']:.0f} | {v['dram_pct']:.0f} |")
        c = call_of(k)
        if c:
            traffic[c] += v["dram"]; t_r[c] += v["dram_r"]; t_w[c] += v["dram_w"]
    tj = os.path.join(prof, "traffic.json")
    allt = json.load(open(tj)) if os.path.exists(tj) else {}
    allt[workload] = {c: {"dram_bytes": int(b), "dram_read_bytes": int(t_r[c]), "dram_write_bytes": int(t_w[c]),
                          "source": f"profiles/{name}_summary.md"} for c, b in traffic.items()}
    json.dump(allt, open(tj, "w"), indent=1, sort_keys=True)
    if bench:
        out += ["", "DRAM READ traffic vs algorithmic READ bytes per call (the writes of a 20-100 MB kernel are still dirty in the 126 MB L2 "
                "when it ends, so only the read side can show re-reads): " +
                ", ".join(f"{c} {t_r[c]/1e6:.1f}/{bench['kernels'][c].get('algorithmic_read_bytes', 0)/1e6:.1f} MB read, "
                          f"{t_w[c]/1e6:.1f}/{bench['kernels'][c].get('algorithmic_write_bytes', 0)/1e6:.1f} MB written" for c in traffic if c in bench["kernels"]), ""]
open(os.path.join(prof, f"{name}_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
