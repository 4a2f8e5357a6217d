#!/usr/bin/env python
"""Per-source-line instruction counts / stall samples for one kernel of an .ncu-rep (development tool).

    python scripts/ncu_lines.py REPORT.ncu-rep KERNEL_SUBSTRING [LIB.so]
Joins ncu's SASS page (per-instruction metrics) with nvdisasm's line info of the kernel in
the library that was profiled (the in-tree .so must be the build that was profiled)."""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, kern = sys.argv[1], sys.argv[2]
lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dssmd_b200", "libdssmd_b200.so")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows, h = [], None
for r in csv.reader(io.StringIO(out)):
    if r and r[0] == "Address":
        if h is not None:
            break
        h = {k: i for i, k in enumerate(r)}
        continue
    if h and r and r[0].startswith("0x"):
        rows.append(r)
base = int(rows[0][0], 16)
metrics = {int(r[0], 16) - base: (int(r[h["Instructions Executed"]] or 0), int(r[h["# Samples"]] or 0), r[h["Source"]].strip()) for r in rows}

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
line_of = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur_fn, cur_line, ok = None, None, False
    for ln in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            cur_fn = m.group(1); ok = kern in cur_fn and not line_of; cur_line = None
            continue
        if not ok:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            line_of[int(m.group(1), 16)] = cur_line
agg = collections.defaultdict(lambda: [0, 0])
for off, (ex, smp, src) in metrics.items():
    k = line_of.get(off)
    agg[k][0] += ex; agg[k][1] += smp
tot_ex = sum(v[0] for v in agg.values()) or 1
tot_s = sum(v[1] for v in agg.values()) or 1
srcs = {}
print(f"total executed {tot_ex}, samples {tot_s}, instructions mapped {sum(1 for o in metrics if line_of.get(o))}/{len(metrics)}")
for k, (ex, smp) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    text = ""
    if k:
        path = next((p for p in (os.path.join(os.path.dirname(lib), "csrc", k[0]),) if os.path.exists(p)), None)
        if path:
            srcs.setdefault(path, open(path).read().splitlines())
            text = srcs[path][k[1] - 1].strip()[:90]
    print(f"{str(k):28s} exec {100*ex/tot_ex:5.1f}%  samples {100*smp/tot_s:5.1f}%  {text}")
if os.environ.get("NCU_LINES_UNMAPPED"):
    print("---- unmapped instructions (offset, executed/warp-launch, sass) ----")
    n = 0
    for off in sorted(metrics):
        if line_of.get(off) is None and metrics[off][0] > 0:
            print(f"{off:6x} {metrics[off][0]:9d} {metrics[off][2][:100]}")
            n += 1
            if n > int(os.environ["NCU_LINES_UNMAPPED"]):
                break
