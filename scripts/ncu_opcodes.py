#!/usr/bin/env python
"""Executed warp-instructions by opcode for one kernel of an .ncu-rep (development tool).

    python scripts/ncu_opcodes.py REPORT.ncu-rep KERNEL_SUBSTRING [--dump]"""
import collections, csv, io, subprocess, sys

rep, kern = sys.argv[1], sys.argv[2]
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
ops = collections.Counter()
tot = 0
for r in rows:
    n = int(r[h["Instructions Executed"]] or 0)
    src = r[h["Source"]].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    ops[op.split(".")[0]] += n
    tot += n
    if "--dump" in sys.argv:
        print(f"{n:10d} {int(r[h['# Samples']] or 0):5d}  {src}")
print("total warp-instructions executed", tot)
for op, n in ops.most_common(40):
    print(f"  {op:12s} {n:10d} {100.0*n/tot:5.1f}%")
