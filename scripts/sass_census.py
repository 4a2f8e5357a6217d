#!/usr/bin/env python
"""Instruction census of one kernel's hot loop from the object file (development tool, no GPU needed):

    python scripts/sass_census.py dssmd_b200/_obj/warp_bwd_band2.o 'band2_kernelILi3ELi258ELb1' [--lines]

Disassembles with source line info (nvdisasm -g), finds the largest backward branch (the row loop) and prints the
number of SASS instructions inside it, by opcode and (with --lines) by source line.  The kernels here are bound by
instruction issue (profiles/r01h_summary.md), so this count is the offline proxy for their run time."""
import collections
import os
import re
import subprocess
import sys
import tempfile


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    by_line = "--lines" in sys.argv
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
    start = next(i for i, l in enumerate(txt) if l.startswith(".text.") and re.search(pat, l))
    body = []
    for l in txt[start + 1:]:
        if l.startswith("\t.section") or l.startswith(".text."):
            break
        body.append(l)
    ins = []      # (addr, opcode, srcline, text)
    cur = None
    for l in body:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
        if m:
            text = m.group(2).strip()
            toks = text.split()
            op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            ins.append((int(m.group(1), 16), op.split(".")[0], cur, text))
    # largest backward branch
    best = None
    labels = {}
    for l in body:
        m = re.match(r"^(\.L_x_\d+):", l)
        if m:
            labels[m.group(1)] = None
    # label addresses: the address of the next instruction after the label line
    pending = []
    for l in body:
        m = re.match(r"^(\.L_x_\d+):", l)
        if m:
            pending.append(m.group(1))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/", l)
        if m and pending:
            for p in pending:
                labels[p] = int(m.group(1), 16)
            pending = []
    for a, op, _, text in ins:
        if op == "BRA":
            m = re.search(r"`\((\.L_x_\d+)\)", text)
            if m and labels.get(m.group(1)) is not None and labels[m.group(1)] < a:
                span = a - labels[m.group(1)]
                if best is None or span > best[0]:
                    best = (span, labels[m.group(1)], a)
    if best is None:
        print("no loop found; total instructions:", len(ins))
        return
    _, lo, hi = best
    loop = [i for i in ins if lo <= i[0] <= hi]
    print(f"kernel {pat}: {len(ins)} instructions, row loop {lo:#x}..{hi:#x}: {len(loop)} instructions (static)")
    ops = collections.Counter(i[1] for i in loop)
    print("  " + "  ".join(f"{k}:{v}" for k, v in ops.most_common(28)))
    if by_line:
        lines = collections.Counter(i[2] for i in loop)
        for k, v in sorted(lines.items(), key=lambda kv: (kv[0][0], kv[0][1]) if kv[0] else ("", 0)):
            print(f"   {k}: {v}")


if __name__ == "__main__":
    main()
