#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics per kernel + stall samples by opcode (needs ncu on PATH)."""
import collections
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum']
STALL = 'smsp__pcsamp_warps_issue_stalled_'


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    kre = sys.argv[2] if len(sys.argv) > 2 else None
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[ix['Kernel Name']]
        if kre and kre not in name:
            continue
        print("=====", name[:110])
        for wname in WANT:
            if wname in ix:
                print(f"  {wname:70s} {r[ix[wname]]}")
        st = {h[len(STALL):]: int(float(r[i] or 0)) for h, i in ix.items() if h.startswith(STALL) and 'not_issued' not in h}
        tot = sum(st.values()) or 1
        print("  stalls: " + ", ".join(f"{k}={100*v/tot:.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
    if len(sys.argv) > 3:
        src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{sys.argv[3]}"]))))
        h = None
        agg = collections.Counter(); ex = collections.Counter()
        for r in src:
            if r and r[0] == 'Address':
                h = {k: i for i, k in enumerate(r)}; agg.clear(); ex.clear(); continue
            if h is None or len(r) < len(h) // 2 or not r[0].startswith('0x'):
                continue
            toks = r[h['Source']].split()
            op = toks[1] if toks[0].startswith('@') else toks[0]
            agg[op] += int(r[h['# Samples']] or 0); ex[op] += int(r[h['Instructions Executed']] or 0)
        tot = sum(agg.values()) or 1
        print("  by opcode (samples%, executed):")
        for op, s in agg.most_common(14):
            print(f"    {op:24s} {100*s/tot:5.1f}%  {ex[op]:12d}")
        print("    total executed", sum(ex.values()))


if __name__ == "__main__":
    main()
