#!/usr/bin/env python
"""Short digest of one .ncu-rep (read on the CPU box): duration, occupancy, issue rate, shared-memory wavefronts by kind,
DRAM read / write bytes, the stall reasons, and (with --lines N) the N source lines with the most stall samples.
    python scripts/ncu_brief.py gpurun_out/TAG/prof.ncu-rep [--lines 25]"""
import csv
import io
import subprocess
import sys


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    rows = page(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}

    def g(k):
        v = m.get(k, ("nan", ""))[0]
        return v if v not in ("", "?") else "nan"
    print("kernel:", g("Kernel Name")[:90])
    print(f"duration {g('gpu__time_duration.sum')} us | regs {g('launch__registers_per_thread')} | warps active {float(g('sm__warps_active.avg.pct_of_peak_sustained_active')):.1f} % "
          f"| issue active {float(g('smsp__issue_active.avg.pct')):.1f} % | inst {float(g('smsp__inst_executed.sum'))/1e6:.2f} M")
    print(f"DRAM read {g('dram__bytes_read.sum')} {m.get('dram__bytes_read.sum',('',''))[1]} write {g('dram__bytes_write.sum')} {m.get('dram__bytes_write.sum',('',''))[1]} "
          f"| dram % {float(g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')):.1f} | l1tex data pipe % {float(g('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed')):.1f}")
    for k in ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum",
              "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
              "smsp__inst_executed_op_shared_atom.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum"):
        if k in m:
            print(f"  {k}: {float(g(k))/1e6:.3f} M")
    st = {h.split("issue_stalled_")[1].split("_per_issue")[0]: float(v) for h, v in zip(hdr, vals) if "issue_stalled" in h and "per_issue_active.ratio" in h}
    print("stalls per issue:", "  ".join(f"{k}:{v:.2f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:9]))
    if "--lines" in sys.argv:
        n = int(sys.argv[sys.argv.index("--lines") + 1])
        src = page(rep, "source")
        if src and src[0] and src[0][0] == "Kernel Name":
            src = src[1:]
        h = src[0]
        i_src, i_samp, i_ex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
        stall_cols = [(i, x) for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
        i_wf, i_wfi = h.index("L1 Wavefronts Shared"), h.index("L1 Wavefronts Shared Ideal")
        body = [r for r in src[1:] if len(r) > i_samp]
        tot = sum(float(r[i_samp] or 0) for r in body)
        print(f"top {n} SASS instructions by stall samples (of {tot:.0f}); columns: samples, share, executed, smem wavefronts (ideal), main stall")
        for r in sorted(body, key=lambda r: -float(r[i_samp] or 0))[:n]:
            main_stall = max(stall_cols, key=lambda c: float(r[c[0]] or 0))[1]
            print(f"  {float(r[i_samp]):6.0f} {100*float(r[i_samp])/tot:5.1f}% {float(r[i_ex]):8.0f} {float(r[i_wf] or 0):8.0f} ({float(r[i_wfi] or 0):7.0f}) {main_stall:18s} {r[i_src].strip()[:80]}")
        # stall samples by opcode class
        import collections
        by = collections.Counter()
        for r in body:
            toks = r[i_src].split()
            op = (toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")).split(".")[0]
            by[op] += float(r[i_samp] or 0)
        print("samples by opcode:", "  ".join(f"{k}:{100*v/tot:.1f}%" for k, v in by.most_common(14)))


if __name__ == "__main__":
    main()
