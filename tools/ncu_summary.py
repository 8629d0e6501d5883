#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full --import-source on) into the two text files kept under profiles/:
   <out>_metrics.csv  key counters per profiled launch,  <out>_lines.txt  stall samples / instructions per source line.
Runs here (no GPU needed): python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01b_ncu_tcg"""
import csv
import subprocess
import sys
from collections import defaultdict

WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        # L2 -> SM side (the weight ring): bytes and sectors the L2 served to the SMs' L1/TEX + async-copy path, L2 throughput
        "lts__t_sectors_srcunit_tex.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_bytes_equiv_l1sectormiss_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_sectors.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_sectors.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_uniform.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(out + "_metrics.csv", "w") as f:
        f.write("metric,unit," + ",".join(f"launch{i}" for i in range(len(data))) + "\n")
        for w in ["Kernel Name"] + WANT:
            if w in hdr:
                i = hdr.index(w)
                f.write(f"{w},{units[i]}," + ",".join('"' + r[i] + '"' if "," in r[i] else r[i] for r in data) + "\n")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    num = lambda x: int(x) if x.strip().isdigit() else 0
    h = None
    fname = ""
    m = {}
    cur = None
    tot_s = tot_i = 0
    stall_tot = defaultdict(int)
    for r in csv.reader(src.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            h = r
            cs, ci = h.index("# Samples"), h.index("Instructions Executed")
            sc = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
            continue
        if h is None or len(r) < len(h):
            continue
        if r[0] != "":
            k = (fname, int(r[0]))
            cur = m.setdefault(k, {"k": k, "src": r[1], "s": 0, "i": 0, "st": defaultdict(int)})
        else:
            s, n = num(r[cs]), num(r[ci])
            cur["s"] += s
            cur["i"] += n
            tot_s += s
            tot_i += n
            for c in sc:
                v = num(r[c])
                if v:
                    cur["st"][h[c]] += v
                    stall_tot[h[c]] += v
    with open(out + "_lines.txt", "w") as f:
        f.write(f"# {rep}: {tot_s} warp stall samples, {tot_i} warp instructions (all profiled launches)\n# stall reasons overall:\n")
        ssum = sum(stall_tot.values()) or 1
        for k, v in sorted(stall_tot.items(), key=lambda x: -x[1])[:10]:
            f.write(f"#   {k[6:]:20s} {100 * v / ssum:5.1f}%\n")
        f.write("# file:line  samples%  instructions%  top stall reasons  source\n")
        for l in sorted(m.values(), key=lambda x: -x["s"])[:45]:
            st = sorted(l["st"].items(), key=lambda x: -x[1])[:3]
            f.write(f"{l['k'][0]}:{l['k'][1]:<5d} {100 * l['s'] / max(tot_s, 1):5.1f}% {100 * l['i'] / max(tot_i, 1):5.1f}%  "
                    f"{','.join(a[6:] + '=' + str(b) for a, b in st):45s} {l['src'].strip()[:110]}\n")
    print("wrote", out + "_metrics.csv", out + "_lines.txt")


if __name__ == "__main__":
    main()
