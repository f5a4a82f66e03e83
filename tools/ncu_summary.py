#!/usr/bin/env python3
"""Condenses an `ncu --page raw --csv` export (+ optional `--page source --csv`) into the handful of numbers
DESIGN.md / profiles/ quote.  Usage: tools/ncu_summary.py raw.csv [src.csv] [warp_steps]"""
import collections
import csv
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second", "lts__t_bytes.sum",
        "smsp__inst_executed.sum", "sass__inst_executed_register_spilling", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__cycles_active.avg", "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")][:100])
        for h, u, v in zip(hdr, units, r):
            if h in KEYS:
                print(f"  {h:75s} {v} {u}")
        st = [(float(v), h) for h, v in zip(hdr, r) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
        print("  stalls per issue:", ", ".join(f"{h[34:-23]}={v:.2f}" for v, h in sorted(st, reverse=True)[:7]))
    if len(sys.argv) > 2:
        rows = list(csv.reader(open(sys.argv[2])))
        hdr = rows[1]
        ia, isamp, iex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
        data = []
        for r in rows[2:]:
            if r and r[0] == "Kernel Name":
                break
            if len(r) > iex:
                data.append(r)
        ws = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
        tot_s = sum(int(r[isamp] or 0) for r in data)
        tot_e = sum(int(r[iex] or 0) for r in data)
        print(f"dynamic warp-instructions: {tot_e}  (= {tot_e / ws:.1f} per warp-step)")
        byop = collections.defaultdict(lambda: [0, 0])
        for r in data:
            t = r[ia].split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            byop[op][0] += int(r[isamp] or 0)
            byop[op][1] += int(r[iex] or 0)
        for op, (s, e) in sorted(byop.items(), key=lambda kv: -kv[1][1])[:22]:
            print(f"  {op:10s} {e / ws:8.1f} per warp-step   stall samples {100 * s / max(tot_s, 1):5.1f}%")


main()
