#!/usr/bin/env python
"""Condenses an `ncu --set full` report into the handful of numbers DESIGN.md / bench.py cite.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv
    ncu -i gpurun_out/prof.ncu-rep --page source --csv > src.csv     (optional)
    python tools/ncu_summary.py raw.csv [src.csv] > profiles/rNN_fused_kernel.txt"""
import collections
import csv
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(units, vals)))
        print("kernel:", d.get("Kernel Name", ("", "?"))[1])
        for k in KEYS:
            if k in d:
                print("  %-78s %14s %s" % (k, d[k][1], d[k][0]))
        for k in sorted(d):
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and float(d[k][1] or 0) > 0.02:
                print("  %-78s %14s" % (k.replace("smsp__average_warps_issue_stalled_", "stall/issue: "), d[k][1]))
    if len(sys.argv) > 2:
        rows = list(csv.reader(open(sys.argv[2])))
        hdr = rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        data = rows[2:]
        tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
        opc, smp = collections.Counter(), collections.Counter()
        for r in data:
            op = [t for t in r[ix["Source"]].split() if not t.startswith("@")][0].split(".")[0]
            opc[op] += int(r[ix["Instructions Executed"]])
            smp[op] += int(r[ix["# Samples"]])
        total = sum(opc.values())
        print("executed warp instructions: %d" % total)
        print("  %-10s %9s %9s" % ("opcode", "% instrs", "% samples"))
        for k, v in opc.most_common(24):
            print("  %-10s %8.2f%% %8.2f%%" % (k, 100.0 * v / total, 100.0 * smp[k] / tot))


if __name__ == "__main__":
    main()
