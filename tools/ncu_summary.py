#!/usr/bin/env python
"""Condense an .ncu-rep (one kernel) into the handful of numbers the tuning notes quote.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--top 25]
"""
import csv
import io
import subprocess
import sys


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    r = page(rep, "raw")
    d = dict(zip(r[0], r[2]))
    keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
            "launch__occupancy_limit_registers", "smsp__warps_active.avg.per_cycle_active",
            "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
            "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct",
            "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "smsp__inst_executed_op_branch.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]
    for k in keys:
        print(f"{k}: {d.get(k)}")
    for k in ("dadd", "dmul", "dfma"):
        v = d.get(f"smsp__sass_thread_inst_executed_op_{k}_pred_on.sum.per_cycle_elapsed")
        print(f"thread {k}/cycle (all SMSPs): {v}")
    rows = []
    for k in d:
        if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued"):
            try:
                rows.append((float(d[k]), k.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    tot = sum(v for v, _ in rows) or 1.0
    print("stalls: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for v, k in sorted(rows, reverse=True)[:10]))
    s = page(rep, "source")
    if len(s) > 2:
        hdr = s[1]
        ix = {h: i for i, h in enumerate(hdr)}
        data = s[2:]

        def f(row, k):
            v = row[ix[k]] if k in ix else ""
            try:
                return int(v)
            except ValueError:
                return 0
        tot = sum(f(x, "# Samples") for x in data) or 1
        print(f"-- top {top} instructions by stall samples (of {tot}) --")
        names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        for row in sorted(data, key=lambda x: -f(x, "# Samples"))[:top]:
            st = sorted(((f(row, n), n[6:]) for n in names), reverse=True)[:3]
            print(f"{row[ix['Address']][-5:]} {100 * f(row, '# Samples') / tot:5.2f}%  exec {f(row, 'Instructions Executed'):>11}  "
                  f"{row[ix['Source']][:58]:58s} {st}")


if __name__ == "__main__":
    main()
