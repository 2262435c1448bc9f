#!/usr/bin/env python
"""Summarises gpurun_out/<tag>/{launches.csv, *.ncu-rep} into profiles/<tag>_*.md (the tracked evidence).

    python scripts/ncu_summary.py gpurun_out/r01a profiles/r01a
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_allocated",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__thread_inst_executed_pred_on_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = None
    agg = collections.OrderedDict()
    for r in rows:
        if r[0] == "ID":
            hdr = r; continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("ls1b200::", "")
        v = float(d["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(d["Metric Unit"], 1)
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| {k} | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.1f} | {v[1] / tot:.3f} |\n")


def raw(rep, out, title):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# {title}\n\nsource: `ncu --set full --clock-control none --import-source on` ({rep}); one column per captured launch\n\n")
        names = [r[hdr.index("Kernel Name")].split("(")[0] for r in rows[2:]]
        f.write("| metric | unit | " + " | ".join(n[-40:] for n in names) + " |\n|---|---|" + "---|" * len(names) + "\n")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                f.write(f"| {k} | {units[i]} | " + " | ".join(r[i] for r in rows[2:]) + " |\n")


if __name__ == "__main__":
    src, dst = sys.argv[1], sys.argv[2]
    import glob, os
    if os.path.exists(os.path.join(src, "launches.csv")):
        launches(os.path.join(src, "launches.csv"), dst + "_launches.md")
    for rep in glob.glob(os.path.join(src, "*.ncu-rep")):
        raw(rep, dst + "_" + os.path.basename(rep)[:-8] + ".md", os.path.basename(rep))
