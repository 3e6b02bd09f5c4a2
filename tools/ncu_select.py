#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` into the metrics the profiles/ summaries quote (one column per launch).

    python tools/ncu_select.py gpurun_out/prof.ncu-rep > profiles/rN_ncu_<kernel>.csv
"""
import csv
import subprocess
import sys

KEEP = ("Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct")


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, launches = rows[0], rows[1], rows[2:]
    w = csv.writer(sys.stdout)
    w.writerow(["metric", "unit"] + [r[hdr.index("Kernel Name")].split("(")[0] for r in launches])
    for i, h in enumerate(hdr):
        if h in KEEP or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
            w.writerow([h, units[i]] + [r[i] for r in launches])


if __name__ == "__main__":
    main()
