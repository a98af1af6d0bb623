#!/usr/bin/env python
"""Condense an `ncu --set full` report into the handful of numbers the roofline discussion needs.
usage: tools/ncu_summary.py gpurun_out/x.ncu-rep [> profiles/x_summary.txt]   (runs on the CPU box: ncu -i ... --page raw)"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
    print(f"# {path}: ncu --set full --clock-control none (replayed, cold cache: use shares / percentages, not absolutes)")
    for r in rows[2:]:
        print(f"kernel: {r[name_col] if name_col is not None else '?'}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:68s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    main(sys.argv[1])
