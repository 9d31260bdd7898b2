#!/usr/bin/env python
"""Condense `ncu --set full` reports (gpurun_out/*.ncu-rep) into a small JSON of the metrics DESIGN.md cites.
Usage: python scripts/ncu_summary.py out.json name=report.ncu-rep [name=report.ncu-rep ...]"""
import csv
import io
import json
import subprocess
import sys

KEEP = ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_fmaheavy.sum", "smsp__inst_executed_pipe_lsu.sum",
        "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
        "launch__shared_mem_per_block_dynamic", "sm__inst_executed_pipe_fp32.sum")


def rows(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rd = list(csv.reader(io.StringIO(txt)))
    head, units = rd[0], rd[1]
    out = []
    for r in rd[2:]:
        d = {}
        for h, u, v in zip(head, units, r):
            if h in KEEP:
                d[h] = f"{v} {u}".strip() if u and h != "Kernel Name" else v
        out.append(d)
    return out


def main():
    dst = sys.argv[1]
    res = {}
    for arg in sys.argv[2:]:
        name, path = arg.split("=", 1)
        res[name] = rows(path)
    json.dump(res, open(dst, "w"), indent=1)
    for k, v in res.items():
        for r in v:
            print(k, r.get("Kernel Name", "")[:70], r.get("gpu__time_duration.sum"), r.get("smsp__inst_executed.sum"),
                  r.get("dram__bytes_read.sum"), r.get("dram__bytes_write.sum"), r.get("launch__registers_per_thread"))


if __name__ == "__main__":
    main()
