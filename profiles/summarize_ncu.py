#!/usr/bin/env python
"""Turns an .ncu-rep (ncu --set full) into the short text summary kept under profiles/.
usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep "title" > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def main():
    rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units = rows[0], rows[1]
    print(f"# {title}\n# source: {rep} (ncu --set full --clock-control none --import-source on)\n")
    for r in rows[2:]:
        print("kernel:", r[head.index("Kernel Name")])
        for k in KEYS:
            if k in head:
                i = head.index(k)
                print(f"  {k:72s} {r[i]:>16s} {units[i]}")
        stalls = [(float(r[i].replace(",", "")), n) for i, n in enumerate(head)
                  if n.startswith("smsp__average_warps_issue_stalled") and n.endswith("per_issue_active.ratio")]
        print("  warp stall reasons (warps per issue):")
        for v, n in sorted(stalls, reverse=True)[:8]:
            print(f"    {v:7.3f}  " + n.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
        pipes = [(float(r[i].replace(",", "")), n) for i, n in enumerate(head)
                 if n.startswith("smsp__inst_executed_pipe_") and n.endswith(".sum")]
        print("  instructions by pipe (millions of warp instructions):")
        for v, n in sorted(pipes, reverse=True)[:8]:
            if v > 0:
                print(f"    {v / 1e6:9.1f}  " + n.replace("smsp__inst_executed_pipe_", "").replace(".sum", ""))
        print()


if __name__ == "__main__":
    main()
