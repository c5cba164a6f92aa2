#!/usr/bin/env python3
"""Summarise an .ncu-rep (first profiled kernel) into the metrics DESIGN.md argues with.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/xxx.txt]"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__thread_inst_executed_per_inst_executed.pct",
    "smsp__sass_average_branch_targets_threads_uniform.pct", "smsp__warps_eligible.avg.per_cycle_active", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_bytes.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fmaheavy.sum",
    # collected with --metrics next to --set full (tools/gpu_r02_l.sh): L2 sectors, atomics, instruction caches
    "lts__t_sectors.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_atom.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum", "sm__icc_request_hit_rate.pct", "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed",
]
STALLS = "smsp__average_warps_issue_stalled_{}_per_issue_active.ratio"
REASONS = ["long_scoreboard", "short_scoreboard", "wait", "math_pipe_throttle", "branch_resolving", "no_instruction", "not_selected", "selected", "dispatch_stall",
           "lg_throttle", "mio_throttle", "tex_throttle", "barrier", "membar", "drain", "imc_miss", "sleeping", "misc"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    print("kernel:", vals[col["Kernel Name"]][:100])
    for w in WANT + [STALLS.format(r) for r in REASONS]:
        if w in col:
            print(f"{w:90s} {vals[col[w]]:>18s} {units[col[w]]}")


if __name__ == "__main__":
    main()
