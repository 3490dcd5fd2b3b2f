#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) or such a CSV export: one dict per captured
launch with the counters the roofline discussion uses.  Usage: tools/ncu_summary.py file.ncu-rep|file.csv [out.json]"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__t_bytes.sum": "l2_bytes",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_cycles_pct",
    "sm__inst_executed_pipe_tensor_op_dmma.avg.pct_of_peak_sustained_active": "dmma_pipe_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_cycles_pct",
    "sm__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__occupancy_limit_registers": "occ_lim_regs",
    "launch__occupancy_limit_shared_mem": "occ_lim_smem",
    "smsp__inst_executed.sum": "inst_executed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_sb",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_sb",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_throttle",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "stall_lg_throttle",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio": "stall_mio_throttle",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio": "stall_not_selected",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio": "stall_dispatch",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio": "stall_no_inst",
    "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio": "stall_imc",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio": "stall_branch",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio": "stall_membar",
}


def main():
    rep = sys.argv[1]
    if rep.endswith(".csv"):     # already exported with `ncu -i x.ncu-rep --page raw --csv`
        txt = open(rep).read()
    else:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")][:90]}
        for i, h in enumerate(hdr):
            if h in KEYS:
                try:
                    d[KEYS[h]] = float(r[i].replace(",", ""))
                except ValueError:
                    d[KEYS[h]] = r[i]
                if h in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum"):
                    d[KEYS[h] + "_unit"] = units[i]
        out.append(d)
    js = json.dumps(out, indent=1)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(js)
    print(js)


if __name__ == "__main__":
    main()
