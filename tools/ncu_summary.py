#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full) into the handful of counters DESIGN.md / profiles/ quote.
usage: tools/ncu_summary.py report.ncu-rep [--md]"""
import csv, subprocess, sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall long_scoreboard"),
    ("smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "stall short_scoreboard"),
    ("smsp__average_warp_latency_issue_stalled_barrier.ratio", "stall barrier"),
    ("smsp__average_warp_latency_issue_stalled_wait.ratio", "stall wait"),
    ("smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "stall lg_throttle"),
    ("smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "stall math_pipe"),
    ("smsp__average_warp_latency_issue_stalled_sleeping.ratio", "stall sleeping"),
    ("smsp__average_warp_latency_issue_stalled_membar.ratio", "stall membar"),
    ("smsp__average_warp_latency_issue_stalled_branch_resolving.ratio", "stall branch"),
    ("smsp__average_warp_latency_issue_stalled_no_instruction.ratio", "stall no_inst"),
    ("smsp__average_warp_latency_issue_stalled_not_selected.ratio", "stall not_selected"),
    ("smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio", "stall dispatch"),
    ("smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "stall mio_throttle"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(f"### {d['Kernel Name'][:90]}  (id {d['ID']})")
        for k, label in KEYS:
            if k in d and d[k] != "":
                print(f"- {label}: {d[k]} {u.get(k, '')}")
        print()


if __name__ == "__main__":
    main()
