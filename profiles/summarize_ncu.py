#!/usr/bin/env python3
"""Summarise an .ncu-rep (captured with `ncu --set full`) into the handful of numbers DESIGN/VERDICT
care about.  Usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 read sectors (from SMs)"),
    ("lts__t_sectors_srcunit_tex_op_write.sum", "L2 write sectors (from SMs)"),
    ("lts__t_sectors_srcunit_tex_op_atom.sum", "L2 atomic sectors"),
    ("lts__t_sectors_srcunit_tex_op_red.sum", "L2 reduction sectors"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard (warps/issue)"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_drain_per_issue_active.ratio", "stall drain"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (expected 0: hash/gather, not GEMM)"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        sys.exit("no data in " + rep)
    h, units = rows[0], rows[1]
    ki = h.index("Kernel Name")
    for r in rows[2:]:
        print(f"kernel: {r[ki]}")
        for key, label in WANT:
            if key in h:
                i = h.index(key)
                print(f"  {label:52s} {r[i]} {units[i]}")
        print()


if __name__ == "__main__":
    main()
