#!/usr/bin/env python
"""ncu_summary.py <report.ncu-rep> <out.csv> -- per-launch summary of an `ncu --set full` capture (run where ncu is installed):
duration, DRAM bytes read/written, DRAM %, registers, occupancy, stall reasons.  The csv goes under profiles/."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("Kernel Name", "kernel"), ("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"), ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__shared_mem_per_block_dynamic", "smem_dyn"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "ld_sectors"), ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "ld_requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "st_sectors"), ("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "st_requests"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg"),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    cols = [(hdr.index(k), n) for k, n in KEYS if k in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([n + (f" [{units[i]}]" if units[i] else "") for i, n in cols])
        for r in rows[2:]:
            w.writerow([r[i] for i, _ in cols])
    print(f"{len(rows) - 2} launches -> {out}")


if __name__ == "__main__":
    main()
