"""Dev tool: key numbers per kernel from an .ncu-rep (ncu -i ... --page raw --csv)."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "dur"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
    ("launch__block_size", "block"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__occupancy_limit_registers", "lim_regs"), ("launch__occupancy_limit_shared_mem", "lim_smem"),
    ("launch__occupancy_limit_warps", "lim_warps"),
    ("sm__inst_executed.sum", "inst"), ("smsp__inst_executed.sum", "inst_smsp"), ("sm__inst_executed.avg.per_cycle_active", "ipc"),
    ("smsp__issue_active.avg.pct", "issue%"), ("smsp__issue_inst0.avg.pct_of_peak_sustained_active", "noissue%"),
    ("smsp__average_warp_latency_per_inst_issued.ratio", "cyc/inst"), ("smsp__average_warps_issue_stalled_per_issue_active", "x"),
    ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__cycles_active.avg", "sm_cyc"), ("smsp__cycles_active.avg", "smsp_cyc"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf"),
    ("smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "st_long"),
    ("smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "st_short"),
    ("smsp__warp_issue_stalled_barrier_per_warp_active.pct", "st_bar"),
    ("smsp__warp_issue_stalled_wait_per_warp_active.pct", "st_wait"),
    ("smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "st_math"),
    ("smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "st_mio"),
    ("smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "st_lg"),
    ("smsp__warp_issue_stalled_not_selected_per_warp_active.pct", "st_notsel"),
    ("smsp__warp_issue_stalled_membar_per_warp_active.pct", "st_membar"),
    ("smsp__warp_issue_stalled_sleeping_per_warp_active.pct", "st_sleep"),
    ("smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct", "st_branch"),
    ("smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct", "st_disp"),
    ("smsp__warp_issue_stalled_no_instruction_per_warp_active.pct", "st_noinst"),
    ("smsp__warp_issue_stalled_imc_miss_per_warp_active.pct", "st_imc"),
    ("smsp__warp_issue_stalled_drain_per_warp_active.pct", "st_drain"),
    ("smsp__warp_issue_stalled_tex_throttle_per_warp_active.pct", "st_tex"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(d.get("Kernel Name", "")[:110])
        line = []
        for k, short in KEYS:
            if k in d and d[k] != "":
                line.append("%s=%s" % (short, d[k]))
        print("   " + "  ".join(line))


if __name__ == "__main__":
    main(sys.argv[1])
