"""Condense an `ncu --page raw --csv` export into the handful of counters DESIGN.md / bench.py cite.

  ncu -i prof.ncu-rep --page raw --csv > prof_raw.csv
  python ncu_summary.py prof_raw.csv [kernel-substring] > profiles/<name>.json
One JSON object per profiled launch (the full capture replays each launch ~40 times; these are ncu's merged values).
"""
import csv
import json
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
    "smsp__inst_executed.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes.sum.per_second", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__m_xbar2l1tex_read_sectors_mem_texture.sum",
    "smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio", "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        if want and want not in name:
            continue
        d = {"kernel": name, "id": r[col["ID"]]}
        for k in KEYS:
            if k in col and r[col[k]] != "":
                try:
                    v = float(r[col[k]].replace(",", ""))
                except ValueError:
                    v = r[col[k]]
                d[k] = {"value": v, "unit": units[col[k]]}
        out.append(d)
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
