#!/usr/bin/env python
"""tools/ncu_summary.py REPORT.ncu-rep [...] -- one compact line block per captured kernel launch
(the numbers DESIGN.md and profiles/ quote).  Reads the report with `ncu -i ... --page raw --csv`."""
import csv
import io
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("launch__waves_per_multiprocessor", "waves"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long_sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short_sb"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg_thr"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio_thr"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64pipe%"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64cyc%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "fmaheavy%"),
        ("sm__inst_executed.avg.per_cycle_elapsed", "ipc"),
        ("smsp__inst_executed.sum", "warp_inst"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "ld_sectors"),
        ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "ld_requests"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "st_sectors"),
        ("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "st_requests"),
        ("lts__t_sectors_op_read.sum", "l2_rd_sectors"), ("lts__t_sectors_op_write.sum", "l2_wr_sectors")]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(path, "(empty)")
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            print("%s :: %s" % (path.split("/")[-1], name[:90]))
            items = []
            for key, short in WANT:
                if key in hdr:
                    i = hdr.index(key)
                    v = r[i]
                    try:
                        v = "%.4g" % float(v.replace(",", ""))
                    except ValueError:
                        pass
                    items.append("%s=%s%s" % (short, v, (" " + units[i]) if units[i] and units[i] not in ("%", "inst") else ""))
            for k in range(0, len(items), 6):
                print("    " + "  ".join(items[k:k + 6]))


if __name__ == "__main__":
    main()
