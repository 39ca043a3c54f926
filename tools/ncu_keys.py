"""Key per-launch metrics of every kernel in an .ncu-rep (raw page), for quick comparisons.
usage: python tools/ncu_keys.py <report.ncu-rep> [kernel-substring]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
]


def main():
    rep = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if sub not in d["Kernel Name"]:
            continue
        print("##", d["Kernel Name"][:90], "grid", d.get("Grid Size"), "block", d.get("Block Size"))
        for k in KEYS:
            if k in d:
                print("  %-75s %s %s" % (k, d[k], units[hdr.index(k)]))
        for k in hdr:
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
                try:
                    v = float(d[k])
                except ValueError:
                    continue
                if v >= 0.3:
                    print("  stall %-69s %.2f" % (k.split("issue_stalled_")[1].split("_per_issue")[0], v))


if __name__ == "__main__":
    main()
