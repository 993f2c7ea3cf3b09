"""Print the headline metrics and top stall reasons of every kernel in an .ncu-rep (needs `ncu` on PATH)."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "sm__inst_executed_pipe_fp64.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum", "smsp__inst_executed_pipe_fp64.sum"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units = r[0], r[1]
    for row in r[2:]:
        print("---", row[hdr.index("Kernel Name")][:70])
        for w in KEYS:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:78s} {row[i]:>16s} {units[i]}")
        st = []
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h:
                try:
                    st.append((float(row[i]), h.split("issue_stalled_")[1].split("_per_issue")[0]))
                except Exception:
                    pass
        print("  stalls (warps per issue):", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:7]))


if __name__ == "__main__":
    main(sys.argv[1])
