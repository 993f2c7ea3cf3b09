"""Key metrics of every kernel in an .ncu-rep (`ncu --set full`): time, DRAM bytes, occupancy, issue utilisation, stalls."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]


def main(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if "average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    for r in rows[2:]:
        print("==== " + r[ix["Kernel Name"]])
        for w in WANT:
            if w in ix:
                print("  %-62s %s %s" % (w, r[ix[w]], units[ix[w]]))
        top = sorted(((float(r[ix[h]]), h) for h in stalls if r[ix[h]] not in ("", "n/a")), reverse=True)[:6]
        print("  warps stalled per issue (top): " + ", ".join("%s=%.2f" % (h.split("issue_stalled_")[1].split("_per_")[0], v) for v, h in top))


if __name__ == "__main__":
    main(sys.argv[1])
