"""Aggregate an `ncu --csv` launch list by kernel: average duration (and DRAM bytes when captured), count, share."""
import csv
import sys


def main(path, top=30):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if r and r[0] == "ID":
            hdr, start = r, i + 1
            break
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = {}
    for r in rows[start:]:
        if len(r) <= vi:
            continue
        b = agg.setdefault(r[ki], {}).setdefault(r[mi], [0, 0.0])
        b[0] += 1
        b[1] += float(r[vi].replace(",", ""))
    T = "gpu__time_duration.sum"
    tot = sum(a[T][1] for a in agg.values())
    print("total %.1f us over %d launches" % (tot / 1e3, sum(a[T][0] for a in agg.values())))
    for n, a in sorted(agg.items(), key=lambda x: -x[1][T][1])[:top]:
        t = a[T]
        ex = ""
        if "dram__bytes_read.sum" in a:
            ex = " rd=%7.1f MB wr=%7.1f MB" % (a["dram__bytes_read.sum"][1] / t[0] / 1e6, a["dram__bytes_write.sum"][1] / t[0] / 1e6)
        print("%9.1f us avg  n=%4d  %5.1f%%%s  %s" % (t[1] / t[0] / 1e3, t[0], 100 * t[1] / tot, ex, n[:90]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
