"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import sys


def main(path, top=25):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
        agg[row["Kernel Name"].split("(")[0][:60]].append(v)
    tot = sum(sum(v) for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:top]:
        print(f"{k:60s} n={len(v):4d} avg={sum(v)/len(v):10.1f}us total={sum(v)/1000:8.2f}ms {100*sum(v)/tot:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
