"""Per-instruction stall attribution from `ncu --page source --csv` (needs -lineinfo / --import-source)."""
import csv
import subprocess
import sys


def main(rep, kernel, top=25, reason=None):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for r in rows[2:]:
        if len(r) < 10 or r[0] == "Kernel Name":
            break
        if r[0] == "Address":
            continue
        try:
            d = {h: int(r[ix[h]]) for h in reasons}
            data.append((int(r[ix["Warp Stall Sampling (All Samples)"]]), int(r[ix["Instructions Executed"]]),
                         r[ix["Source"]].strip(), d, len(data)))
        except Exception:
            pass
    tot = sum(d[0] for d in data)
    agg = {h: sum(d[3][h] for d in data) for h in reasons}
    print("samples", tot, " by reason:", ", ".join(f"{h[6:]}={100*v/tot:.1f}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    key = (lambda d: d[3][reason]) if reason else (lambda d: d[0])
    for d in sorted(data, key=key, reverse=True)[:top]:
        main_r = max(d[3].items(), key=lambda kv: kv[1])
        print(f"{d[4]:5d} {d[0]:6d} {100*d[0]/tot:5.1f}% exec={d[1]:8d} {main_r[0][6:]:>12s}  {d[2]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25, sys.argv[4] if len(sys.argv) > 4 else None)
