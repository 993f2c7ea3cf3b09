"""One-line summary of a bench.py JSON line (file argument)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
c5 = (d.get("configs") or {}).get("c5") or {}
print("ms_per_step=%.4f dense_ms=%.4f dense_frac=%.3f step_frac=%.3f elbo=%r fp32_layers=%s c5_ms=%s e2e=%s" % (
    d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline_step"]["frac"], d["elbo_final"],
    d["config"].get("layers_on_the_fp32_special_tie_path"), c5.get("ms_per_step"),
    (d.get("e2e") or {}).get("wall_s_all_runs")))
