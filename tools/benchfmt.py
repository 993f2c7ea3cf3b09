import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["elbo_final"])
    else: print(l[:300])
