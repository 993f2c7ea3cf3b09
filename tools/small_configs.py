"""iter/s on the small BASELINE configs (C1: N=100 synthetic; C2: Karnataka vil1, 4 layers) from the golden fixtures."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.golden_util import Golden  # noqa: E402
from tests.test_gpu_parity import make_engine  # noqa: E402

for name in ("f1_over", "karnataka_vil1"):
    g = Golden(name)
    eng, P = make_engine(g)
    eng.iterate(20, elbo_last=True)
    torch.cuda.synchronize()
    for use_graph in (False, True):
        if use_graph and not hasattr(eng, "enable_graphs"):
            continue
        if use_graph:
            eng.enable_graphs()
        t0 = time.time()
        n = 2000
        for b in range(n // 10):
            eng.iterate(10, elbo_last=True)
            eng.elbo()
        torch.cuda.synchronize()
        dt = time.time() - t0
        print(f"{name}: N={g.N} L={g.L} K={g.K} nnzX={len(g.X_vals)} graphs={use_graph}: {n/dt:.0f} iter/s ({dt/n*1e6:.1f} us/iter, "
              f"ELBO every 10th)")
