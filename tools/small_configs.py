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

# ---- the restarts of one fit, one after the other vs side by side
import warnings  # noqa: E402

import vimure_b200 as vm  # noqa: E402
from tests.test_gpu_parity import build_inputs  # noqa: E402

for name in ("f1_over", "karnataka_vil1"):
    g = Golden(name)
    X, R = build_inputs(g)
    fk = dict(g.fit_kwargs)
    fk.update(num_realisations=5, max_iter=200)
    for conc in (False, True, False, True):
        m = vm.VimureModel(mutuality=True, convergence_tol=0.0)  # tol 0: every restart runs all 200 iterations
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            t0 = time.time()
            m.fit(X, R=R, init="fast", concurrent_realisations=conc, **fk)
            torch.cuda.synchronize()
            dt = time.time() - t0
        print(f"{name}: 5 restarts x 200 iterations, side by side={conc}: fit {dt*1e3:.1f} ms, CAVI loop "
              f"{m.timings['cavi_loop']*1e3:.1f} ms = {5*200/m.timings['cavi_loop']:.0f} iter/s, maxL={m.maxL:.6f}  "
              + " ".join(f"{k}={v*1e3:.1f}" for k, v in m.timings.items()))
