"""Debug helper: run the first phases of a tiny ego-mask problem for several K, phase by phase, with VM_DEBUG_SYNC=1."""
import os
import sys

os.environ["VM_DEBUG_SYNC"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import vimure_b200.synthetic as syn  # noqa: E402
from vimure_b200 import _packing  # noqa: E402
from vimure_b200._engine import CaviEngine  # noqa: E402

PRI = dict(alpha_theta=0.1, beta_theta=0.1, alpha_lambda=10.0, beta_lambda=10.0, alpha_eta=0.5, beta_eta=1.0)
for K in [int(a) for a in sys.argv[1:]] or [3, 5, 8, 9, 12, 16]:
    N = 40
    net = syn.Multitensor(N=N, L=1, K=K, C=2, avg_degree=6, eta=0.5, seed=21).build_X(mutuality=0.5, seed=22)
    P = _packing.pack(net.X.subs, net.X.vals, 1, N, N, K, net.R, "cuda", tile_h=64)
    eng = CaviEngine(P, PRI, mutuality=True, eps=1e-12)
    rs = np.random.RandomState(1).random_sample
    pr_u = np.zeros((P.U, K))
    pr_u[:, 0] = 1.0
    eng.set_state(0.1 * rs((1, N)) + 0.1, 0.1 * rs((1, N)) + 0.1, 10 * rs((1, K)) + 10, 10 * rs((1, K)) + 10, 0.7,
                  1.0 + float(net.X.vals.sum()), pr_u, 1e-12)
    torch.cuda.synchronize()
    print("K", K, "U", P.U, "I1", P.I1, "n_gchunk", P.n_gchunk, "set_state ok", flush=True)
    for ph in ("gamma", "phi", "rho", "finish"):
        try:
            eng.phase(ph, 0)
            torch.cuda.synchronize()
            print("  phase", ph, "ok", flush=True)
        except Exception as e:  # noqa: BLE001
            print("  phase", ph, "FAILED:", str(e)[:200], flush=True)
            sys.exit(1)
print("all ok")
