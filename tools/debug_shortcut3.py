"""Debug helper: K=3 case of test_simple_special_ties_in_the_dense_kernel: worst ties of the slab after 2 iterations."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import vimure_b200.synthetic as syn  # noqa: E402
from vimure_b200 import _packing  # noqa: E402
from vimure_b200._engine import CaviEngine  # noqa: E402

PRI = dict(alpha_theta=0.1, beta_theta=0.1, alpha_lambda=10.0, beta_lambda=10.0, alpha_eta=0.5, beta_eta=1.0)
L, N, K = 2, 640, 3
net = syn.Multitensor(N=N, L=L, K=K, C=2, avg_degree=8, eta=0.5, seed=3).build_X(mutuality=0.5, seed=4)
out = {}
for mode in ("short", "fp64"):
    if mode == "fp64":
        os.environ["VM_NO_SIMPLE"] = "1"
    P = _packing.pack(net.X.subs, net.X.vals, L, N, N, K, net.R, "cuda", tile_h=32)
    os.environ.pop("VM_NO_SIMPLE", None)
    eng = CaviEngine(P, PRI, mutuality=True, eps=1e-12)
    rs = np.random.RandomState(3).random_sample
    st = dict(gamma_shp=0.1 * rs((L, N)) + 0.1, phi_shp=10.0 * rs((L, K)) + 10.0, gamma_rte=0.1 * rs((L, N)) + 0.1,
              phi_rte=10.0 * rs((L, K)) + 10.0, nu_shp=0.5 * rs(1)[0] + 0.5)
    keep = (P.t["u_has_x"] & P.t["u_reported"]).cpu().numpy()
    pr_u = np.zeros((P.U, K))
    pr_u[:, 0] = 1.0
    pr = 1 + 0.01 * rs((int(keep.sum()), K))
    pr_u[keep] = pr / pr.sum(axis=1)[:, None]
    eng.set_state(st["gamma_shp"], st["gamma_rte"], st["phi_shp"], st["phi_rte"], st["nu_shp"],
                  1.0 + float(net.X.vals.sum()), pr_u, 1e-12)
    for it in range(2):
        eng.iterate(1)
        torch.cuda.synchronize()
        out[(mode, it)] = (eng.rho_u32.cpu().numpy().copy(), eng.rho_slab().cpu().numpy().copy())
    out[mode] = (P, eng)
P, eng = out["short"]
t = {k: v.cpu().numpy() for k, v in P.t.items() if k.startswith("u_")}
for it in range(2):
    a, b = out[("short", it)][0], out[("fp64", it)][0]
    d = np.abs(a - b).max(axis=1)
    w = np.argsort(-d)[:6]
    print("it", it, "rho_u32 worst", d[w])
    for u in w:
        print("  u", u, "lrow", t["u_lrow"][u], "col", t["u_col"][u], "cnt", t["u_cnt"][u], "m0", t["u_m0"][u], "x0", t["u_x0"][u],
              "xT0", t["u_xT0"][u], "px", t["u_px"][u], "pxt", t["u_pxt"][u], "simple", t["u_simple"][u], "single", t["u_single"][u],
              "short", a[u], "fp64", b[u])
    sa, sb = out[("short", it)][1], out[("fp64", it)][1]
    ds = np.abs(sa - sb)
    print("   slab max abs diff", ds.max(), "at", np.unravel_index(np.argmax(ds), ds.shape))
