"""Debug helper: one iteration without ELBO with and without the shortcut ties, compare every intermediate."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import vimure_b200.synthetic as syn  # noqa: E402
from vimure_b200 import _packing  # noqa: E402
from vimure_b200._engine import CaviEngine  # noqa: E402

PRI = dict(alpha_theta=0.1, beta_theta=0.1, alpha_lambda=10.0, beta_lambda=10.0, alpha_eta=0.5, beta_eta=1.0)
N, K, L = 1100, 2, 1
net = syn.StandardSBM(N=N, L=L, K=K, C=2, avg_degree=10, seed=10).build_X(mutuality=0.5, seed=20)
res = {}
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
    for ph in ("gamma", "phi"):
        eng.phase(ph, 0)
    eng.phase("rho", 0)
    torch.cuda.synchronize()
    res[mode] = dict(A=eng.red3.cpu().numpy().copy(), fixA=eng.fixA.cpu().numpy().copy(), rho32=eng.rho_u32.cpu().numpy().copy(),
                     lc=eng.layer_consts.cpu().numpy().copy(), flags=eng.dev_flags.cpu().numpy().copy(),
                     fixP=eng.fixP.cpu().numpy().copy(), phi0=eng.phi0.cpu().numpy().copy(),
                     px=P.t["u_px"].cpu().numpy(), pxt=P.t["u_pxt"].cpu().numpy(), lrow=P.t["u_lrow"].cpu().numpy(),
                     col=P.t["u_col"].cpu().numpy(), ncx=P.n_cx, U=P.U, ncxblk=getattr(P, "n_cxblk", None), nublk=P.n_ublk,
                     nodetab=eng.nodetab.cpu().numpy().copy(), slab=eng.rho_slab().cpu().numpy().copy())
a, b = res["short"], res["fp64"]
print("U", a["U"], "n_cx", a["ncx"], "n_cxblk", a["ncxblk"], "n_ublk", a["nublk"], "lc", a["lc"], "flags", a["flags"])
A0, A1 = a["A"][:N * K].reshape(N, K), b["A"][:N * K].reshape(N, K)
bad = np.nonzero(np.abs(A0 - A1).max(axis=1) > 1e-6 * np.abs(A1).max(axis=1))[0]
print("reporters with different A:", len(bad), bad[:20], bad[-5:] if len(bad) else "")
print("A short", A0[bad[:5]], "A fp64", A1[bad[:5]])
print("extras short", a["A"][N * K:], "fp64", b["A"][N * K:])
f0, f1 = a["fixA"].reshape(N, K) / 2.0**44, b["fixA"].reshape(N, K) / 2.0**44
print("fixA diff reporters:", int((np.abs(f0 - f1).max(axis=1) > 1e-6).sum()), "max", np.abs(f0 - f1).max())
print("fixA short/ fp64 at bad:", f0[bad[:5]], f1[bad[:5]])
d = np.abs(a["rho32"] - b["rho32"]).max(axis=1)
sc = a["px"] > 0
print("rho_u32 max diff shortcut ties", d[sc].max(), "complex", d[~sc].max() if (~sc).any() else None)
w = np.argsort(-d)[:5]
print("worst ties", w, d[w], "px", a["px"][w], "pxt", a["pxt"][w], "lrow", a["lrow"][w], "col", a["col"][w],
      a["rho32"][w], b["rho32"][w])
print("slab max diff", np.abs(a["slab"] - b["slab"]).max())
print("phi0", a["phi0"], b["phi0"], "fixP", a["fixP"] / 2.0**30)
nt = a["nodetab"].reshape(N, 4)
print("nodetab[:3]", nt[:3], "act sum", nt[:, 3].sum())

# ---- several full iterations, single engine: shortcut vs fp64 path
print("---- multi-iteration")
engs = {}
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
    engs[mode] = eng
for it in range(4):
    out = {}
    for mode, eng in engs.items():
        eng.iterate(1)
        torch.cuda.synchronize()
        out[mode] = (eng.params(), eng.red3.cpu().numpy().copy(), eng.fixA.cpu().numpy().copy(), eng.dev_flags.cpu().numpy().copy(),
                     eng.layer_consts.cpu().numpy().copy())
    pa, pb = out["short"][0], out["fp64"][0]
    msg = []
    for k in ("gamma_shp", "gamma_rte", "phi_shp", "phi_rte", "nu_shp"):
        x, y = np.asarray(pa[k]), np.asarray(pb[k])
        msg.append("%s %.2e" % (k, float(np.max(np.abs(x - y) / np.abs(y)))))
    A0, A1 = out["short"][1][:N * K].reshape(N, K), out["fp64"][1][:N * K].reshape(N, K)
    badA = np.nonzero(np.abs(A0 - A1).max(axis=1) > 1e-6 * np.abs(A1).max(axis=1))[0]
    print("it", it, " ".join(msg), "| A bad reporters", len(badA), badA[:8], "extras", out["short"][1][N * K:], out["fp64"][1][N * K:],
          "flags", out["short"][3][:2], "simple flag", out["short"][4][2 * K + 4])
    if len(badA):
        print("   A short", A0[badA[:3]], "A fp64", A1[badA[:3]])
