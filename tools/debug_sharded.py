"""Debug helper: 3 row-block shards vs one engine, one iteration without ELBO: which reporters' statistics differ."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import vimure_b200.synthetic as syn  # noqa: E402
from vimure_b200 import _packing  # noqa: E402
from vimure_b200._engine import CaviEngine  # noqa: E402
from vimure_b200.model import shard_rows  # noqa: E402

PRI = dict(alpha_theta=0.1, beta_theta=0.1, alpha_lambda=10.0, beta_lambda=10.0, alpha_eta=0.5, beta_eta=1.0)
L, N, K, W = 1, 1100, 2, 3
net = syn.StandardSBM(N=N, L=L, K=K, C=2, avg_degree=10, seed=10).build_X(mutuality=0.5, seed=20)
rs = np.random.RandomState(3).random_sample
st = dict(gamma_shp=0.1 * rs((L, N)) + 0.1, phi_shp=10.0 * rs((L, K)) + 10.0, gamma_rte=0.1 * rs((L, N)) + 0.1,
          phi_rte=10.0 * rs((L, K)) + 10.0, nu_shp=0.5 * rs(1)[0] + 0.5)
table = 1 + 0.01 * np.random.RandomState(4).random_sample((1 << 16, K))


def mk(row0, nloc):
    P = _packing.pack(net.X.subs, net.X.vals, L, N, N, K, net.R, "cuda", row0=row0, nloc=nloc, tile_h=32)
    e = CaviEngine(P, PRI, mutuality=True, eps=1e-12)
    keep = (P.t["u_has_x"] & P.t["u_reported"]).cpu().numpy()
    flat = P.t["u_gflat"].cpu().numpy()
    pr_u = np.zeros((P.U, K))
    pr_u[:, 0] = 1.0
    pr = table[(flat[keep] * 2654435761) % (1 << 16)]
    pr_u[keep] = pr / pr.sum(axis=1)[:, None]
    e.set_state(st["gamma_shp"], st["gamma_rte"], st["phi_shp"], st["phi_rte"], st["nu_shp"],
                1.0 + float(np.sum(net.X.vals)), pr_u, 1e-12)
    return e


one = mk(0, N)
engs = [mk(*shard_rows(N, W, r)) for r in range(W)]


def allreduce(attr):
    tot = sum(getattr(e, attr) for e in engs)
    for e in engs:
        getattr(e, attr).copy_(tot)


allreduce("red3")
print("init A diff", float((engs[0].red3 - one.red3).abs().max()))
for it in range(2):
    for e in engs:
        e.phase("gamma")
    allreduce("red1")
    for e in engs:
        e.phase("phi")
    allreduce("red2")
    for e in engs:
        e.phase("rho", 0)
    fa = sum(e.fixA for e in engs).cpu().numpy().reshape(N, K)
    rp = [e.rowpart.cpu().numpy() for e in engs]
    allreduce("red3")
    for e in engs:
        e.phase("finish", 0)
    one.phase("gamma"); one.phase("phi"); one.phase("rho", 0)
    f1 = one.fixA.cpu().numpy().reshape(N, K)
    A0 = engs[0].red3.cpu().numpy()[:N * K].reshape(N, K)
    A1 = one.red3.cpu().numpy()[:N * K].reshape(N, K)
    one.phase("finish", 0)
    bad = np.nonzero(np.abs(A0 - A1).max(axis=1) > 1e-6 * np.abs(A1).max(axis=1))[0]
    badf = np.nonzero(fa[:, 1] != f1[:, 1])[0]
    print("it", it, "A bad reporters", len(bad), bad[:6], bad[-3:] if len(bad) else "", "| fixA (sum of shards) != single:", len(badf),
          badf[:6], badf[-3:] if len(badf) else "")
    if len(bad):
        print("   A shards", A0[bad[:3]], "single", A1[bad[:3]], "fixA shards", fa[bad[:3]] / 2.0**44, "single", f1[bad[:3]] / 2.0**44)
    print("   extras", engs[0].red3.cpu().numpy()[N * K:], one.red3.cpu().numpy()[N * K:], "n_cx", [e.P.n_cx for e in engs], one.P.n_cx,
          "U", [e.P.U for e in engs], one.P.U, "simple flags", [float(e.layer_consts[2 * K + 4]) for e in engs])
