"""Why does (not) a layer of an all-reporter-mask problem take the fp32 special-tie path?  Prints the guard's ingredients."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vimure_b200 import _packing  # noqa: E402
from vimure_b200._engine import CaviEngine  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
L, K = 2, 2
dev = torch.device("cuda", 0)
net = bench.make_network(N, L, K, config="c4")
subs = torch.from_numpy(np.stack(net.X.subs).astype(np.int32)).to(dev)
vals = torch.from_numpy(np.asarray(net.X.vals).astype(np.int32)).to(dev)
P = _packing.pack(subs, vals, L, N, net.M, K, net.R, dev, tile_h=128)
eng = CaviEngine(P, bench.PRIORS, mutuality=True, eps=1e-12)
print("all32_mode", eng.all32_mode, "U", P.U, "I", P.I, "max x", int(vals.max()))
st, prng = bench.draw_state(L, net.M, K)
keep = P.t["u_has_x"] & P.t["u_reported"]
pr_u = torch.rand((P.U, K), dtype=torch.float64, device=dev).mul_(0.01).add_(1.0)
pr_u /= pr_u.sum(dim=-1, keepdim=True)
pr_u.masked_fill_(~keep[:, None], 0.0)
pr_u[:, 0].masked_fill_(~keep, 1.0)
eng.set_state(st["gamma_shp"], st["gamma_rte"], st["phi_shp"], st["phi_rte"], st["nu_shp"],
              bench.PRIORS["beta_eta"] + float(vals.sum()), pr_u, 1e-12)
print("simple_consts", eng.simple_consts.cpu().numpy())
for it in range(6):
    eng.iterate(1)
    lc = eng.layer_consts.cpu().numpy().reshape(L, 3 * K + 5)
    El = eng.E_lambda.cpu().numpy()
    Ell = eng.Elog_lambda.cpu().numpy()
    elt = eng.Elog_theta.cpu().numpy()
    print("it", it, "SIMPLE", lc[:, 2 * K + 4], "DEAD", lc[:, 2 * K + 3], "S_all", lc[:, 2 * K], "E_lambda", El.ravel(),
          "Elog_lambda", Ell.ravel(), "min Elog_theta", elt.min(axis=1), "G_nu", eng.nu.cpu().numpy())
eng.iterate(1, elbo_last=True)
print("elbo", eng.elbo())
