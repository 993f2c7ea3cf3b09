"""A/B timing of several builds of libvimure_b200.so in ONE process on the same packed network and initial state.

    python tools/ab_libs.py [--config c3|c5s] [--nodes N] [--iters 40] name=path.so [name=path.so ...]

For every library: set the same state, run `--iters` CAVI iterations without the ELBO (CUDA events), then one ELBO
iteration; prints ms/iteration, the final ELBO and the largest relative difference of the gamma/phi/nu posteriors against
the FIRST library (a full-size parity check between kernel variants).
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vimure_b200 import _capi, _packing  # noqa: E402
from vimure_b200._engine import CaviEngine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c3")
    ap.add_argument("--nodes", type=int, default=0)
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("libs", nargs="+")
    a = ap.parse_args()
    L, K, N = 1, 2, a.nodes or 20000
    if a.config == "c5s":
        L, K, N = 4, 3, a.nodes or 16000
    dev = torch.device("cuda", 0)
    net = bench.make_network(N, L, K, config=a.config)
    P = _packing.pack(net.X.subs, net.X.vals, L, N, net.M, K, net.R, dev, tile_h=128)
    st, prng = bench.draw_state(L, net.M, K)
    keep = (P.t["u_has_x"] & P.t["u_reported"]).cpu().numpy()
    pr_u = np.zeros((P.U, K))
    pr_u[:, 0] = 1.0
    pr = 1 + 0.01 * prng.random_sample((int(keep.sum()), K))
    pr_u[keep] = pr / pr.sum(axis=1)[:, None]
    nu_rte = bench.PRIORS["beta_eta"] + float(net.X.vals.sum())
    eng = CaviEngine(P, bench.PRIORS, mutuality=True, eps=1e-12)
    ref = None
    print("config %s N=%d L=%d K=%d U=%d I=%d" % (a.config, N, L, K, P.U, P.I))
    for spec in a.libs:
        name, path = spec.split("=", 1)
        eng.lib = _capi.open_library(os.path.abspath(path))
        eng.set_state(st["gamma_shp"], st["gamma_rte"], st["phi_shp"], st["phi_rte"], st["nu_shp"], nu_rte, pr_u, 1e-12)
        eng.iterate(5)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.iterate(a.iters)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / a.iters
        eng.iterate(1, elbo_last=True)
        elbo = eng.elbo()
        p = eng.params()
        vec = np.concatenate([p["gamma_shp"].ravel(), p["gamma_rte"].ravel(), p["phi_shp"].ravel(), p["phi_rte"].ravel(),
                              [p["nu_shp"]]])
        rho_u = eng.rho_u32.clone()
        if ref is None:
            ref = (vec, rho_u)
            d = dr = 0.0
        else:
            d = float(np.max(np.abs(vec - ref[0]) / np.maximum(np.abs(ref[0]), 1e-300)))
            dr = float((rho_u - ref[1]).abs().max())
        print("%-12s %.4f ms/iter  elbo=%.9f  max rel diff params vs first=%.3e  max abs diff rho_u=%.3e"
              % (name, ms, elbo, d, dr), flush=True)


if __name__ == "__main__":
    main()
