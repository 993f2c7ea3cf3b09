"""Edge cases of the path against the CPU oracle: completely underflowed rows (Q3), a large EPS (closed form no longer
negligible), a layer without any report, reporters that are a subset of the nodes (M < N), K = 12 (a general-kernel-only K; the default K = max(X)+1 of count data lands there)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def _run(L, N, M, K, subs, vals, mask, spec, state, priors, eps=1e-12, mutuality=True, iters=4, rtol_p=1e-5, rtol_e=1e-6,
         tile_h=8):
    _cuda()
    from oracle.cavi_numpy import OracleCAVI
    from vimure_b200 import _packing
    from vimure_b200._engine import CaviEngine

    P = _packing.pack(subs, vals, L, N, M, K, mask, "cuda", tile_h=tile_h, mutuality=mutuality)
    eng = CaviEngine(P, priors, mutuality=mutuality, eps=eps)
    keep = (P.t["u_has_x"] & P.t["u_reported"]).cpu().numpy()
    prng = np.random.RandomState(5)
    pr_u = np.zeros((P.U, K))
    pr_u[:, 0] = 1.0
    pr = 1 + 0.01 * prng.random_sample((int(keep.sum()), K))
    pr_u[keep] = pr / pr.sum(axis=1)[:, None]
    nu_rte = priors["beta_eta"] + float(np.sum(vals))
    eng.set_state(state["gamma_shp"], state["gamma_rte"], state["phi_shp"], state["phi_rte"],
                  state["nu_shp"] if mutuality else 1e-6, nu_rte if mutuality else 1.0, pr_u, eps)
    o = OracleCAVI(L, N, M, K, subs, vals, spec, mutuality=mutuality, EPS=eps, **priors)
    flat = P.t["u_gflat"].cpu().numpy()[keep]
    ties = np.stack([flat // (N * N), (flat // N) % N, flat % N], axis=1)
    o.set_state(state["gamma_shp"], state["gamma_rte"], state["phi_shp"], state["phi_rte"], state["nu_shp"],
                o.default_pr_rho(ties, pr_u[keep]))
    for it in range(iters):
        eng.iterate(1, elbo_last=True)
        o.iterate()
        p = eng.params()
        for k in ("gamma_shp", "gamma_rte", "phi_shp", "phi_rte"):
            np.testing.assert_allclose(p[k], getattr(o, k), rtol=rtol_p, err_msg=f"{k} it{it}")
        if mutuality:
            np.testing.assert_allclose(p["nu_shp"], o.nu_shp, rtol=rtol_p, err_msg=f"nu it{it}")
        np.testing.assert_allclose(eng.elbo(), o.elbo(), rtol=rtol_e, err_msg=f"elbo it{it}")
    rho = eng.rho_slab().cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(rho, o.rho, rtol=3e-5, atol=1e-30)
    return eng, o


def _net(N, M, L, K, seed, eta=0.4, law="sbm"):
    import vimure_b200.synthetic as syn

    mk = syn.StandardSBM if law == "sbm" else syn.Multitensor
    kw = dict(N=N, M=M, L=L, K=K, C=2, avg_degree=6, seed=seed)
    if law != "sbm":
        kw["eta"] = eta
    return mk(**kw).build_X(mutuality=eta, seed=seed + 1)


def _state(L, M, K, seed, th=(0.1, 0.1), lam=(10.0, 10.0)):
    rs = np.random.RandomState(seed).random_sample
    return dict(gamma_shp=th[0] * rs((L, M)) + th[0], gamma_rte=th[1] * rs((L, M)) + th[1],
                phi_shp=lam[0] * rs((L, K)) + lam[0], phi_rte=lam[1] * rs((L, K)) + lam[1], nu_shp=0.5 * rs(1)[0] + 0.5)


PRI = dict(alpha_theta=0.1, beta_theta=0.1, alpha_lambda=10.0, beta_lambda=10.0, alpha_eta=0.5, beta_eta=1.0)


def test_completely_underflowed_rows_ego():
    """Huge E[theta]: S * E[lambda_0] > 745 for most ties -> the reference's exp() gives 0 for every k and the row
    stays all-zero (model.py:807-811, Q3).  Statistics, ELBO and the slab must follow."""
    net = _net(48, 48, 1, 2, seed=3)
    st = _state(1, 48, 2, 1)
    # gamma_rte ~ 2N E[lambda_0] after the first update, so S E[lambda_0] ~ 2 alpha_theta / (2N): alpha_theta >> 745 N
    # and a lambda prior strong enough to keep E[lambda] ~ 1 (otherwise the model renormalises lambda away)
    # half of the reporters only: ties between two ordinary reporters stay alive
    at = np.full((1, 48), 0.1)
    at[0, :24] = 2e5
    st["gamma_shp"] = at * (1 + st["gamma_shp"])
    st["phi_shp"] = st["phi_shp"] * 0 + 1e12
    st["phi_rte"] = st["phi_rte"] * 0 + 1e12
    pri = dict(PRI, alpha_theta=at, beta_theta=np.full((1, 48), 0.1), alpha_lambda=1e12, beta_lambda=1e12)
    spec = {"kind": "ego", "rep": np.ones((1, 48), dtype=np.uint8), "diag": True}
    eng, o = _run(1, 48, 48, 2, np.stack(net.X.subs), net.X.vals, net.R, spec, st, pri, iters=3)
    dead = o.rho.sum(-1) == 0
    assert dead.any() and not dead.all(), "scenario must contain dead AND live rows"


def test_completely_underflowed_rows_all_mask():
    import vimure_b200 as vm
    import vimure_b200.synthetic as syn

    y = syn.StandardSBM(N=40, M=12, L=2, K=2, C=2, avg_degree=5, seed=2)
    X, _ = syn.dense_reporting_X(y, M=12, mutuality=0.3, seed=4)
    st = _state(2, 12, 2, 2)
    st["gamma_shp"] = st["gamma_shp"] * 0 + 5e5  # S E[lambda_0] ~ M alpha_theta / N^2 >> 745
    st["phi_shp"] = st["phi_shp"] * 0 + 1e12
    st["phi_rte"] = st["phi_rte"] * 0 + 1e12
    pri = dict(PRI, alpha_theta=5e5, alpha_lambda=1e12, beta_lambda=1e12)
    eng, o = _run(2, 40, 12, 2, np.stack(X.subs), X.vals, vm.masks.AllMask(2, 40, 12), {"kind": "all", "dense_input": True},
                  st, pri, iters=3)
    assert (o.rho.sum(-1) == 0).any()


def test_large_eps_closed_form_not_negligible():
    net = _net(64, 64, 2, 3, seed=7, law="gm")
    spec = {"kind": "ego", "rep": np.ones((2, 64), dtype=np.uint8), "diag": True}
    _run(2, 64, 64, 3, np.stack(net.X.subs), net.X.vals, net.R, spec, _state(2, 64, 3, 3), PRI, eps=0.02, iters=4,
         rtol_p=2e-5, rtol_e=2e-6)


def test_layer_without_reports_and_subset_of_reporters():
    import vimure_b200 as vm

    net = _net(56, 56, 1, 2, seed=11)
    s = np.stack(net.X.subs)
    rep = np.zeros((2, 40), dtype=np.uint8)
    rep[:, ::2] = 1  # only every other node among the first 40 is a reporter (M = 40 < N = 56)
    keep = (s[3] < 40) & (rep[0][np.minimum(s[3], 39)] == 1)
    subs, vals = s[:, keep].copy(), net.X.vals[keep]
    # layer 1 exists (L = 2) but carries no report at all
    mask = vm.masks.EgoMask(2, 56, 40, rep=rep, diag=False)
    spec = {"kind": "ego", "rep": rep, "diag": False}
    _run(2, 56, 40, 2, subs, vals, mask, spec, _state(2, 40, 2, 4), PRI, iters=4)


def test_k12_large_k_build():
    net = _net(40, 40, 1, 12, seed=21, law="gm", eta=0.5)
    spec = {"kind": "ego", "rep": np.ones((1, 40), dtype=np.uint8), "diag": True}
    _run(1, 40, 40, 12, np.stack(net.X.subs), net.X.vals, net.R, spec, _state(1, 40, 12, 6), PRI, iters=3)
