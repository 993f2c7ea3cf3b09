"""Parity of the CUDA path (through the C ABI) with the reference's golden vectors and the CPU oracle.

Tolerances are the ones BASELINE.json's north_star states: per-iteration theta/lambda/eta within rtol 1e-5,
ELBO trajectory within rtol 1e-6, rho-argmax adjacency identical except where the top posteriors differ by < 1e-6.
"""
import warnings

import numpy as np
import pytest

from tests.golden_util import ALL_FIXTURES, LARGE_FIXTURES, Golden

pytestmark = pytest.mark.gpu

RTOL_PARAM = 1e-5
RTOL_ELBO = 1e-6


def _cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch


def build_inputs(g, structured=False):
    import vimure_b200 as vm

    X = vm.sptensor.sptensor(tuple(g.X_subs), g.X_vals, shape=(g.L, g.N, g.N, g.M))
    spec = g.R_spec
    if spec["kind"] == "all":
        R = np.ones((g.L, g.N, g.N, g.M))
    elif structured and spec["kind"] == "ego":
        R = vm.masks.EgoMask(g.L, g.N, g.M, rep=spec["rep"], diag=spec["diag"])
    else:
        subs, vals = g.R_coo()
        R = vm.sptensor.sptensor(tuple(subs), vals, shape=(g.L, g.N, g.N, g.M))
        if spec.get("dense_input") and spec["kind"] == "coo":
            R = R.toarray()
    return X, R


def make_engine(g, row0=0, nloc=None, structured=True, tile_h=64):
    torch = _cuda()
    import vimure_b200 as vm
    from vimure_b200 import _packing
    from vimure_b200._engine import CaviEngine

    X, R = build_inputs(g, structured=structured)
    mask = vm.masks.from_input(R, g.L, g.N, g.M)
    P = _packing.pack(g.X_subs, g.X_vals, g.L, g.N, g.M, g.K, mask, "cuda", row0=row0, nloc=nloc, tile_h=tile_h,
                      mutuality=g.mutuality)
    eps = g.fit_kwargs.get("EPS", 1e-12)
    eng = CaviEngine(P, g.priors(), mutuality=g.mutuality, eps=eps)
    st = g.init_state()
    # prior of the special ties from the injected (l,i,j) -> values
    flat = P.t["u_gflat"].cpu().numpy()
    pr_u = np.zeros((P.U, g.K))
    pr_u[:, 0] = 1.0
    ties = st["pr_ties"]
    if len(ties):
        tf = (ties[:, 0] * g.N + ties[:, 1]) * g.N + ties[:, 2]
        o = np.argsort(tf)
        pos = np.minimum(np.searchsorted(tf[o], flat), len(tf) - 1)
        hit = tf[o][pos] == flat
        pr_u[hit] = st["pr_vals"][o][pos[hit]]
    nu_rte = g.priors()["beta_eta"] + g.X_vals.sum() if g.mutuality else 1.0
    nu_shp = st["nu_shp"] if g.mutuality else 1e-6
    eng.set_state(st["gamma_shp"], st["gamma_rte"], st["phi_shp"], st["phi_rte"], nu_shp, nu_rte, pr_u, eps)
    return eng, P


@pytest.mark.parametrize("name", ALL_FIXTURES)
def test_per_iteration_parity_with_reference(name):
    g = Golden(name)
    eng, P = make_engine(g)
    z = g.z
    for it in range(g.n_iter):
        eng.iterate(1, elbo_last=True)
        p = eng.params()
        np.testing.assert_allclose(p["gamma_shp"], z["it_gamma_shp"][it], rtol=RTOL_PARAM, err_msg=f"gamma_shp it{it}")
        np.testing.assert_allclose(p["gamma_rte"], z["it_gamma_rte"][it], rtol=RTOL_PARAM, err_msg=f"gamma_rte it{it}")
        np.testing.assert_allclose(p["phi_shp"], z["it_phi_shp"][it], rtol=RTOL_PARAM, err_msg=f"phi_shp it{it}")
        np.testing.assert_allclose(p["phi_rte"], z["it_phi_rte"][it], rtol=RTOL_PARAM, err_msg=f"phi_rte it{it}")
        if g.mutuality:
            np.testing.assert_allclose(p["nu_shp"], z["it_nu_shp"][it], rtol=RTOL_PARAM, err_msg=f"nu_shp it{it}")
        np.testing.assert_allclose(eng.elbo(), z["it_elbo"][it], rtol=RTOL_ELBO, err_msg=f"elbo it{it}")
    rho = eng.rho_slab().cpu().numpy().astype(np.float64)
    if "rho_final" in z.files:
        ref = z["rho_final"]
        np.testing.assert_allclose(rho, ref, rtol=2e-5, atol=1e-30)
        top = np.sort(ref, axis=-1)
        clear = (top[..., -1] - top[..., -2]) >= 1e-6
        assert np.array_equal(np.argmax(rho, -1)[clear], np.argmax(ref, -1)[clear])
    else:
        t = z["rho_final_ties"]
        np.testing.assert_allclose(rho[t[:, 0], t[:, 1], t[:, 2]], z["rho_final_vals"], rtol=2e-5, atol=1e-30)
        np.testing.assert_allclose(rho.sum(axis=1), z["rho_final_colsum"], rtol=1e-5)
        np.testing.assert_allclose(rho.sum(axis=2), z["rho_final_rowsum"], rtol=1e-5)
        assert int(np.argmax(rho, -1).sum()) == int(z["rho_argmax_sum"])


@pytest.mark.parametrize("name", LARGE_FIXTURES)
@pytest.mark.parametrize("tile_h", [128, 32])
def test_shortcut_iterations_parity_with_reference(name, tile_h):
    """The iterations WITHOUT ELBO are the ones the production loop runs 9 times out of 10: there the fast dense kernel
    evaluates the SIMPLE and SINGLE special ties itself in fp32 (vm_ctx.simple_mode).  Goldens with N >= 512 reach that
    path: parameters after every such iteration against the unmodified reference (rtol 1e-5), the ELBO on the reference's
    cadence (iteration 1, 10, last) within 1e-6, final rho / argmax as in the per-iteration test."""
    g = Golden(name)
    eng, P = make_engine(g, tile_h=tile_h)
    assert eng.simple_mode and int(P.t["u_single"].sum()) > 0 and int(P.t["u_simple"].sum()) > 0
    z = g.z
    K = g.K
    for it in range(g.n_iter):
        elbo_it = it == 0 or (it + 1) % 10 == 0 or it == g.n_iter - 1
        eng.iterate(1, elbo_last=elbo_it)
        if it == 1:  # the layers did take the shortcut (flag written by k_phi_finish)
            lc = eng.layer_consts.cpu().numpy().reshape(g.L, 3 * K + 5)
            assert (lc[:, 2 * K + 4] == 1.0).all()
        p = eng.params()
        for k in ("gamma_shp", "gamma_rte", "phi_shp", "phi_rte"):
            np.testing.assert_allclose(p[k], z["it_" + k][it], rtol=RTOL_PARAM, err_msg=f"{k} it{it}")
        np.testing.assert_allclose(p["nu_shp"], z["it_nu_shp"][it], rtol=RTOL_PARAM, err_msg=f"nu_shp it{it}")
        if elbo_it:
            np.testing.assert_allclose(eng.elbo(), z["it_elbo"][it], rtol=RTOL_ELBO, err_msg=f"elbo it{it}")
    rho = eng.rho_slab().cpu().numpy().astype(np.float64)
    t = z["rho_final_ties"]
    np.testing.assert_allclose(rho[t[:, 0], t[:, 1], t[:, 2]], z["rho_final_vals"], rtol=2e-5, atol=1e-30)
    np.testing.assert_allclose(rho.sum(axis=1), z["rho_final_colsum"], rtol=1e-5)
    np.testing.assert_allclose(rho.sum(axis=2), z["rho_final_rowsum"], rtol=1e-5)
    assert int(np.argmax(rho, -1).sum()) == int(z["rho_argmax_sum"])


@pytest.mark.parametrize("name", ["f1_over", "karnataka_vil1", "gm_l2_k3", "dense_reporting"])
def test_sharded_phases_match_single_rank(name):
    """Emulate 3 ranks on one GPU: three row-block engines, the three statistics vectors summed by hand
    (what the NCCL all-reduce does), must reproduce the reference trajectory too."""
    torch = _cuda()
    from vimure_b200.model import shard_rows

    g = Golden(name)
    W = 3
    engs = [make_engine(g, *shard_rows(g.N, W, r), tile_h=16)[0] for r in range(W)]
    # make_engine ran the initial statistics per shard without a reduction: redo the reduction by hand
    F = engs[0].C

    def allreduce(attr):
        tot = sum(getattr(e, attr) for e in engs)
        for e in engs:
            getattr(e, attr).copy_(tot)

    allreduce("red3")
    z = g.z
    for it in range(min(g.n_iter, 6)):
        for e in engs:
            e.phase("gamma")
        allreduce("red1")
        for e in engs:
            e.phase("phi")
        allreduce("red2")
        for e in engs:
            e.phase("rho", F["VM_F_ELBO"])
        allreduce("red3")
        for e in engs:
            e.phase("finish", F["VM_F_ELBO"])
        for e in engs:
            p = e.params()
            np.testing.assert_allclose(p["gamma_shp"], z["it_gamma_shp"][it], rtol=RTOL_PARAM)
            np.testing.assert_allclose(p["gamma_rte"], z["it_gamma_rte"][it], rtol=RTOL_PARAM)
            np.testing.assert_allclose(p["phi_rte"], z["it_phi_rte"][it], rtol=RTOL_PARAM)
            np.testing.assert_allclose(e.elbo(), z["it_elbo"][it], rtol=RTOL_ELBO)


def test_fit_api_reproduces_reference_f1_and_trace():
    """The reference's only known answer (test_model.py:117-188): F1 0.92 +- 0.01, with the reference's own
    seeded initialisation (init='reference' consumes the same RNG stream), 2 realisations x 21 iterations."""
    _cuda()
    from sklearn.metrics import f1_score

    import vimure_b200 as vm

    for name in ("f1_over", "f1_under"):
        g = Golden(name)
        X, R = build_inputs(g)
        model = vm.VimureModel(mutuality=True)
        fk = dict(g.fit_kwargs)
        fk["num_realisations"] = 2
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            model.fit(X, R=R, init="reference", **fk)
        ref = g.z["ref2_trace"]  # realisation, seed, iter, elbo, reached
        tr = model.trace
        assert list(tr["iter"]) == [int(v) for v in ref[:, 2]]
        assert list(tr["seed"]) == [int(v) for v in ref[:, 1]]
        np.testing.assert_allclose(tr["elbo"].to_numpy(), ref[:, 3], rtol=RTOL_ELBO)
        np.testing.assert_allclose(model.maxL, float(g.z["ref2_maxL"]), rtol=RTOL_ELBO)
        np.testing.assert_allclose(model.nu_shp_f, float(g.z["ref2_nu_shp_f"]), rtol=RTOL_PARAM)
        np.testing.assert_allclose(model.gamma_shp_f, g.z["ref2_gamma_shp_f"], rtol=RTOL_PARAM)
        Y_rec = vm.utils.apply_rho_threshold(model, threshold=0.5)[0].flatten()
        f1 = f1_score(g.z["Y_true"].flatten(), Y_rec)
        assert abs(f1 - float(g.z["ref_f1"])) < 1e-9
        assert abs(f1 - (0.92 if name == "f1_over" else 0.97)) < 1e-2
        assert type(model.nu_shp) == np.float64 and type(model.nu_rte) == np.float64
        assert model.rho.shape == (g.L, g.N, g.N, g.K)


def test_fit_convergence_rule_matches_reference():
    """sbm_k3 was fitted by the reference with the default stop rule: it stopped after 40 iterations."""
    _cuda()
    import vimure_b200 as vm

    g = Golden("sbm_k3")
    X, R = build_inputs(g)
    model = vm.VimureModel()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model.fit(X, R=R, init_state=g.init_state(), **g.fit_kwargs)
    ref = g.z["trace"]
    assert list(model.trace["iter"]) == [int(v) for v in ref[:, 1]]
    np.testing.assert_allclose(model.trace["elbo"].to_numpy(), ref[:, 2], rtol=RTOL_ELBO)
    assert [bool(v) for v in model.trace["reached_convergence"]] == [bool(v) for v in ref[:, 3]]
    assert model.n_iter_ == g.n_iter


def test_device_special_functions():
    torch = _cuda()
    import ctypes

    import scipy.special as sp

    from vimure_b200 import _capi

    lib = _capi.load()
    x = np.concatenate([np.logspace(-6, 6, 4000), np.linspace(0.05, 30, 3000)])
    xt = torch.as_tensor(x, device="cuda")
    dg, lg = torch.empty_like(xt), torch.empty_like(xt)
    rc = lib.vm_test_special(ctypes.c_void_p(xt.data_ptr()), ctypes.c_void_p(dg.data_ptr()),
                             ctypes.c_void_p(lg.data_ptr()), x.size, None)
    assert rc == 0
    torch.cuda.synchronize()
    np.testing.assert_allclose(dg.cpu().numpy(), sp.psi(x), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(lg.cpu().numpy(), sp.gammaln(x), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("name", ["gm_l2_k3", "dense_reporting", "custom_mask", "sbm_n520"])
def test_two_engines_are_bit_identical(name):
    """Every reduction has a fixed order and the per-reporter corrections are integer atomics: two engines over the same
    packed problem (different tile heights for the small ones) give the same bits, slab included."""
    g = Golden(name)
    a, _ = make_engine(g, tile_h=8 if g.N < 512 else 128)
    b, _ = make_engine(g, tile_h=8 if g.N < 512 else 128)
    for it in range(4):
        a.iterate(1, elbo_last=(it % 2 == 1))
        b.iterate(1, elbo_last=(it % 2 == 1))
        pa, pb = a.params(), b.params()
        for k in ("gamma_shp", "gamma_rte", "phi_shp", "phi_rte", "nu_shp"):
            np.testing.assert_array_equal(pa[k], pb[k])
    assert a.elbo() == b.elbo()
    import torch

    assert torch.equal(a.rho_slab(), b.rho_slab())


@pytest.mark.parametrize("name", ["undirected", "rho_prior", "nomut", "gm_l2_k3"])
def test_fit_with_reference_seeded_init_reproduces_reference_trace(name):
    """`init="reference"` consumes the reference's RNG stream (model.py:470-500, 570-592): a seeded fit reproduces the
    reference's ELBO trace without any injected state -- incl. the symmetrised prior of undirected=True and a user rho_prior."""
    _cuda()
    import vimure_b200 as vm

    g = Golden(name)
    X, R = build_inputs(g)
    model = vm.VimureModel(**g.model_kwargs)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model.fit(X, R=R, init="reference", **g.fit_kwargs)
    ref = g.z["trace"]  # realisation, iter, elbo, reached
    assert list(model.trace["iter"]) == [int(v) for v in ref[:, 1]]
    np.testing.assert_allclose(model.trace["elbo"].to_numpy(), ref[:, 2], rtol=RTOL_ELBO)
    np.testing.assert_allclose(model.maxL, float(g.z["maxL"]), rtol=RTOL_ELBO)
    np.testing.assert_allclose(model.gamma_shp, g.z["it_gamma_shp"][-1], rtol=RTOL_PARAM)
    np.testing.assert_allclose(model.pr_rho.reshape(-1, g.K)[
        (g.z["init_pr_ties"][:, 0] * g.N + g.z["init_pr_ties"][:, 1]) * g.N + g.z["init_pr_ties"][:, 2]],
        g.z["init_pr_vals"], rtol=1e-12)
