"""The reference's own API tests (test/test_model.py) re-expressed against vimure_b200.VimureModel (GPU)."""
import warnings

import numpy as np
import pandas as pd
import pytest

from tests.golden_util import Golden
from tests.test_gpu_parity import build_inputs

pytestmark = pytest.mark.gpu


def _cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def check_final_parameters(model, L, N, M, K):
    """reference test_model.py:15-52"""
    assert type(model.rho) == np.ndarray and model.rho.shape == (L, N, N, K) and model.rho.sum() != 0
    for nm, shape in (("gamma_shp", (L, M)), ("gamma_rte", (L, M)), ("phi_shp", (L, K)), ("phi_rte", (L, K))):
        v = getattr(model, nm)
        assert type(v) == np.ndarray and v.shape == shape and v.sum() > 0
    assert type(model.nu_shp) == np.float64 and model.nu_shp >= 0
    assert type(model.nu_rte) == np.float64 and model.nu_rte >= 0


@pytest.fixture(scope="module")
def fitted():
    _cuda()
    import vimure_b200 as vm
    import vimure_b200.synthetic as syn

    net = syn.StandardSBM(N=100, M=100, L=1, K=2, C=2, avg_degree=4, seed=3).build_X(mutuality=0.5, seed=4)
    models = {}
    for mut in (True, False):
        m = vm.VimureModel(mutuality=mut)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m.fit(net.X, K=2, R=net.R, num_realisations=1, max_iter=100, seed=1)
        models[mut] = m
    return net, models


def test_parameters_and_warnings():
    """reference test_model.py:59-115"""
    _cuda()
    import vimure_b200 as vm
    import vimure_b200.synthetic as syn

    net = syn.StandardSBM(N=20, M=20, L=1, K=3, C=2, avg_degree=2, seed=1).build_X(seed=2)
    model = vm.VimureModel()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model.fit(net.X, K=3, R=net.R, max_iter=30)
    check_final_parameters(model, 1, 20, 20, 3)
    with pytest.warns(UserWarning, match="Reporters Mask was not informed"):
        vm.VimureModel().fit(net.X, K=3, max_iter=2)
    with pytest.warns(UserWarning, match="Parameter K was None"):
        vm.VimureModel().fit(net.X, R=net.R, max_iter=2)
    with pytest.raises(ValueError, match="Dimensions of reporter mask"):
        vm.VimureModel().fit(net.X, K=3, R=np.ones((1, 20, 20, 3)), max_iter=2)
    with pytest.raises(ValueError, match="theta_prior must be a 2D tuple"):
        vm.VimureModel().fit(net.X, K=3, R=net.R, theta_prior=[0.1, 0.1])
    with pytest.raises(ValueError, match="alpha_lambda matrix is not valid"):
        vm.VimureModel().fit(net.X, K=3, R=net.R, alpha_lambda=np.ones((1, 2)), beta_lambda=np.ones((1, 3)))
    with pytest.raises(ValueError, match="rho_prior has to have shape"):
        vm.VimureModel().fit(net.X, K=3, R=net.R, rho_prior=np.ones((1, 20, 19)))
    with pytest.warns(UserWarning, match="Overriding mutuality"):
        m = vm.VimureModel(undirected=True)
    assert m.mutuality is False
    with pytest.raises(ValueError, match="has to be symmetric"):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m.fit(net.X, K=3, R=net.R, max_iter=2)


def test_inferred_model_methods(fitted):
    """reference test_model.py:362-447"""
    net, models = fitted
    for mut, model in models.items():
        with pytest.raises(ValueError, match="'method' should be one of"):
            model.get_inferred_model(method="NotImplemented")
        Y = model.get_inferred_model(method="rho_max")
        assert Y.shape == (model.L, model.N, model.N) and Y.sum() > 0
        assert np.array_equal(Y, np.argmax(model.rho_f, axis=-1))
    model = models[True]
    with pytest.raises(ValueError, match="you must set the threshold"):
        model.get_inferred_model(method="fixed_threshold")
    with pytest.raises(ValueError, match="you must set the threshold"):
        model.get_inferred_model(method="fixed_threshold", threshold=2)
    Y = model.get_inferred_model(method="fixed_threshold", threshold=0.5)
    assert Y.shape == (1, 100, 100) and Y.sum() > 0
    assert np.array_equal(Y, (model.rho_f[..., 1] >= 0.5).astype(float))
    Ym = model.get_inferred_model(method="rho_mean")
    np.testing.assert_allclose(Ym, model.rho_f[..., 1])
    Yh = model.get_inferred_model(method="heuristic_threshold")
    assert Yh.shape == (1, 100, 100)
    thr = 0.54 * model.G_exp_nu - 0.01
    assert np.array_equal(Yh, (model.rho_f[..., 1] >= thr).astype(int))
    with pytest.warns(UserWarning, match="threshold methods is incompatible"):
        Y2 = models[False].get_inferred_model(method="heuristic_threshold")
    assert np.array_equal(Y2, models[False].get_inferred_model("rho_max"))
    post = model.get_posterior_estimates()
    assert set(post) == {"nu", "theta", "lambda", "rho"}
    assert post["theta"].shape == (1, 100) and post["lambda"].shape == (1, 2) and post["rho"].shape == (1, 100, 100, 2)


def test_consumers_read_the_best_restart_whichever_slab_holds_it():
    """rho_f is the posterior of the BEST restart (model.py:925-942): the consumers must read it whether it is the last
    restart's slab, a device copy kept from an earlier restart (sequential restarts) or another engine's slab (restarts
    side by side) -- all four methods against the host copy of rho_f."""
    _cuda()
    import vimure_b200 as vm
    from tests.golden_util import Golden
    from tests.test_gpu_parity import build_inputs

    g = Golden("f1_over")
    X, R = build_inputs(g)
    for conc in (False, True):
        model = vm.VimureModel(mutuality=True)
        fk = dict(g.fit_kwargs)
        fk.update(num_realisations=3, max_iter=11, concurrent_realisations=conc, seed=3)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            model.fit(X, R=R, **fk)
        rf = model.rho_f
        assert np.array_equal(model.get_inferred_model("rho_max"), np.argmax(rf, axis=-1))
        np.testing.assert_allclose(model.get_inferred_model("rho_mean"), rf[..., 1], rtol=0, atol=0)
        assert np.array_equal(model.get_inferred_model("fixed_threshold", threshold=0.3), (rf[..., 1] >= np.float32(0.3)).astype(float))
        thr = np.float32(0.54 * model.G_exp_nu - 0.01)
        assert np.array_equal(model.get_inferred_model("heuristic_threshold"), (rf[..., 1] >= thr).astype(int))


def test_sample_inferred_model(fitted):
    """reference test_model.py:430-438"""
    net, models = fitted
    Y = models[True].sample_inferred_model(N=10)
    assert len(Y) == 10
    for y in Y:
        assert y.shape == (1, 100, 100)


def test_reference_compatible_attributes(fitted):
    net, models = fitted
    m = models[True]
    assert m.pr_rho.shape == (1, 100, 100, 2) and np.allclose(m.pr_rho.sum(-1), 1.0)
    assert m.logpr_rho.shape == m.pr_rho.shape
    assert m.data_T_vals.shape == net.X.vals.shape
    Xd = net.X.toarray()
    l, i, j, r = net.X.subs
    assert np.array_equal(m.data_T_vals, Xd[l, j, i, r])
    assert list(m.trace.columns) == ["realisation", "seed", "iter", "elbo", "runtime", "reached_convergence"]
    assert m.G_exp_theta_f.shape == (1, 100) and m.G_exp_lambda_f.shape == (1, 2)
    assert m.get_params()["mutuality"] is True  # sklearn BaseEstimator surface


def test_dataframe_input_equals_tensor_input():
    """reference test_model.py:450-481: fitting a DataFrame == fitting the parsed tensors."""
    _cuda()
    import vimure_b200 as vm
    from vimure_b200.io import read_from_edgelist

    g = Golden("f1_over")
    l, i, j, m = g.X_subs
    df = pd.DataFrame({"ego": i, "alter": j, "reporter": m, "layer": l, "weight": g.X_vals})
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = vm.VimureModel().fit(df, seed=1, num_realisations=1, max_iter=30)
        net = read_from_edgelist(df, K=2)
        b = vm.VimureModel().fit(net.X, K=net.K, R=net.R, seed=1, num_realisations=1, max_iter=30)
    assert a.K == b.K == 2
    np.testing.assert_array_equal(a.gamma_shp, b.gamma_shp)
    np.testing.assert_array_equal(a.trace["elbo"].to_numpy(), b.trace["elbo"].to_numpy())
    check_final_parameters(a, net.L, net.N, net.N, 2)


@pytest.mark.parametrize("name,kw", [("f1_over", {}), ("karnataka_vil1", {"convergence_tol": 0.1})])
def test_concurrent_realisations_match_sequential(name, kw):
    """The restarts of model.py:386-437 run side by side (one stream + one state per restart) must give exactly what
    they give one after the other: same seeds, same trace, same best restart, same posteriors."""
    _cuda()
    import vimure_b200 as vm

    g = Golden(name)
    X, R = build_inputs(g)
    fk = dict(g.fit_kwargs)
    fk["num_realisations"] = 4
    fk["max_iter"] = 35  # ELBO at 1, 10, 20, 30, 35: restarts may stop at different iterations
    fits = {}
    for conc in (False, True):
        m = vm.VimureModel(mutuality=True, **kw)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m.fit(X, R=R, init="reference", concurrent_realisations=conc, **fk)
        fits[conc] = m
    a, b = fits[False], fits[True]
    for col in ("realisation", "seed", "iter", "reached_convergence"):
        assert list(a.trace[col]) == list(b.trace[col])
    np.testing.assert_array_equal(a.trace["elbo"].to_numpy(), b.trace["elbo"].to_numpy())
    assert a.maxL == b.maxL and a.seed == b.seed and a.n_iter_ == b.n_iter_
    for nm in ("gamma_shp", "gamma_rte", "phi_shp", "phi_rte", "nu_shp", "gamma_shp_f", "gamma_rte_f", "phi_shp_f",
               "phi_rte_f", "nu_shp_f", "G_exp_theta_f", "G_exp_lambda_f", "G_exp_nu_f"):
        np.testing.assert_array_equal(np.asarray(getattr(a, nm)), np.asarray(getattr(b, nm)), err_msg=nm)
    np.testing.assert_array_equal(a.rho, b.rho)
    np.testing.assert_array_equal(a.rho_f, b.rho_f)
    np.testing.assert_array_equal(a.pr_rho, b.pr_rho)
    np.testing.assert_array_equal(a.get_inferred_model(), b.get_inferred_model())
    np.testing.assert_array_equal(a.get_inferred_model(method="fixed_threshold", threshold=0.3),
                                  b.get_inferred_model(method="fixed_threshold", threshold=0.3))
    # and the default ("auto") takes the side-by-side path for a launch-bound problem like this one
    m = vm.VimureModel(mutuality=True, **kw)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m.fit(X, R=R, init="reference", **fk)
    assert m._engine_f is not None
    np.testing.assert_array_equal(m.trace["elbo"].to_numpy(), a.trace["elbo"].to_numpy())


def test_device_sampling_follows_rho(fitted):
    """`sample_inferred_model(rng="device")` (vm_sample): reproducible, keyed by seed, and distributed like
    multinomial(n, rho).argmax(-1) of reference model.py:1086-1088."""
    net, models = fitted
    m = models[True]
    rho1 = m.rho_f[..., 1]
    a = m.sample_inferred_model(N=3, seed=7, rng="device")
    b = m.sample_inferred_model(N=3, seed=7, rng="device")
    assert len(a) == 3 and all(x.shape == (1, 100, 100) and x.dtype.kind == "i" for x in a)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)
    assert (a[0] != a[1]).any()  # seed + i: different streams
    np.testing.assert_array_equal(m.sample_inferred_model(N=2, seed=8, rng="device")[0],
                                  m._engine_of_rho_f().sample(2, 8).cpu().numpy())
    # one trial per tie: P(Y = 1) = rho_1; frequencies over 400 seeds within 6 sigma (+ slack)
    eng = m._engine_of_rho_f()
    n = 400
    freq = np.zeros(rho1.shape)
    for s in range(n):
        freq += eng.sample(1, 1000 + s).cpu().numpy()
    freq /= n
    sigma = np.sqrt(np.maximum(rho1 * (1 - rho1), 1e-12) / n)
    assert np.all(np.abs(freq - rho1) <= 6 * sigma + 5e-3), float(np.max(np.abs(freq - rho1) / (sigma + 1e-3)))
    mid = (rho1 > 0.2) & (rho1 < 0.8)
    if mid.any():  # and it does fluctuate where the posterior is undecided
        assert np.abs(freq - rho1)[mid].max() > 0
    # many trials: the most frequent category is the most probable one wherever the posterior is decided
    big = eng.sample(2001, 5).cpu().numpy()
    decided = np.abs(rho1 - 0.5) > 0.1
    np.testing.assert_array_equal(big[decided], (rho1 > 0.5)[decided].astype(big.dtype))
    # the numpy path still exists and agrees in law on the decided ties
    host = m.sample_inferred_model(N=1, seed=7, rng="numpy")[0]
    assert host.shape == (1, 100, 100)
