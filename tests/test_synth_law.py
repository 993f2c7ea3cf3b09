"""The LAW of this repo's sparse report generators against the reference generator's moments
(tests/golden/synth_moments.npz, made by oracle/gen_synth_moments.py from `vimure.synthetic._build_X`,
reference synthetic.py:138-209): same ground truth Y, same reliabilities theta, independent seeds.  Their RNG streams
differ from numpy's by design, so what is compared are per-class sums -- each a sum over thousands of independent
(reporter, pair) draws -- within 5 standard errors."""
import os

import numpy as np
import pytest

from oracle.gen_synth_moments import class_stats

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "synth_moments.npz")


def _net_from_fixture(z):
    import vimure_b200.synthetic as syn
    from vimure_b200.sptensor import sptensor

    L, N, M, K = (int(v) for v in z["dims"])
    net = syn.StandardSBM(N=N, M=M, L=L, K=K, C=2, avg_degree=8, seed=3)
    net.Y_subs, net.Y_vals = z["Y_subs"].astype(np.int64), z["Y_vals"].astype(np.int64)
    net.Y = sptensor(tuple(net.Y_subs), net.Y_vals, shape=(L, N, N))
    return net


def _dense(subs, vals, shape):
    X = np.zeros(shape, dtype=np.int64)
    X[tuple(subs)] = vals
    return X


def _compare(stats, ref):
    """stats, ref: (seeds, 5 classes, 5 statistics)."""
    assert np.array_equal(stats[:, :, 0], np.broadcast_to(ref[0, :, 0], stats[:, :, 0].shape))  # slot counts are exact
    for c in range(5):
        for q, name in ((1, "sum X"), (2, "#X>0"), (3, "sum X^2"), (4, "sum X X^T")):
            a, b = stats[:, c, q], ref[:, c, q]
            se = np.sqrt(a.var(ddof=1) / len(a) + b.var(ddof=1) / len(b))
            assert abs(a.mean() - b.mean()) <= 5 * se + 1e-9, (c, name, a.mean(), b.mean(), se)


def test_host_generator_follows_the_reference_law():
    z = np.load(FIX)
    net = _net_from_fixture(z)
    Yd = net.Y.toarray()
    stats = []
    for sd in range(40):
        net.build_X(mutuality=float(z["eta"]), theta=z["theta"], seed=1000 + sd)
        stats.append(class_stats(_dense(net.X.subs, net.X.vals, net.X.shape), Yd))
    _compare(np.stack(stats), z["stats"])


@pytest.mark.gpu
def test_device_generator_follows_the_reference_law():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    z = np.load(FIX)
    net = _net_from_fixture(z)
    Yd = net.Y.toarray()
    L, N, M, K = (int(v) for v in z["dims"])
    stats = []
    for sd in range(40):
        subs, vals = net.build_X_device(mutuality=float(z["eta"]), theta=z["theta"], seed=2000 + sd)
        stats.append(class_stats(_dense(subs.cpu().numpy(), vals.cpu().numpy(), (L, N, N, M)), Yd))
    _compare(np.stack(stats), z["stats"])


@pytest.mark.gpu
def test_device_generator_shards_agree_with_the_whole():
    """Counter-based RNG: the entries a rank generates for its row block are the same ones a single rank generates for
    the whole network; with emit_transposed it also gets exactly the reciprocal entries of its rows that it does not own."""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import vimure_b200.synthetic as syn
    from vimure_b200.model import shard_rows

    L, N, K = 2, 700, 3
    net = syn.Multitensor(N=N, L=L, K=K, C=2, avg_degree=8, eta=0.5, seed=4)

    def key(subs):
        s = subs.cpu().numpy().astype(np.int64)
        return ((s[0] * N + s[1]) * N + s[2]) * N + s[3]

    subs, vals = net.build_X_device(mutuality=0.5, seed=9)
    k_all = key(subs)
    o = np.argsort(k_all)
    k_all, v_all = k_all[o], vals.cpu().numpy()[o]
    assert len(np.unique(k_all)) == len(k_all) and (v_all > 0).all()
    # twice the same call: same set
    subs2, vals2 = net.build_X_device(mutuality=0.5, seed=9)
    k2 = key(subs2)
    o2 = np.argsort(k2)
    assert np.array_equal(k2[o2], k_all) and np.array_equal(vals2.cpu().numpy()[o2], v_all)
    i_all, j_all = (k_all // (N * N)) % N, (k_all // N) % N
    for W in (3,):
        for r in range(W):
            row0, nloc = shard_rows(N, W, r)
            s, v = net.build_X_device(mutuality=0.5, seed=9, row0=row0, nloc=nloc, emit_transposed=True)
            ks = key(s)
            os_ = np.argsort(ks)
            own = (i_all >= row0) & (i_all < row0 + nloc)
            tr = ~own & (j_all >= row0) & (j_all < row0 + nloc)
            want = own | tr
            assert np.array_equal(ks[os_], k_all[want]) and np.array_equal(v.cpu().numpy()[os_], v_all[want])
            s, v = net.build_X_device(mutuality=0.5, seed=9, row0=row0, nloc=nloc, emit_transposed=False)
            ks = key(s)
            os_ = np.argsort(ks)
            assert np.array_equal(ks[os_], k_all[own]) and np.array_equal(v.cpu().numpy()[os_], v_all[own])


@pytest.mark.gpu
def test_posterior_synthetic_network_from_a_fitted_model():
    """`PosteriorSyntheticNetwork` (reference synthetic.py:964-1177): Y from rho_f, reports from the posterior Gammas."""
    import warnings

    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import vimure_b200 as vm
    import vimure_b200.synthetic as syn

    net = syn.Multitensor(N=120, L=2, K=2, C=2, avg_degree=8, eta=0.3, seed=3).build_X(mutuality=0.3, seed=4)
    model = vm.VimureModel(mutuality=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model.fit(net.X, R=net.R, K=2, seed=5, max_iter=21)
    ps = syn.PosteriorSyntheticNetwork(model, seed_Y=11).build_Y()
    # numpy stream for a small problem: exactly the reference's draw
    pv = model.rho_f / model.rho_f.sum(axis=-1, keepdims=True)
    ref_Y = np.random.default_rng(11).multinomial(n=1, pvals=pv, size=(2, 120, 120)).argmax(axis=-1)
    assert np.array_equal(ps.Y.toarray(), ref_Y)
    ps.build_X(seed_X=7)
    prng = np.random.RandomState(7)
    np.testing.assert_allclose(ps.theta, prng.gamma(shape=model.gamma_shp_f, scale=1.0 / model.gamma_rte_f, size=(2, 120)))
    assert ps.X.shape == (2, 120, 120, 120) and len(ps.X.vals) > 0 and ps.X.vals.min() >= 1
    l, i, j, m = ps.X.subs
    assert np.all((m == i) | (m == j))  # self-reporter mask
    # reporters see their true ties far more often than their non-ties
    Yd = ps.Y.toarray() > 0
    hit = Yd[l, i, j].mean()
    assert hit > 5 * Yd.mean()
    assert ps.X_union.shape == (2, 120, 120) and len(ps.X_intersection.vals) <= len(ps.X_union.vals)
    # the device stream for Y agrees in law: same marginal frequency of ties
    psd = syn.PosteriorSyntheticNetwork(model, seed_Y=11).build_Y(rng="device")
    assert abs(len(psd.Y_vals) - len(ps.Y_vals)) < 6 * np.sqrt(len(ps.Y_vals)) + 10
