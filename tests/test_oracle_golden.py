"""The numpy oracle (oracle/cavi_numpy.py) against the golden vectors produced by the unmodified
reference: this is what pins the oracle (CPU only)."""
import numpy as np
import pytest

from oracle.cavi_numpy import OracleCAVI
from tests.golden_util import ALL_FIXTURES, Golden


def make_oracle(g):
    o = OracleCAVI(g.L, g.N, g.M, g.K, g.X_subs, g.X_vals, g.R_spec, mutuality=g.mutuality,
                   EPS=g.fit_kwargs.get("EPS", 1e-12), **g.priors())
    st = g.init_state()
    pr = o.default_pr_rho(st["pr_ties"], st["pr_vals"])
    o.set_state(st["gamma_shp"], st["gamma_rte"], st["phi_shp"], st["phi_rte"], st["nu_shp"], pr)
    return o


@pytest.mark.parametrize("name", ALL_FIXTURES)
def test_oracle_matches_reference_trajectory(name):
    g = Golden(name)
    o = make_oracle(g)
    z = g.z
    for it in range(g.n_iter):
        o.iterate()
        np.testing.assert_allclose(o.gamma_shp, z["it_gamma_shp"][it], rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(o.gamma_rte, z["it_gamma_rte"][it], rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(o.phi_shp, z["it_phi_shp"][it], rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(o.phi_rte, z["it_phi_rte"][it], rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(o.nu_shp, z["it_nu_shp"][it], rtol=1e-10)
        np.testing.assert_allclose(o.elbo(), z["it_elbo"][it], rtol=1e-10)
    if "rho_final" in z.files:
        np.testing.assert_allclose(o.rho, z["rho_final"], rtol=1e-9, atol=1e-300)
    else:
        t = z["rho_final_ties"]
        np.testing.assert_allclose(o.rho[t[:, 0], t[:, 1], t[:, 2]], z["rho_final_vals"], rtol=1e-9)
        np.testing.assert_allclose(o.rho.sum(axis=1), z["rho_final_colsum"], rtol=1e-9)
    assert int(np.argmax(o.rho, axis=-1).sum()) == int(z["rho_argmax_sum"])


def test_oracle_convergence_rule_sbm_k3():
    """Default convergence rule (model.py:1021-1056): the reference stopped after 40 iterations."""
    g = Golden("sbm_k3")
    o = make_oracle(g)
    elbo, trace = o.run(max_iter=g.fit_kwargs["max_iter"])
    ref = g.z["trace"]
    assert [t[0] for t in trace] == [int(v) for v in ref[:, 1]]
    np.testing.assert_allclose([t[1] for t in trace], ref[:, 2], rtol=1e-10)
    assert [bool(t[2]) for t in trace] == [bool(v) for v in ref[:, 3]]
    np.testing.assert_allclose(elbo, float(g.z["maxL"]), rtol=1e-10)
