"""Runs the UNMODIFIED reference (`/root/reference/src/python/vimure`) under the `oracle/shims`
stand-ins and records, for one realisation, the injected initial state and the state after every
CAVI iteration.  TEST INFRASTRUCTURE ONLY: it exists to pin the numpy restatement
(`oracle/cavi_numpy.py`) and to generate the golden vectors under `tests/golden/`.

`/root/reference` exists only in the build container; nothing in `tests -m gpu`, `smoke()` or
`bench.py` may import this module.
"""
import os
import sys
import warnings

import numpy as np

REFERENCE_SRC = "/root/reference/src/python"
SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def import_reference():
    """Import the reference `vimure` package (needs /root/reference)."""
    if not os.path.isdir(REFERENCE_SRC):
        raise RuntimeError("reference sources are not available on this machine")
    for p in (SHIMS, REFERENCE_SRC):
        if p not in sys.path:
            sys.path.insert(0, p)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import vimure as vm  # noqa: E402
        import vimure.utils  # noqa: F401
        import vimure._io  # noqa: F401
    return vm


def run_reference_trace(X, fit_kwargs, model_kwargs=None, record_rho=True):
    """Fit the reference for ONE realisation and record per-iteration state.

    Returns a dict with
      init   : gamma_shp/gamma_rte/phi_shp/phi_rte/nu_shp/nu_rte and pr_rho (dense) as drawn by
               the reference (`model.py:458-617`),
      iters  : per-iteration arrays gamma_shp, gamma_rte, phi_shp, phi_rte, nu_shp (after
               `_update_CAVI`, `model.py:623-660`) and the ELBO evaluated at EVERY iteration with the
               reference's own `__ELBO` (`model.py:948-1019`; it is side-effect free),
      model  : the fitted reference model object.
    """
    vm = import_reference()
    model_kwargs = dict(model_kwargs or {})
    fit_kwargs = dict(fit_kwargs)
    fit_kwargs["num_realisations"] = 1

    rec = {"init": {}, "iters": []}

    class Recorder(vm.model.VimureModel):
        def _initialize_old_variables(self):
            super()._initialize_old_variables()
            rec["init"] = dict(
                gamma_shp=np.array(self.gamma_shp, dtype=float),
                gamma_rte=np.array(self.gamma_rte, dtype=float),
                phi_shp=np.array(self.phi_shp, dtype=float),
                phi_rte=np.array(self.phi_rte, dtype=float),
                nu_shp=float(self.nu_shp),
                nu_rte=float(self.nu_rte),
                pr_rho=np.array(self.pr_rho, dtype=float),
            )

        def _update_CAVI(self, data, subs_nz, data_T_vals=None):
            out = super()._update_CAVI(data, subs_nz, data_T_vals)
            elbo = self._VimureModel__ELBO(self.X, self.data_T, self.subs_nz)
            rec["iters"].append(
                dict(
                    gamma_shp=np.array(self.gamma_shp, dtype=float),
                    gamma_rte=np.array(self.gamma_rte, dtype=float),
                    phi_shp=np.array(self.phi_shp, dtype=float),
                    phi_rte=np.array(self.phi_rte, dtype=float),
                    nu_shp=float(self.nu_shp),
                    elbo=float(elbo),
                    rho=np.array(self.rho, dtype=float) if record_rho else None,
                )
            )
            return out

    # the reference's constructor is keyword-only
    model = Recorder(**model_kwargs)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model.fit(X, **fit_kwargs)
    rec["model"] = model
    return rec
