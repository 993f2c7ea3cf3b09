"""CPU oracle: a vectorised fp64 numpy restatement of the reference CAVI loop.

TEST INFRASTRUCTURE ONLY.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import this module; the product package never does.

Parity status: PINNED.  `tests/test_oracle_golden.py` checks this restatement against the golden
vectors in `tests/golden/*.npz`, which were produced by running the unmodified reference
(`oracle/gen_golden.py`) -- per-iteration gamma/phi/nu, the ELBO at every iteration and the final
rho agree to ~1e-12 relative.

What is restated (all `file:line` relative to `/root/reference/src/python/vimure/`):
  * `_update_cache`                       model.py:662-696
  * `_update_gamma` + `_sp_uttkrp_theta`  model.py:698-727, 832-859
  * `_update_phi` + `_sp_uttkrp_lambda`   model.py:729-761, 861-887
  * `_update_rho` + `_sp_uttkrp_rho`      model.py:763-818, 889-923
  * `_update_nu`                          model.py:820-830
  * `__ELBO`, `_calculate_mean_poisson`, `_gamma_elbo_term`, `_categorical_elbo_term`
                                          model.py:948-1019, 1220-1313
  * `_check_for_convergence`              model.py:1021-1056
including the number-changing quirks Q1-Q6 of SURVEY.md section 3.3.

The reporter mask R is given as a *spec* so that large masks never have to be materialised:
  {"kind": "ego", "rep": (L,M) 0/1, "diag": bool}   reporter m reports row m and column m
  {"kind": "all"}                                   every reporter reports every tie (dense R == 1)
  {"kind": "coo", "subs": (4,nnz), "vals": (nnz,), "dense_input": bool}
"""
import numpy as np
import scipy.special as sp


def _gamma_elbo_term(pa, pb, qa, qb):
    # model.py:1300-1303
    return sp.gammaln(qa) - pa * np.log(qb) + (pa - qa) * sp.psi(qa) + qa * (1 - pb / qb)


class OracleCAVI:
    def __init__(self, L, N, M, K, X_subs, X_vals, R_spec, mutuality=True, alpha_theta=0.1, beta_theta=0.1,
                 alpha_lambda=10.0, beta_lambda=10.0, alpha_eta=0.5, beta_eta=1.0, EPS=1e-12):
        self.L, self.N, self.M, self.K = int(L), int(N), int(M), int(K)
        self.mutuality = bool(mutuality)
        self.EPS = float(EPS)
        X_subs = np.asarray(X_subs)
        self.xl, self.xi, self.xj, self.xm = (X_subs[d].astype(np.int64) for d in range(4))
        self.xv = np.asarray(X_vals).astype(np.float64)
        self.sumX = self.xv.sum()
        self.alpha_theta = np.broadcast_to(np.asarray(alpha_theta, dtype=float), (L, M)).copy()
        self.beta_theta = np.broadcast_to(np.asarray(beta_theta, dtype=float), (L, M)).copy()
        self.alpha_lambda = np.broadcast_to(np.asarray(alpha_lambda, dtype=float), (L, K)).copy()
        self.beta_lambda = np.broadcast_to(np.asarray(beta_lambda, dtype=float), (L, K)).copy()
        self.alpha_eta, self.beta_eta = float(alpha_eta), float(beta_eta)
        self.R = dict(R_spec)
        self._prepare_pairs()

    # ---------------------------------------------------------------- setup
    def _key(self, l, i, j, m):
        return ((l * self.N + i) * self.N + j) * self.M + m

    def _in_R(self, l, i, j, m):
        """Multiplicity of (l,i,j,m) in R (0/1 for a proper mask)."""
        kind = self.R["kind"]
        if kind == "all":
            return np.ones(len(l))
        if kind == "ego":
            rep = np.asarray(self.R["rep"]).astype(bool)
            ok = ((i == m) | (j == m)) & rep[l, m]
            if not self.R["diag"]:
                ok &= ~((i == m) & (j == m))
            return ok.astype(float)
        subs = np.asarray(self.R["subs"]).astype(np.int64)
        rk = np.sort(self._key(subs[0], subs[1], subs[2], subs[3]))
        k = self._key(l, i, j, m)
        return (np.searchsorted(rk, k, side="right") - np.searchsorted(rk, k, side="left")).astype(float)

    def _prepare_pairs(self):
        """Reciprocal value X[l,j,i,m] for every X entry (`data_T_vals`, model.py:152-161) and the
        R-membership flags the ELBO needs (model.py:977-995, 1269-1286)."""
        k = self._key(self.xl, self.xi, self.xj, self.xm)
        order = np.argsort(k, kind="stable")
        ks = k[order]
        kt = self._key(self.xl, self.xj, self.xi, self.xm)
        pos = np.searchsorted(ks, kt)
        pos_c = np.minimum(pos, max(len(ks) - 1, 0))
        found = (ks[pos_c] == kt) if len(ks) else np.zeros(0, dtype=bool)
        self.xT = np.where(found, self.xv[order][pos_c], 0.0) if len(ks) else np.zeros(0)
        self.x_inR = self._in_R(self.xl, self.xi, self.xj, self.xm)  # X entry itself present in R
        self.xT_inR = self._in_R(self.xl, self.xj, self.xi, self.xm)  # its transposed position present in R

    def default_pr_rho(self, pr_ties=None, pr_vals=None):
        """One-hot prior everywhere (model.py:536-556) except on the given ties."""
        pr = np.zeros((self.L, self.N, self.N, self.K))
        pr[..., 0] = 1.0
        if pr_ties is not None and len(pr_ties):
            pr_ties = np.asarray(pr_ties)
            pr[pr_ties[:, 0], pr_ties[:, 1], pr_ties[:, 2], :] = pr_vals
        return pr

    def set_state(self, gamma_shp, gamma_rte, phi_shp, phi_rte, nu_shp, pr_rho, nu_rte=None):
        """Inject the initial state (model.py:561-605)."""
        self.gamma_shp = np.array(gamma_shp, dtype=float)
        self.gamma_rte = np.array(gamma_rte, dtype=float)
        self.phi_shp = np.array(phi_shp, dtype=float)
        self.phi_rte = np.array(phi_rte, dtype=float)
        if self.mutuality:
            self.nu_shp = float(nu_shp)
            self.nu_rte = float(self.beta_eta + self.sumX) if nu_rte is None else float(nu_rte)
            self.G_exp_nu = np.exp(sp.psi(self.nu_shp) - np.log(self.nu_rte))
        else:  # model.py:597-600
            self.nu_shp, self.nu_rte, self.G_exp_nu = 0.000001, 1.0, 0.0
        self.pr_rho = np.array(pr_rho, dtype=float)
        self.logpr_rho = np.log(self.pr_rho + self.EPS)
        self.rho = self.pr_rho.copy()
        self._cache()

    # ---------------------------------------------------------------- pieces
    def _cache(self):
        # model.py:676-696
        self.G_exp_theta = np.exp(sp.psi(self.gamma_shp) - np.log(self.gamma_rte))
        self.G_exp_lambda = np.exp(sp.psi(self.phi_shp) - np.log(self.phi_rte))
        if not self.mutuality:
            self.dz1 = np.repeat(self.xv[:, None], 1, axis=1)  # (I,1), broadcast over k
            self.dz2 = None
            return
        self.G_exp_nu = np.exp(sp.psi(self.nu_shp) - np.log(self.nu_rte))
        z1 = self.G_exp_theta[self.xl, self.xm][:, None] * self.G_exp_lambda[self.xl, :]
        z2 = self.G_exp_nu * self.xT
        den = z1 + z2[:, None]
        den[den == 0] = 1
        self.dz1 = self.xv[:, None] * z1 / den
        self.dz2 = self.xv[:, None] * z2[:, None] / den

    def _reporter_sums(self, W, use_vals=False):
        """sum over R entries J of W[l,i,j] scattered to (l,m)   [W is (L,N,N)] -> (L,M)."""
        L, N, M = self.L, self.N, self.M
        kind = self.R["kind"]
        if kind == "all":
            return np.repeat(W.sum(axis=(1, 2))[:, None], M, axis=1)
        if kind == "ego":
            rep = np.asarray(self.R["rep"]).astype(float)
            out = np.zeros((L, M))
            rows = W.sum(axis=2)[:, :M]
            cols = W.sum(axis=1)[:, :M]
            dg = np.einsum("lii->li", W)[:, :M]
            out = rows + cols - (dg if self.R["diag"] else 2 * dg)
            return out * rep
        subs = np.asarray(self.R["subs"]).astype(np.int64)
        w = W[subs[0], subs[1], subs[2]]
        if use_vals:
            w = w * np.asarray(self.R["vals"])
        out = np.zeros((L, M))
        np.add.at(out, (subs[0], subs[3]), w)
        return out

    def _tie_sums(self, E):
        """S[l,i,j] = sum over R entries of E[l,m] * R.vals   (model.py:766-792) -> (L,N,N)."""
        L, N, M = self.L, self.N, self.M
        kind = self.R["kind"]
        if kind == "all":
            return np.broadcast_to(E.sum(axis=1)[:, None, None], (L, N, N)).copy()
        if kind == "ego":
            rep = np.asarray(self.R["rep"]).astype(float)
            Er = np.zeros((L, N))
            Er[:, :M] = E * rep
            S = Er[:, :, None] + Er[:, None, :]
            idx = np.arange(N)
            S[:, idx, idx] = Er if self.R["diag"] else 0.0
            return S
        subs = np.asarray(self.R["subs"]).astype(np.int64)
        S = np.zeros((L, N, N))
        np.add.at(S, (subs[0], subs[1], subs[2]), E[subs[0], subs[3]] * np.asarray(self.R["vals"]))
        return S

    # ---------------------------------------------------------------- updates
    def update_gamma(self):
        # shape: model.py:700, 851-859
        rho_I = self.rho[self.xl, self.xi, self.xj, :]
        tmp = (rho_I * self.dz1).sum(axis=1)
        acc = np.zeros((self.L, self.M))
        np.add.at(acc, (self.xl, self.xm), tmp)
        self.gamma_shp = self.alpha_theta + acc
        # rate: model.py:704-718 (sparse R ignores R.vals; dense R multiplies by R -- Q4)
        E_phi_rho = np.einsum("lijk,lk->lij", self.rho, self.phi_shp / self.phi_rte)
        use_vals = bool(self.R.get("dense_input", False))
        self.gamma_rte = self.beta_theta + self._reporter_sums(E_phi_rho, use_vals=use_vals)

    def update_phi(self):
        # shape: model.py:731, 880-887
        rho_I = self.rho[self.xl, self.xi, self.xj, :]
        tmp = rho_I * self.dz1
        acc = np.zeros((self.L, self.K))
        np.add.at(acc, self.xl, tmp)
        self.phi_shp = self.alpha_lambda + acc
        # rate: model.py:742-749  (R.vals never used)
        Eg = self.gamma_shp / self.gamma_rte
        out = np.zeros((self.L, self.K))
        kind = self.R["kind"]
        if kind == "coo":
            subs = np.asarray(self.R["subs"]).astype(np.int64)
            t = self.rho[subs[0], subs[1], subs[2], :] * Eg[subs[0], subs[3]][:, None]
            np.add.at(out, subs[0], t)
        else:
            for k in range(self.K):
                A_k = self._reporter_sums(self.rho[..., k])  # (L,M): sum of rho_k over ties reported by m
                out[:, k] = (A_k * Eg).sum(axis=1)
        self.phi_rte = self.beta_lambda + out

    def update_rho(self):
        # model.py:763-818, 908-923 (no max-subtraction; normalise only where the sum is > 0 -- Q3)
        S = self._tie_sums(self.gamma_shp / self.gamma_rte)
        Exp_theta_lambda = np.einsum("lij,lk->lijk", S, self.phi_shp / self.phi_rte)
        ElogT = sp.psi(self.gamma_shp) - np.log(self.gamma_rte)
        ElogL = sp.psi(self.phi_shp) - np.log(self.phi_rte)
        tmp = (ElogT[self.xl, self.xm][:, None] + ElogL[self.xl, :]) * self.dz1
        add = np.zeros_like(self.rho)
        np.add.at(add, (self.xl, self.xi, self.xj), tmp)
        log_rho = self.logpr_rho + add - Exp_theta_lambda
        with np.errstate(under="ignore"):
            rho = np.exp(log_rho)
        s = rho.sum(axis=3)
        ok = s > 0
        rho[ok] /= s[ok, None]
        self.rho = rho

    def update_nu(self):
        # model.py:820-830
        self.nu_shp = self.alpha_eta + (self.dz2 * self.rho[self.xl, self.xi, self.xj, :]).sum()

    def iterate(self):
        """One `_update_CAVI` (model.py:623-660)."""
        self._cache()
        self.update_gamma()
        self._cache()
        self.update_phi()
        self._cache()
        self.update_rho()
        if self.mutuality:
            self._cache()  # G_exp_nu now reflects the OLD nu_shp; it stays stale for the ELBO (Q2)
            self.update_nu()

    # ---------------------------------------------------------------- ELBO
    def elbo(self):
        """`__ELBO` (model.py:948-1019) incl. Q1 (exp(rho)) and Q2 (stale G_exp_nu)."""
        Et = self.gamma_shp / self.gamma_rte
        El = self.phi_shp / self.phi_rte
        Eeta = self.nu_shp / self.nu_rte
        # term 1: - sum_J sum_k rho_k (Et El_k + Eeta * X[l,j,i,m])      model.py:957-965, 1257-1291
        t1 = 0.0
        for k in range(self.K):
            t1 += (self._reporter_sums(self.rho[..., k]) * Et).sum(axis=1) @ El[:, k]
        if self.mutuality:
            rs = self.rho.sum(axis=3)
            # X entry I=(l,i,j,m) is the "X_T" value of the R entry (l,j,i,m)
            t1 += Eeta * (self.xv * self.xT_inR * rs[self.xl, self.xj, self.xi]).sum()
        elbo = -t1
        # term 2: sum_I x log(EPS + sum_k exp(rho_k)(Gt Gl_k + Gnu xT)) over X entries present in R
        erho = np.exp(self.rho[self.xl, self.xi, self.xj, :])
        Gnu = self.G_exp_nu if self.mutuality else 0.0
        mean = self.G_exp_theta[self.xl, self.xm][:, None] * self.G_exp_lambda[self.xl, :] + (Gnu * self.xT)[:, None]
        val = (erho * mean).sum(axis=1) * self.x_inR
        elbo += (self.xv * np.log(val + self.EPS)).sum()
        # gamma terms  model.py:997-1011
        elbo += _gamma_elbo_term(self.alpha_theta, self.beta_theta, self.gamma_shp, self.gamma_rte).sum()
        elbo += _gamma_elbo_term(self.alpha_lambda, self.beta_lambda, self.phi_shp, self.phi_rte).sum()
        elbo += _gamma_elbo_term(self.alpha_eta, self.beta_eta, self.nu_shp, self.nu_rte)
        # categorical term  model.py:1306-1313
        elbo += (self.rho * (np.log(self.pr_rho + self.EPS) - np.log(self.rho + self.EPS))).sum()
        return float(elbo)

    # ---------------------------------------------------------------- driver
    def run(self, max_iter, convergence_tol=0.1, decision=1, record=None):
        """The `while` loop of `fit` (model.py:399-426) with `_check_for_convergence` (model.py:1021-1056)."""
        INF = 1e10
        coincide, it, conv, elbo = 0, 1, False, -INF
        trace = []
        while not conv and it <= max_iter:
            self.iterate()
            if it == 1 or it % 10 == 0 or it == max_iter:
                old = elbo
                elbo = self.elbo()
                coincide = coincide + 1 if abs(elbo - old) < convergence_tol else 0
            if coincide > decision:
                conv = True
            if record is not None:
                record(self, it)
            it += 1
            if (it - 1) % 10 == 0:
                trace.append((it - 1, elbo, conv))
        return elbo, trace
