"""Randomised parity sweep: the CUDA path (through the C ABI) against the CPU oracle on seeded random problems that vary
everything the kernels branch on -- layers, nodes (multiples of 4 or not: fast / general dense kernel, full / partial
column tiles), reporters (M <= N), categories (K = 2..5), the mask structure (ego, ego with a subset of reporters and no
diagonal, all-reporter, general COO with X entries outside R and repeated reporters per tie), mutuality, the row-tile
height, ties with several X entries and self-loops."""
import numpy as np
import pytest

from tests.test_gpu_edges import PRI, _run, _state

pytestmark = pytest.mark.gpu


def _random_problem(seed):
    import vimure_b200 as vm
    import vimure_b200.synthetic as syn

    rng = np.random.RandomState(1000 + seed)
    L = int(rng.randint(1, 4))
    K = int(rng.choice([2, 2, 3, 4, 5]))
    N = int(rng.choice([36, 61, 128, 250, 516, 700]))
    kind = ["ego", "ego_subset", "all", "coo"][seed % 4]
    if kind in ("all", "coo"):
        N = min(N, 128)  # the oracle materialises the explicit mask
    mutuality = bool(rng.rand() < 0.75)
    tile_h = int(rng.choice([8, 32, 128]))
    M = N if kind == "ego" else int(rng.randint(5, min(N, 40) + 1))
    eta = float(rng.choice([0.0, 0.3, 0.6]))
    law = syn.StandardSBM if rng.rand() < 0.5 else syn.Multitensor
    kw = dict(N=N, M=M, L=L, K=K, C=2, avg_degree=float(rng.choice([3, 8])), seed=seed)
    if law is syn.Multitensor:
        kw["eta"] = max(eta, 0.1)
    y = law(**kw)
    if kind == "all":
        X, _ = syn.dense_reporting_X(y, M=M, mutuality=eta, seed=seed + 1)
        subs, vals = np.stack(X.subs).astype(np.int64), np.asarray(X.vals)
        mask, spec = vm.masks.AllMask(L, N, M), {"kind": "all", "dense_input": True}
    else:
        net = y.build_X(mutuality=eta, seed=seed + 1)
        subs, vals = np.stack(net.X.subs).astype(np.int64), np.asarray(net.X.vals)
        if kind == "ego":
            mask = net.R
            spec = {"kind": "ego", "rep": np.ones((L, M), dtype=np.uint8), "diag": True}
        else:
            rep = (rng.rand(L, M) < 0.7).astype(np.uint8)
            rep[:, 0] = 1
            keep = (subs[3] < M)
            keep &= rep[subs[0], np.minimum(subs[3], M - 1)] == 1
            subs, vals = subs[:, keep], vals[keep]
            ego = vm.masks.EgoMask(L, N, M, rep=rep, diag=False)
            if kind == "ego_subset":
                mask, spec = ego, {"kind": "ego", "rep": rep, "diag": False}
            else:
                # a general mask: most of the ego entries (some dropped, so that X has entries outside R), a few weights
                # of 2, and extra reporters on random ties
                R = ego.to_sptensor()
                rs, rv = np.stack(R.subs).astype(np.int64), np.asarray(R.vals).astype(np.float64)
                k2 = rng.rand(rs.shape[1]) < 0.9
                rs, rv = rs[:, k2], rv[k2]
                rv[rng.rand(rv.size) < 0.05] = 2.0
                n_extra = 3 * N
                ex = np.stack([rng.randint(0, L, n_extra), rng.randint(0, N, n_extra), rng.randint(0, N, n_extra),
                               rng.randint(0, M, n_extra)])
                allk = np.concatenate([rs, ex], axis=1)
                _, first = np.unique(np.ravel_multi_index(tuple(allk), (L, N, N, M)), return_index=True)
                first.sort()
                rs = allk[:, first]
                rv = np.concatenate([rv, np.ones(n_extra)])[first]
                mask = vm.masks.CooMask(tuple(rs), rv, (L, N, N, M), dense_input=False)
                spec = {"kind": "coo", "subs": rs, "vals": rv, "dense_input": False}
        # a few self-loops and second reports on existing ties (several X entries per tie)
        if subs.shape[1] > 10 and kind != "all":
            take = rng.choice(subs.shape[1], size=min(20, subs.shape[1]), replace=False)
            extra = subs[:, take].copy()
            extra[3] = (extra[3] + 1 + rng.randint(0, max(M - 1, 1), extra.shape[1])) % M  # another reporter, same tie
            loops = subs[:, take[:5]].copy()
            loops[2] = loops[1]
            allk = np.concatenate([subs, extra, loops], axis=1)
            allv = np.concatenate([vals, rng.randint(1, 4, extra.shape[1]), rng.randint(1, 3, loops.shape[1])])
            _, first = np.unique(np.ravel_multi_index(tuple(allk), (L, N, N, M)), return_index=True)
            first.sort()
            subs, vals = allk[:, first], allv[first]
    if subs.shape[1] == 0:
        pytest.skip("empty random problem")
    return dict(L=L, N=N, M=M, K=K, subs=subs, vals=vals, mask=mask, spec=spec, mutuality=mutuality, tile_h=tile_h,
                kind=kind)


@pytest.mark.parametrize("seed", list(range(12)))
def test_random_problem_matches_oracle(seed):
    p = _random_problem(seed)
    st = _state(p["L"], p["M"], p["K"], seed + 50)
    _run(p["L"], p["N"], p["M"], p["K"], p["subs"], p["vals"], p["mask"], p["spec"], st, PRI, mutuality=p["mutuality"],
         iters=3, tile_h=p["tile_h"])
