"""Moments of the reference's report generator (`vimure.synthetic._build_X` under the self-reporter mask,
`/root/reference/src/python/vimure/synthetic.py:63-232`) on a fixed ground truth Y and fixed reliabilities theta, over
several seeds.  TEST INFRASTRUCTURE ONLY -- the fixture `tests/golden/synth_moments.npz` pins the LAW of this repo's
sparse generators (host: vimure_b200.synthetic.build_X, device: vm_synth_ego), whose RNG streams are not numpy's.

    python oracle/gen_synth_moments.py        # needs /root/reference (build container only)

Statistics, per seed and per class c = (Y_ij > 0, Y_ji > 0) of the ordered pair as seen by a reporter m in {i, j}, i != j:
    n[c]  number of (tie, reporter) slots;   s1[c] = sum X;   p[c] = #(X > 0);   s2[c] = sum X^2;   sx[c] = sum X * X^T
plus the same for the self ties (class 4).  X = X[l,i,j,m], X^T = X[l,j,i,m].
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.ref_runner import import_reference  # noqa: E402


def class_stats(Xd, Yd):
    """Xd: dense (L,N,N,M) counts, Yd: dense (L,N,N) ground truth.  Returns (5, 5) array [class, (n, s1, p, s2, sx)]."""
    L, N, _, M = Xd.shape
    out = np.zeros((5, 5))
    for l in range(L):
        for m in range(M):
            for (rows, cols) in ((np.full(N, m), np.arange(N)), (np.arange(N), np.full(N, m))):
                i, j = rows, cols
                off = i != j
                x = Xd[l, i, j, m].astype(float)
                xt = Xd[l, j, i, m].astype(float)
                cls = 2 * (Yd[l, i, j] > 0).astype(int) + (Yd[l, j, i] > 0).astype(int)
                for c in range(4):
                    s = off & (cls == c)
                    out[c] += [s.sum(), x[s].sum(), (x[s] > 0).sum(), (x[s] ** 2).sum(), (x[s] * xt[s]).sum()]
            x = float(Xd[l, m, m, m])
            out[4] += [1, x, x > 0, x * x, x * x]
    return out


def main():
    vm = import_reference()
    N, L, K, eta = 120, 1, 2, 0.5
    gt = vm.synthetic.StandardSBM(N=N, M=N, L=L, K=K, C=2, avg_degree=8, sparsify=True, seed=3)
    Yd = gt.Y.toarray()
    theta = np.random.RandomState(5).gamma(shape=2.0, scale=0.5, size=(L, N))
    seeds = list(range(100, 116))
    stats = []
    for sd in seeds:
        gt._build_X(mutuality=eta, theta=theta, flag_self_reporter=True, cutoff_X=False, seed=sd)
        stats.append(class_stats(gt.X.toarray(), Yd))
        print("seed", sd, stats[-1][:, 1])
    ys = np.stack([np.asarray(s) for s in gt.Y.subs]).astype(np.int32)
    path = os.path.join(ROOT, "tests", "golden", "synth_moments.npz")
    np.savez_compressed(path, dims=np.array([L, N, N, K]), eta=np.array(eta), theta=theta, Y_subs=ys,
                        Y_vals=np.asarray(gt.Y.vals).astype(np.int32), seeds=np.array(seeds), stats=np.stack(stats))
    print("wrote", path)


if __name__ == "__main__":
    main()
