"""Reporter-mask (R) representations.

The reference stores the mask R (L x N x N x M) either as a dense all-ones array (default when R is not
given, `model.py:207-211`) or as an explicit COO tensor (`_io.py:229-242`, `synthetic.py:1184-1204`), whose
size is nnz(R) ~ 2*L*M*N for the usual "a reporter reports the ties she is part of" mask -- 8e8 entries at
N = 20k.  vimure_b200 recognises the two structured cases so that they never have to be materialised:

  EgoMask : reporter m is node m and reports row m and column m of every layer she is active in, with or
            without the (m,m) tie (`build_self_reporter_mask` includes it, `read_from_edgelist` does not);
  AllMask : every reporter reports every tie (R == 1);
  CooMask : anything else, explicit.

`ReporterMask.from_input` accepts what the reference accepts for `R` (dense ndarray / dtensor / sptensor-like)
plus these classes themselves.
"""
import numpy as np
import torch

from .sptensor import is_sparse_like


class ReporterMask:
    kind = None
    dense_input = False  # True when the user passed a dense array (the reference's dtensor branches)

    def to_sptensor(self):
        raise NotImplementedError


class AllMask(ReporterMask):
    kind = "all"

    def __init__(self, L, N, M):
        self.L, self.N, self.M = int(L), int(N), int(M)
        self.dense_input = True

    @property
    def shape(self):
        return (self.L, self.N, self.N, self.M)

    def entry_multiplicity(self, l, i, j, m):
        return torch.ones(l.shape, dtype=torch.float64, device=l.device)

    def tie_reported(self, l, i, j):
        return torch.ones(l.shape, dtype=torch.bool, device=l.device)

    def toarray(self):
        return np.ones(self.shape)


class EgoMask(ReporterMask):
    kind = "ego"

    def __init__(self, L, N, M, rep=None, diag=True):
        self.L, self.N, self.M = int(L), int(N), int(M)
        if self.M > self.N:
            raise ValueError("EgoMask needs M <= N (reporter m is node m)")
        if rep is None:
            rep = np.ones((self.L, self.M), dtype=np.uint8)
        rep = np.asarray(rep)
        if rep.ndim == 1:  # list of reporter node ids, active in every layer
            r = np.zeros((self.L, self.M), dtype=np.uint8)
            r[:, rep.astype(np.int64)] = 1
            rep = r
        self.rep = np.ascontiguousarray(rep.astype(np.uint8)).reshape(self.L, self.M)
        self.diag = bool(diag)

    @property
    def shape(self):
        return (self.L, self.N, self.N, self.M)

    def _rep_t(self, device):
        return torch.as_tensor(self.rep.astype(bool), device=device)

    def entry_multiplicity(self, l, i, j, m):
        rep = self._rep_t(l.device)
        ok = ((i == m) | (j == m)) & rep[l, m]
        if not self.diag:
            ok &= ~((i == m) & (j == m))
        return ok.to(torch.float64)

    def tie_reported(self, l, i, j):
        rep = self._rep_t(l.device)
        M = self.M
        ri = rep[l, i.clamp(max=M - 1)] & (i < M)
        rj = rep[l, j.clamp(max=M - 1)] & (j < M)
        off = (ri | rj) & (i != j)
        dg = (i == j) & ri & self.diag
        return off | dg

    def to_sptensor(self):
        from .sptensor import sptensor

        subs = []
        N = self.N
        for l in range(self.L):
            for m in np.nonzero(self.rep[l])[0]:
                o = np.delete(np.arange(N), m)
                subs.append(np.stack([np.full(N - 1, l), np.full(N - 1, m), o, np.full(N - 1, m)]))
                subs.append(np.stack([np.full(N - 1, l), o, np.full(N - 1, m), np.full(N - 1, m)]))
                if self.diag:
                    subs.append(np.array([[l], [m], [m], [m]]))
        s = np.concatenate(subs, axis=1) if subs else np.zeros((4, 0), dtype=np.int64)
        return sptensor(tuple(s), np.ones(s.shape[1], dtype=np.int64), shape=self.shape)


class CooMask(ReporterMask):
    kind = "coo"

    def __init__(self, subs, vals, shape, dense_input=False):
        self.subs = np.stack([np.asarray(s).astype(np.int64) for s in subs])
        self.vals = np.asarray(vals).astype(np.float64)
        self.shape = tuple(int(d) for d in shape)
        self.L, self.N, _, self.M = self.shape
        self.dense_input = bool(dense_input)
        self._keys = None

    def _key(self, l, i, j, m):
        return ((l * self.N + i) * self.N + j) * self.M + m

    def _sorted_keys(self, device):
        if self._keys is None or self._keys.device != device:
            s = torch.as_tensor(self.subs, device=device)
            self._keys = torch.sort(self._key(s[0], s[1], s[2], s[3]))[0]
            self._tkeys = torch.unique(((s[0] * self.N + s[1]) * self.N + s[2]))
        return self._keys

    def entry_multiplicity(self, l, i, j, m):
        keys = self._sorted_keys(l.device)
        k = self._key(l, i, j, m)
        hi = torch.searchsorted(keys, k, right=True)
        lo = torch.searchsorted(keys, k, right=False)
        return (hi - lo).to(torch.float64)

    def tie_reported(self, l, i, j):
        self._sorted_keys(l.device)
        tk = self._tkeys
        k = (l * self.N + i) * self.N + j
        if tk.numel() == 0:
            return torch.zeros(l.shape, dtype=torch.bool, device=l.device)
        pos = torch.searchsorted(tk, k).clamp(max=tk.numel() - 1)
        return tk[pos] == k


def _ego_entries_distinct(l, i, j, m, L, N, M):
    """No (l, i, j, m) occurs twice, given that every entry has i == m or j == m: an entry is then identified by its
    reporter, the side the reporter is on and the other node -- one flag per such slot (O(n)) instead of a sort / hash
    of the 4-subscript keys (np.unique took 0.34 s of a 0.4 s fit set-up on the 6.8e5-entry Karnataka mask)."""
    n = l.size
    slots = 2 * L * M * N
    if slots > 8 * n + (1 << 20):  # far sparser than an ego mask can be: the sort is cheaper than the flag array
        key = np.sort(((l * N + i) * N + j) * M + m)
        return not bool(np.any(key[1:] == key[:-1]))
    row_side = i == m  # reporter is the source (this includes the diagonal entry)
    slot = ((l * M + m) * 2 + np.where(row_side, 0, 1)) * N + np.where(row_side, j, i)
    seen = np.zeros(slots, dtype=bool)
    seen[slot] = True
    return int(np.count_nonzero(seen)) == n


def _detect_ego(subs, vals, L, N, M):
    """Return an EgoMask if the COO mask is exactly an ego mask, else None."""
    if M > N or len(vals) == 0:
        return None
    l, i, j, m = (np.asarray(s).astype(np.int64) for s in subs)
    if not np.all(vals == 1):
        return None
    if not np.all((i == m) | (j == m)):
        return None
    cnt = np.bincount(l * M + m, minlength=L * M).reshape(L, M)
    rep = cnt > 0
    ndiag = int(np.count_nonzero((i == m) & (j == m)))
    for diag in (True, False):
        full = 2 * N - 1 if diag else 2 * N - 2
        if np.all(cnt[rep] == full) and ndiag == (int(rep.sum()) if diag else 0):
            if _ego_entries_distinct(l, i, j, m, L, N, M):
                return EgoMask(L, N, M, rep=rep.astype(np.uint8), diag=diag)
    return None


def from_input(R, L, N, M):
    """Coerce what the user passed as R (reference `model.py:199-213`: dense array, dtensor or sptensor)."""
    if isinstance(R, ReporterMask):
        if tuple(R.shape) != (L, N, N, M):
            raise ValueError("Dimensions of reporter mask (R) do not match L x N x N x M")
        return R
    if is_sparse_like(R):
        shape = tuple(int(d) for d in R.shape)
        if shape != (L, N, N, M):
            raise ValueError("Dimensions of reporter mask (R) do not match L x N x N x M")
        vals = np.asarray(R.vals)
        ego = _detect_ego(R.subs, vals, L, N, M)
        if ego is not None:
            return ego
        return CooMask(R.subs, vals, shape, dense_input=False)
    Rd = np.asarray(R)
    if Rd.shape != (L, N, N, M):
        raise ValueError("Dimensions of reporter mask (R) do not match L x N x N x M")
    Rd = Rd.astype(int)  # preprocess(): utils.py:241-242
    if np.all(Rd == 1):
        return AllMask(L, N, M)
    subs = np.nonzero(Rd)
    vals = Rd[subs]
    # the reference keeps a sufficiently sparse array as sptensor (utils.py:243-246, is_sparse utils.py:87-112)
    sparse = Rd.size > (len(vals) + 1) * Rd.ndim
    if sparse:
        ego = _detect_ego(subs, vals, L, N, M)
        if ego is not None:
            return ego
    elif not np.all(vals == 1):
        raise ValueError("A dense reporter mask must be binary (0/1).")
    return CooMask(subs, vals, Rd.shape, dense_input=not sparse)
