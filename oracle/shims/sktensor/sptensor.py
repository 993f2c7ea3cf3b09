"""COO container part of the ``sktensor`` stand-in (see package docstring). Test infrastructure only."""
import numpy as np


class sptensor(object):
    """Sparse COO tensor: keeps ``subs`` (tuple of index arrays) and ``vals`` in construction order."""

    def __init__(self, subs, vals, shape=None, dtype=None, accumfun=None, issorted=False):
        if not isinstance(subs, tuple):
            raise ValueError("Subscripts must be a tuple of array-likes")
        if len(subs[0]) != len(vals):
            raise ValueError("Subscripts and values must be of equal length")
        if dtype is None:
            dtype = np.array(vals).dtype
        for s in subs:
            if len(s) and np.array(s).dtype.kind not in "iu":
                raise ValueError("Subscripts must be integers")
        self.subs = subs
        self.vals = np.array(vals, dtype=dtype)
        self.dtype = dtype
        self.issorted = issorted
        self.accumfun = accumfun
        if shape is None:
            self.shape = tuple(int(np.max(s)) + 1 for s in subs)
        else:
            self.shape = tuple(int(d) for d in shape)
        self.ndim = len(subs)

    def __getitem__(self, idx):
        if len(idx) != self.ndim:
            raise ValueError("subscripts must be complete")
        sel = np.ones(len(self.vals), dtype=bool)
        for d in range(self.ndim):
            sel = np.logical_and(np.asarray(self.subs[d]) == idx[d], sel)
        vals = self.vals[sel]
        if len(vals) == 0:
            vals = 0
        elif len(vals) > 1:
            if self.accumfun is None:
                raise ValueError("Duplicate entries without specified accumulation function")
            vals = self.accumfun(vals)
        return vals

    def __len__(self):
        return len(self.vals)

    def toarray(self):
        A = np.zeros(self.shape, dtype=self.vals.dtype)
        if len(self.vals):
            A.put(np.ravel_multi_index(tuple(np.asarray(s) for s in self.subs), self.shape), self.vals)
        return A


def fromarray(A):
    """Create a sptensor from a dense numpy array."""
    A = np.asarray(A)
    subs = np.nonzero(A)
    vals = A[subs]
    return sptensor(subs, vals, shape=A.shape, dtype=A.dtype)
