"""Minimal COO / dense tensor containers accepted (and produced) by vimure_b200.

The reference passes data around as `sktensor.sptensor` / `sktensor.dtensor` objects (third-party
scikit-tensor; reference `utils.py:115-181, 220-248`, `model.py:147-170`).  vimure_b200 does not depend on
that package: any object with `.subs` (tuple of 4 index arrays), `.vals` and `.shape` is accepted as a
sparse tensor, and these two classes are what the package itself returns.
"""
import numpy as np


class sptensor(object):
    """Sparse COO tensor; keeps `subs` / `vals` in construction order."""

    def __init__(self, subs, vals, shape=None, dtype=None):
        if not isinstance(subs, tuple):
            raise ValueError("Subscripts must be a tuple of array-likes")
        if len(subs[0]) != len(vals):
            raise ValueError("Subscripts and values must be of equal length")
        self.subs = tuple(np.asarray(s) for s in subs)
        self.vals = np.asarray(vals) if dtype is None else np.asarray(vals, dtype=dtype)
        self.dtype = self.vals.dtype
        if shape is None:
            shape = tuple(int(np.max(s)) + 1 for s in self.subs)
        self.shape = tuple(int(d) for d in shape)
        self.ndim = len(self.subs)

    def __len__(self):
        return len(self.vals)

    def toarray(self):
        A = np.zeros(self.shape, dtype=self.vals.dtype)
        if len(self.vals):
            A[tuple(self.subs)] = self.vals
        return A

    @staticmethod
    def fromarray(A):
        A = np.asarray(A)
        subs = np.nonzero(A)
        return sptensor(subs, A[subs], shape=A.shape, dtype=A.dtype)


class dtensor(np.ndarray):
    """Dense tensor (an ndarray subclass), the counterpart of `sktensor.dtensor`."""

    def __new__(cls, input_array):
        return np.asarray(input_array).view(cls)

    def toarray(self):
        return np.asarray(self)


def is_sparse_like(X):
    return hasattr(X, "subs") and hasattr(X, "vals") and hasattr(X, "shape")
