"""Small helpers of the public surface (reference `utils.py`)."""
import numpy as np

from .sptensor import dtensor, is_sparse_like, sptensor


def match_arg(x, lst):
    return [el for el in lst if x == el]


def is_sparse(X):
    """Heuristic of the reference (utils.py:87-112): sparse iff size > (nnz + 1) * ndim."""
    return X.size > (X.nonzero()[0].size + 1) * X.ndim


def preprocess(X):
    """ndarray -> int sptensor (if sparse enough) or dtensor; tensors pass through (reference utils.py:220-248)."""
    if is_sparse_like(X) or isinstance(X, dtensor):
        return X
    X = np.asarray(X)
    if not X.dtype == np.dtype(int).type:
        X = X.astype(int)
    return sptensor.fromarray(X) if is_sparse(X) else dtensor(X)


def get_optimal_threshold(model):
    """https://arxiv.org/pdf/2112.11396.pdf pg 8 (reference utils.py:200-204; reads G_exp_nu, not G_exp_nu_f)."""
    return 0.54 * model.G_exp_nu - 0.01


def apply_rho_threshold(model, threshold=None):
    """Binarise rho_f[..., 1]: 1.0 where it reaches the threshold, else 0.0 (restated from reference utils.py:207-217;
    host-side helper -- `VimureModel.get_inferred_model` does this on the device, `vm_infer` mode 1)."""
    thr = get_optimal_threshold(model) if threshold is None else threshold
    r1 = np.asarray(model.rho_f)[..., 1]
    return (r1 >= thr).astype(r1.dtype)


def calculate_overall_reciprocity(Y):
    return np.logical_and(Y > 0, np.transpose(Y, axes=(1, 0)) > 0).sum() / Y.sum()
