"""Stand-in for the un-vendored third-party ``sktensor`` (scikit-tensor-py3 0.4.2) module.

TEST INFRASTRUCTURE ONLY.  The reference (`/root/reference/src/python/vimure`) imports
``sktensor`` (pinned as an un-vendored zip in ``src/python/poetry.lock:255-271``) but only uses it
as a COO container: ``sptensor(subs, vals, shape=, dtype=)`` with ``.subs .vals .shape .ndim
.toarray() [tuple-index]``, ``dtensor(ndarray)`` and ``sktensor.sptensor.fromarray``
(SURVEY.md section 8c lists every call site).  None of the CAVI arithmetic lives in sktensor.
This module restates that published container behaviour so that the UNMODIFIED reference can be
imported in the build container to pin the oracle and generate golden vectors.  It is never
imported by the product package.
"""
import numpy as np

from . import sptensor as _sptensor_mod
from .sptensor import sptensor, fromarray  # noqa: F401


class dtensor(np.ndarray):
    """Dense tensor: an ndarray subclass (only isinstance checks and ndarray methods are used)."""

    def __new__(cls, input_array):
        return np.asarray(input_array).view(cls)

    def toarray(self):
        return np.asarray(self)


__all__ = ["sptensor", "dtensor", "fromarray"]
