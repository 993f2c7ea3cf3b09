"""vimure_b200: B200-native CAVI for VIMuRe (drop-in for `vimure.model.VimureModel`).

Like the reference package (`vimure/__init__.py:1-2`) it exposes `model` and `synthetic`; `io`, `utils`, `masks` and
`sptensor` are reachable as sub-modules."""
from . import io, masks, model, sptensor, synthetic, utils  # noqa: F401
from .model import VimureModel  # noqa: F401

__all__ = ["model", "synthetic", "io", "masks", "sptensor", "utils", "VimureModel"]
