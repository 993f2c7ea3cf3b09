"""vimure_b200: B200-native CAVI for VIMuRe (drop-in for `vimure.model.VimureModel`)."""
from . import masks, model, sptensor, utils  # noqa: F401
from .model import VimureModel  # noqa: F401

__all__ = ["model", "masks", "sptensor", "utils", "VimureModel"]
