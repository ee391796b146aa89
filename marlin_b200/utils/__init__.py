from .mt_utils import MTUtils

__all__ = ["MTUtils"]
