"""pyseistr_b200 — B200-native (sm_100a) implementation of pyseistr's structure-oriented
filtering hot path behind the reference's own ``*c`` entry points.

    from pyseistr_b200 import dip3dc, somf3dc, somean3dc, dip2dc, somf2dc, somean2dc

The compute path is hand-written CUDA in ``libpst_b200.so`` (C-ABI: include/pst_b200.h).
Importing this package does not need a GPU; calling an entry point does.
"""
from .api import (dip2dc, dip3dc, pwpaintc, rgt, sint2dc, sint3dc, smoothc, soint2dc, soint3dc, somean2dc, somean3dc, somf2dc, somf3dc)  # noqa: F401
from ._lib import Context, PstError, default_context  # noqa: F401

__all__ = ["dip3dc", "dip2dc", "somf3dc", "somean3dc", "somf2dc", "somean2dc", "smoothc", "soint3dc", "soint2dc", "sint3dc", "sint2dc", "pwpaintc", "rgt",
           "Context", "PstError", "default_context"]
