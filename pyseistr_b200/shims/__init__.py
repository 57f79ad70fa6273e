"""Drop-in replacements of the reference's CPython extension modules (same module and function names, same positional
arguments, same 1-D float32 return arrays), backed by libpst_b200 on B200.

Put this directory in front of ``sys.path`` (``pyseistr_b200.shims.activate()``) and the reference's own Python wrappers
(pyseistr/dip3d.py, somf3d.py, somean3d.py, somf2d.py, somean2d.py, soint3d.py, soint2d.py, sint.py, smooth.py) run
unmodified on the GPU: their ``from dipcfun import *`` etc. resolve to these modules.  Only the functions of the hot
path are provided; anything else raises AttributeError like a missing symbol would.
"""
import os
import sys


def activate():
    """Make ``import dipcfun`` / ``sofcfun`` / ``sof3dcfun`` / ``soint3dcfun`` / ``soint2dcfun`` resolve to the shims."""
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    return here
