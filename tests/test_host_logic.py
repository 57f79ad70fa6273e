"""CPU suite: host-side logic of the package (argument handling, synthetic generators)."""
import numpy as np
import pytest

import pyseistr_b200 as ps
from pyseistr_b200 import synth


def test_synth_deterministic_and_normalised():
    a = synth.cube(32, 10, 6, seed=3)
    b = synth.cube(32, 10, 6, seed=3)
    assert a.dtype == np.float32 and a.shape == (32, 10, 6) and a.flags.f_contiguous
    assert np.array_equal(a, b)
    assert not np.array_equal(a, synth.cube(32, 10, 6, seed=4))
    clean = synth.cube(32, 10, 6, seed=3, noise=0.0)
    assert abs(np.abs(clean).max() - 1.0) < 1e-6
    p = synth.cube(32, 10, 1, seed=3)
    assert p.shape == (32, 10)
    e = synth.erratic(a, ntraces=4)
    changed = np.any(e != a, axis=0).sum()
    assert changed == 4


def test_wrappers_validate_before_touching_the_gpu():
    with pytest.raises(ValueError):
        ps.dip3dc(np.zeros((8, 8), np.float32))
    with pytest.raises(ValueError):
        ps.dip2dc(np.zeros((8, 8, 2), np.float32))
    with pytest.raises(ValueError):
        ps.smoothc(np.zeros((8, 8, 2), np.float32), rect=[3, 3, 1], repeat=0)
    with pytest.raises(NotImplementedError):
        ps.soint2dc(np.zeros((8, 8), np.float32), np.ones((8, 8)), np.zeros((8, 8)), prec=1)


def test_public_names_match_reference_entry_points():
    # reference pyseistr/__init__.py:81-96 exports these C-variant names
    for n in ("dip3dc", "dip2dc", "somf3dc", "somean3dc", "somf2dc", "somean2dc", "smoothc", "soint3dc", "sint3dc"):
        assert callable(getattr(ps, n))


def test_signatures_match_reference_wrappers():
    """Argument names, order and defaults of the reference's *c wrappers (pyseistr/dip3d.py:59, dip2d.py:115,
    somf3d.py:54, somean3d.py:36, somf2d.py:60, somean2d.py:36, soint3d.py:65, sint.py:97, smooth.py:115); the only
    addition is the trailing ctx=None."""
    import inspect
    want = {
        "dip3dc": "(din, niter=5, liter=10, order=2, eps_dv=0.01, eps_cg=1, tol_cg=1e-06, rect=[5, 5, 5], verb=1, runc=1, mask=None, ctx=None)",
        "dip2dc": "(din, niter=5, liter=20, order=2, eps_dv=0.01, eps_cg=1, tol_cg=1e-06, rect=[10, 10, 1], verb=1, mask=None, ctx=None)",
        "somf3dc": "(dn, dipi, dipx, r1, r2, eps, order, option=1, verb=1, ctx=None)",
        "somean3dc": "(dn, dipi, dipx, r1, r2, eps, order, verb=0, ctx=None)",
        "somf2dc": "(dn, dip, ns, order, eps, option=1, verb=1, ctx=None)",
        "somean2dc": "(dn, dip, ns, order, eps, adj=0, verb=1, ctx=None)",
        "soint3dc": "(din, mask, dipi, dipx, order=1, niter=100, njs=[1, 1], drift=0, seed=202223, hasmask=1, var=0, verb=1, ctx=None)",
        "soint2dc": "(din, mask, dip, order=1, niter=100, njs=[1, 1], drift=0, hasmask=1, twoplane=0, prec=0, verb=1, ctx=None)",
        "sint3dc": "(din, mask, dipi, dipx, niter=100, eps=0.01, ns1=1, ns2=1, order1=1, order2=1, verb=1, ctx=None)",
        "smoothc": "(din, rect=[1, 1, 1], diff=[0, 0, 0], box=[0, 0, 0], repeat=1, adj=1, ctx=None)",
    }
    for name, sig in want.items():
        assert str(inspect.signature(getattr(ps, name))) == sig, (name, str(inspect.signature(getattr(ps, name))))
