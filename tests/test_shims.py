"""The reference-named shim modules (pyseistr_b200/shims: dipcfun, sofcfun, sof3dcfun, soint3dcfun, soint2dcfun).

CPU: every function the reference's hot-path wrappers import exists with the reference's positional arity, and -- when
/root/reference is present -- the reference's OWN wrapper source (pyseistr/dip3d.py) runs against the shim and reaches the
C-ABI (which refuses without a GPU: no CPU fallback).  GPU: the shims return what the package's entry points return."""
import importlib
import importlib.util
import inspect
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

ARITY = {"dipcfun": {"dipc": 15, "smoothcf": 15}, "sof3dcfun": {"csomean3d": 11, "csomf3d": 13},
         "sofcfun": {"csomean2d": 10, "csomf2d": 11}, "soint3dcfun": {"csoint3d": 16, "csint3d": 14},
         "soint2dcfun": {"csoint2d": 15, "csint2d": 10}, "paint2dcfun": {"cpaint2d": 8, "cpaint3d": 8}}


@pytest.fixture(scope="module")
def shims():
    from pyseistr_b200 import shims as s
    s.activate()
    mods = {}
    for name in ARITY:
        sys.modules.pop(name, None)
        mods[name] = importlib.import_module(name)
        assert os.path.dirname(mods[name].__file__) == os.path.dirname(s.__file__), name
    return mods


def test_shim_modules_have_the_reference_signatures(shims):
    for mod, fns in ARITY.items():
        for fn, n in fns.items():
            f = getattr(shims[mod], fn)
            assert len(inspect.signature(f).parameters) == n, (mod, fn)


def test_reference_wrapper_source_runs_on_the_shim(shims):
    """pyseistr/dip3d.py of the reference, unmodified, imports dipcfun -> the shim -> pst_dip."""
    src = "/root/reference/pyseistr/dip3d.py"
    if not os.path.exists(src):
        pytest.skip("reference tree not present on this box")
    import types
    pkg = types.ModuleType("pyseistr_ref")            # the reference's package directory WITHOUT its __init__ (matplotlib)
    pkg.__path__ = [os.path.dirname(src)]
    sys.modules["pyseistr_ref"] = pkg
    mod = importlib.import_module("pyseistr_ref.dip3d")
    from pyseistr_b200 import _lib
    d = np.zeros((16, 4, 3), np.float32)
    if _lib.load().pst_device_count() > 0:
        di, dx = mod.dip3dc(d, verb=0)
        assert di.shape == (16, 4, 3) and dx.shape == (16, 4, 3)
    else:
        with pytest.raises(_lib.PstError) as e:
            mod.dip3dc(d, verb=0)
        assert e.value.code == -2


@pytest.mark.gpu
def test_shims_match_entry_points(shims):
    import pyseistr_b200 as ps
    from pyseistr_b200 import synth
    n1, n2, n3 = 48, 14, 6
    d = synth.erratic(synth.cube(n1, n2, n3, seed=31), ntraces=4)
    F = lambda a: np.float32(a).flatten(order="F")
    di, dx = ps.dip3dc(d, 2, 4, verb=0)
    flat = shims["dipcfun"].dipc(F(d), n1, n2, n3, 2, 4, 2, 0.01, 1, 1e-6, 5, 5, 5, 0, 0)
    assert np.array_equal(flat.reshape(n1, n2, n3, 2, order="F")[..., 0], di)
    f = shims["sof3dcfun"].csomf3d(F(d), F(di), F(dx), n1, n2, n3, 2, 2, 9, 1, 2, 0.01, 0)
    assert np.array_equal(f.reshape(n1, n2, n3, order="F"), ps.somf3dc(d, di, dx, 2, 2, 0.01, 2, verb=0))
    m = shims["sof3dcfun"].csomean3d(F(d), F(di), F(dx), n1, n2, n3, 1, 2, 1, 0.01, 0)
    assert np.array_equal(m.reshape(n1, n2, n3, order="F"), ps.somean3dc(d, di, dx, 1, 2, 0.01, 1))
    p2 = ps.dip2dc(d[:, :, 0], 2, 5, 2, rect=[5, 5, 1], verb=0)
    g = shims["sofcfun"].csomf2d(F(d[:, :, 0]), F(p2), n1, n2, 1, 3, 7, 1, 2, 0.01, 0)
    assert np.array_equal(g.reshape(n1, n2, order="F"), ps.somf2dc(d[:, :, 0], p2, 3, 2, 0.01, verb=0))
    s = shims["dipcfun"].smoothcf(F(d), n1, n2, n3, 1, 0, 3, 4, 2, 0, 0, 0, 0, 0, 0)
    assert np.array_equal(s.reshape(n1, n2, n3, order="F"), ps.smoothc(d, [3, 4, 2], adj=0))
    keep = np.random.default_rng(3).random((n2, n3)) > 0.5
    mask = np.zeros_like(d); mask[:, keep] = 1
    a = shims["soint3dcfun"].csoint3d(F(d * mask), F(mask), F(di), F(dx), n1, n2, n3, 1, 1, 1, 5, 0, 202223, 1, 0.0, 0)
    assert np.array_equal(a.reshape(n1, n2, n3, order="F"), ps.soint3dc(d * mask, mask, di, dx, order=1, niter=5, verb=0))
    b = shims["soint3dcfun"].csint3d(F(d * mask), F(di), F(dx), F(mask), n1, n2, n3, 3, 1, 1, 1, 1, 0, 0.01)
    assert np.array_equal(b.reshape(n1, n2, n3, order="F"), ps.sint3dc(d * mask, mask, di, dx, niter=3, verb=0))
