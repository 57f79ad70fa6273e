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


def test_public_names_match_reference_entry_points():
    # reference pyseistr/__init__.py:81-96 exports these C-variant names
    for n in ("dip3dc", "dip2dc", "somf3dc", "somean3dc", "somf2dc", "somean2dc", "smoothc"):
        assert callable(getattr(ps, n))
