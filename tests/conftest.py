import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def golden(name):
    """Load one fixture generated from the compiled reference (tests/golden/make_golden.py)."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: z[k] for k in z.files}


def golden_names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = np.sqrt((b * b).sum())
    return float(np.sqrt(((a - b) ** 2).sum()) / (den if den > 0 else 1.0))


@pytest.fixture(scope="session")
def port():
    """The plain-C oracle restatement (oracle/pst_oracle.c), built on demand."""
    from oracle import port as p
    p.build()
    return p


@pytest.fixture(scope="session")
def ctx():
    """One GPU context for the whole -m gpu session."""
    from pyseistr_b200 import _lib
    c = _lib.default_context(0)
    yield c
