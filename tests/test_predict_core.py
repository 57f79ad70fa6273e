"""CPU suite: the per-trace core of the plane-wave prediction kernel (pyseistr_b200/csrc/pst_predict_core.h: ring-indexed
windows, interior / edge variants, running pointers) compiled for the host (tests/native/predict_host.cpp, no FMA
contraction, like the library) must reproduce the oracle's predict1_step / predict2_step bit for bit: both orders, both
directions, one and two parents, trace lengths around every unrolling boundary."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_fp = ctypes.POINTER(ctypes.c_float)


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("predict") / "predict_host.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", os.path.join(ROOT, "pyseistr_b200", "csrc"),
                    "-o", so, os.path.join(ROOT, "tests", "native", "predict_host.cpp")], check=True)
    lib = ctypes.CDLL(so)
    lib.predict_host.restype = ctypes.c_int
    lib.predict_host.argtypes = [_fp, _fp, _fp, _fp] + [ctypes.c_int] * 6 + [ctypes.c_float, _fp]
    lib.predict_warp_host.restype = ctypes.c_int
    lib.predict_warp_host.argtypes = lib.predict_host.argtypes
    return lib


@pytest.fixture(scope="module")
def port():
    from oracle import port as p
    p.build()
    return p


def _p(a):
    return a.ctypes.data_as(_fp)


@pytest.mark.parametrize("fn", ["predict_host", "predict_warp_host"])
@pytest.mark.parametrize("nw", [1, 2])
@pytest.mark.parametrize("two", [0, 1])
def test_core_matches_oracle(host, port, nw, two, fn):
    """predict_host: the thread-per-trace core (pst_predict_core.h); predict_warp_host: the warp-per-trace core
    (pst_predict_warp.h), lanes emulated phase by phase."""
    rng = np.random.default_rng(10 * nw + two)
    ntr = 7
    for n1 in (2 * nw + 2, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 19, 20, 21, 24, 25, 26, 30, 31, 32, 33, 34, 35, 50, 62, 63, 64, 65, 66, 67, 95, 96, 97, 101, 200):
        if n1 < 2 * nw + 2:
            continue
        for forw1, forw2 in ((0, 0), (1, 0), (0, 1), (1, 1)):
            x1 = rng.standard_normal((n1, ntr)).astype(np.float32)
            x2 = rng.standard_normal((n1, ntr)).astype(np.float32)
            g1 = rng.uniform(-1.2, 1.2, (n1, ntr)).astype(np.float32)
            g2 = rng.uniform(-1.2, 1.2, (n1, ntr)).astype(np.float32)
            x1[rng.random(x1.shape) < 0.1] = 0.0
            out = np.full((n1, ntr), 7.0, np.float32)
            rc = getattr(host, fn)(_p(x1), _p(g1), _p(x2), _p(g2), n1, ntr, nw, two, forw1, forw2, ctypes.c_float(1e-4), _p(out))
            assert rc == 0
            for t in range(ntr):
                if two:
                    want = port.predict(x1[:, t].copy(), g1[:, t].copy(), nw, forw1, 1e-4, x2[:, t].copy(), g2[:, t].copy(), forw2)
                else:
                    want = port.predict(x1[:, t].copy(), g1[:, t].copy(), nw, forw1, 1e-4)
                assert np.array_equal(out[:, t].view(np.uint32), np.asarray(want, np.float32).view(np.uint32)), (n1, nw, two, forw1, forw2, t)
