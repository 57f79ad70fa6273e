"""CPU suite: the per-line core of the L2-resident checkpoint + recompute smoother (pyseistr_b200/csrc/pst_tri_l2_core.h)
is the arithmetic of the CUDA kernel in pst_tri_l2.cu.  It is compiled for the host here (tests/native/tri_l2_host.cpp,
no FMA contraction, like the library) and must reproduce the oracle's ps_smooth2 bit for bit on every axis, radius,
ragged length, signed zeros and in place -- while consuming exactly the block stream the kernel's TMA ring issues."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("tri_l2") / "tri_l2_host.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", os.path.join(ROOT, "pyseistr_b200", "csrc"),
                    "-o", so, os.path.join(ROOT, "tests", "native", "tri_l2_host.cpp")], check=True)
    lib = ctypes.CDLL(so)
    lib.tri_l2_host.restype = ctypes.c_int
    lib.tri_l2_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 5
    return lib


@pytest.fixture(scope="module")
def port():
    from oracle import port as p
    p.build()
    return p


def _run(lib, x, axis, nb, inplace):
    n1, n2, n3 = x.shape
    src = np.asfortranarray(x, dtype=np.float32).copy(order="F")
    dst = src if inplace else np.full_like(src, np.float32(7.0), order="F")
    rcode = lib.tri_l2_host(src.ctypes.data, dst.ctypes.data, n1, n2, n3, axis, nb)
    assert rcode == 0, rcode
    return dst


def _want(port, x, axis, nb):
    rect = [1, 1, 1]
    rect[axis] = nb
    return port.smooth3(x, rect)


@pytest.mark.parametrize("shape", [(70, 37, 9), (33, 64, 40), (12, 5, 131), (32, 32, 32), (100, 11, 64), (1000, 3, 2), (1024, 2, 3)])
def test_core_matches_oracle_every_axis(host, port, shape):
    rng = np.random.default_rng(sum(shape))
    x = np.asfortranarray(rng.standard_normal(shape).astype(np.float32))
    for axis in range(3):
        for nb in (2, 3, 5, 8, 10, 16):
            if nb > shape[axis]:
                continue
            for inplace in (False, True):
                got = _run(host, x, axis, nb, inplace)
                assert np.array_equal(got, _want(port, x, axis, nb)), (shape, axis, nb, inplace)


@pytest.mark.parametrize("nx", [5, 6, 10, 11, 22, 31, 32, 33, 54, 63, 64, 65, 96, 97, 128])
def test_core_ragged_lengths_and_signed_zeros(host, port, nx):
    """Lengths around the block boundaries (both fold zones in one block, a top block of pure padding, ...), with
    exact zeros and negative zeros in the input (the sign of a zero must not leak into the sums)."""
    rng = np.random.default_rng(nx)
    x = rng.standard_normal((nx, 3, 2)).astype(np.float32)
    x[rng.random(x.shape) < 0.2] = 0.0
    x[rng.random(x.shape) < 0.1] = -0.0
    x = np.asfortranarray(x)
    for nb in (2, 3, 4, 5, 6, 7, 8, 10, 16):
        if nb > nx:
            continue
        got = _run(host, x, 0, nb, True)
        want = _want(port, x, 0, nb)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (nx, nb)
