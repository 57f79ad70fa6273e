"""CPU suite: the per-line core of the register kernels of the MULTI-GPU axis-3 smoothing pass
(pyseistr_b200/csrc/pst_tri3_reg_core.h) is the arithmetic of tri3_reg_fwd_kernel / tri3_reg_bwd_kernel in pst_dip.cu.
It is compiled for the host here (tests/native/tri3_reg_host.cpp, no FMA contraction, like the library) and run rank
after rank over the n3-slabs of one volume, in place: every rank role (first / interior / last), one and several
128- or 32-plane chunks per rank, every instantiated radius -- bit-identical to the oracle's ps_smooth2 along axis 3,
i.e. to the single-GPU result."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("tri3_reg") / "tri3_reg_host.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", os.path.join(ROOT, "pyseistr_b200", "csrc"),
                    "-o", so, os.path.join(ROOT, "tests", "native", "tri3_reg_host.cpp")], check=True)
    lib = ctypes.CDLL(so)
    lib.tri3_reg_host.restype = ctypes.c_int
    lib.tri3_reg_host.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 5
    return lib


@pytest.fixture(scope="module")
def port():
    from oracle import port as p
    p.build()
    return p


def _run(lib, x, nb, nranks):
    n1, n2, n3 = x.shape
    y = np.asfortranarray(x, dtype=np.float32).copy(order="F")
    rc = lib.tri3_reg_host(y.ctypes.data, n1, n2, n3, nb, nranks)
    assert rc == 0, rc
    return y


@pytest.mark.parametrize("shape,nranks", [((5, 7, 64), 2), ((4, 5, 96), 3), ((3, 9, 256), 8), ((6, 4, 224), 7),   # one 32-plane chunk per rank
                                          ((3, 5, 256), 2), ((2, 3, 1024), 8), ((3, 3, 384), 3),                  # one 128-plane chunk
                                          ((4, 3, 192), 3), ((2, 5, 192), 2), ((3, 2, 128), 2),                   # 2 / 3 / 2 chunks of 32
                                          ((2, 3, 1024), 4), ((2, 2, 1024), 2), ((3, 2, 768), 2)])                # 2 / 4 / 3 chunks of 128
def test_multi_rank_walks_match_oracle(host, port, shape, nranks):
    rng = np.random.default_rng(sum(shape) + nranks)
    x = np.asfortranarray(rng.standard_normal(shape).astype(np.float32))
    for nb in (2, 3, 4, 5, 6, 8):
        if shape[2] // nranks >= 128 or 2 * nb <= 32:
            got = _run(host, x, nb, nranks)
            want = port.smooth3(x, (1, 1, nb)).reshape(x.shape, order="F")
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (shape, nranks, nb)


def test_signed_zeros_and_constant_lines(host, port):
    """Exact zeros and negative zeros in the input (the sign of a zero must not leak through a carry), whole zero lines."""
    rng = np.random.default_rng(3)
    x = rng.standard_normal((4, 6, 256)).astype(np.float32)
    x[rng.random(x.shape) < 0.3] = 0.0
    x[rng.random(x.shape) < 0.1] = -0.0
    x[1, 2, :] = 0.0
    x[2, 3, :] = -0.0
    x = np.asfortranarray(x)
    for nranks in (2, 8):
        for nb in (2, 5, 8):
            got = _run(host, x, nb, nranks)
            want = port.smooth3(x, (1, 1, nb)).reshape(x.shape, order="F")
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (nranks, nb)


def test_geometries_outside_the_kernels_are_refused(host):
    x = np.zeros((2, 2, 100), np.float32, order="F")
    assert host.tri3_reg_host(x.ctypes.data, 2, 2, 100, 5, 2) == -1      # 50-plane slabs
    x = np.zeros((2, 2, 64), np.float32, order="F")
    assert host.tri3_reg_host(x.ctypes.data, 2, 2, 64, 5, 1) == -1       # one rank: the single-GPU smoothers
    assert host.tri3_reg_host(x.ctypes.data, 2, 2, 64, 7, 2) == -3       # radius not instantiated
