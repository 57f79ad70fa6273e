"""CPU suite: the checkpoint + recompute smoother core (pyseistr_b200/csrc/pst_tri_rc_core.h) is the per-line code of
the CUDA kernels in pst_tri_rc.cu.  It is compiled for the host here (tests/native/tri_rc_host.cpp, no FMA
contraction, like the library) and must reproduce the oracle's ps_smooth2 bit for bit on every axis, radius, block
length, ragged length and in place."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("tri_rc") / "tri_rc_host.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", os.path.join(ROOT, "pyseistr_b200", "csrc"),
                    "-o", so, os.path.join(ROOT, "tests", "native", "tri_rc_host.cpp")], check=True)
    lib = ctypes.CDLL(so)
    lib.tri_rc_host.restype = ctypes.c_int
    lib.tri_rc_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 6
    return lib


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    """The CUDA kernels' own source run on the host: one thread per CUDA thread, a barrier per warp."""
    so = str(tmp_path_factory.mktemp("tri_rc_emul") / "tri_rc_emul.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-pthread", "-I",
                    os.path.join(ROOT, "pyseistr_b200", "csrc"), "-o", so,
                    os.path.join(ROOT, "tests", "native", "tri_rc_emul.cpp")], check=True)
    lib = ctypes.CDLL(so)
    lib.tri_rc_emul.restype = ctypes.c_int
    lib.tri_rc_emul.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 6
    lib.path = so
    return lib


@pytest.fixture(scope="module")
def port():
    from oracle import port as p
    p.build()
    return p


def _run(lib, x, axis, nb, rc, inplace):
    n1, n2, n3 = x.shape
    src = np.asfortranarray(x, dtype=np.float32).copy(order="F")
    dst = src if inplace else np.full_like(src, np.float32(7.0), order="F")
    rcode = lib.tri_rc_host(src.ctypes.data, dst.ctypes.data, n1, n2, n3, axis, nb, rc)
    assert rcode == 0, rcode
    return dst


@pytest.mark.parametrize("shape", [(70, 37, 9), (33, 64, 40), (12, 5, 131), (32, 32, 32), (100, 11, 64)])
def test_core_matches_oracle_every_axis(host, port, shape):
    rng = np.random.default_rng(sum(shape))
    x = np.asfortranarray(rng.standard_normal(shape).astype(np.float32))
    for axis in range(3):
        nx = shape[axis]
        for nb in (2, 3, 4, 5, 6, 7, 8, 10, 16):
            if nb > nx:
                continue
            rect = [1, 1, 1]
            rect[axis] = nb
            want = port.smooth3(x, rect)
            for rc in (16, 32):
                if rc == 16 and nb > 8:
                    continue
                for inplace in (False, True):
                    got = _run(host, x, axis, nb, rc, inplace)
                    assert np.array_equal(got.view(np.uint32), np.asfortranarray(want).view(np.uint32)), (shape, axis, nb, rc, inplace)


def test_core_edge_lengths(host, port):
    """nx == nb, nx just above / below a block multiple, lines of signed zeros."""
    rng = np.random.default_rng(3)
    for nx in (5, 6, 9, 10, 11, 22, 27, 31, 32, 33, 54, 59, 63, 64, 65, 96):
        x = np.asfortranarray(rng.standard_normal((nx, 3, 2)).astype(np.float32))
        x[:, 1, 0] = -0.0
        x[::2, 2, 1] = 0.0
        for nb in (2, 5, 8, 10, 16):
            if nb > nx:
                continue
            want = np.asfortranarray(port.smooth3(x, [nb, 1, 1]))
            for rc in (16, 32):
                if rc == 16 and nb > 8:
                    continue
                got = _run(host, x, 0, nb, rc, True)
                assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (nx, nb, rc)


def test_kernels_emulated_on_host_match_oracle(emul):
    """Strided and contiguous kernels with the launcher's own geometry (partial warps, ragged blocks, in place).  Run in
    a child process (tests/native/run_emul.py): a stuck warp barrier must not hang pytest."""
    import json
    import sys
    cases = []
    for shape in ((70, 37, 9), (12, 5, 131), (33, 64, 5)):
        for axis in range(3):
            for nb in (2, 5, 10, 16):
                if nb > shape[axis]:
                    continue
                for rc in (16, 32):
                    cases.append([list(shape), axis, nb, rc, int((nb + rc // 16 + axis) % 2 == 0)])
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "native", "run_emul.py"), emul.path, "tri_rc_emul",
                        json.dumps(cases)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
