"""Child process of tests/test_tri_sys_emul.py and tests/test_tri_rc_core.py: runs emulated kernels (a deadlocked
emulation aborts, which must not take the pytest process with it) and compares with the oracle bit for bit.

    python run_emul.py <lib.so> <entry> '<json list of [shape, axis, nb, param, inplace]>'

Prints one line per case; exit status 0 when every case matched."""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import port  # noqa: E402


def main():
    lib = ctypes.CDLL(sys.argv[1])
    fn = getattr(lib, sys.argv[2])
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 6
    port.build()
    bad = 0
    for shape, axis, nb, param, inplace in json.loads(sys.argv[3]):
        rng = np.random.default_rng(sum(shape) + nb + axis)
        x = np.asfortranarray(rng.standard_normal(shape).astype(np.float32))
        rect = [1, 1, 1]
        rect[axis] = nb
        want = np.asfortranarray(port.smooth3(x, rect))
        src = x.copy(order="F")
        dst = src if inplace else np.full_like(src, np.float32(7.0), order="F")
        rc = fn(src.ctypes.data, dst.ctypes.data, *shape, axis, nb, param)
        ok = rc == 0 and np.array_equal(dst.view(np.uint32), want.view(np.uint32))
        print("OK" if ok else "BAD", shape, axis, nb, param, inplace, "rc", rc, flush=True)
        bad += not ok
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
