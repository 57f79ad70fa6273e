"""Micro-driver for profiling: one shaping-operator application (3 smoothing passes) on a
device-resident volume.  Usage: python tools/prof_smooth.py n1 n2 n3 [r1 r2 r3] [reps]"""
import ctypes
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyseistr_b200 as ps  # noqa: E402
from pyseistr_b200 import _lib  # noqa: E402

a = [int(v) for v in sys.argv[1:]]
n1, n2, n3 = a[:3]
r = a[3:6] if len(a) >= 6 else [5, 5, 5]
reps = a[6] if len(a) >= 7 else 3
ctx = ps.Context(0)
N = n1 * n2 * n3
x = np.random.default_rng(0).standard_normal(N, dtype=np.float32)
d = ctx.alloc(4 * N)
ctx.h2d(d, x)
ctx.set_profile(True)
for it in range(reps):
    ctx.reset_stats()
    t = time.perf_counter()
    _lib.check(ctx.lib.pst_smooth3_dev(ctx.handle, d, n1, n2, n3, *r, 1, 0))
    dt = time.perf_counter() - t
    st = ctx.stats()
    ms = dict(zip(_lib.KERNEL_CLASSES, st["class_ms"]))
    print(f"rep {it}: wall {dt*1e3:.2f} ms; " + ", ".join(
        f"{k} {v:.3f} ms ({8*N/v/1e6:.0f} GB/s)" for k, v in ms.items() if v > 0))
