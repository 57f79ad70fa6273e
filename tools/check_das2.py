import os, sys, ctypes
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyseistr_b200 as ps
from pyseistr_b200 import synth, _lib
from oracle import port
n1, n2 = 3000, 860
ctx = ps.default_context(0)
x = synth.cube(n1, n2, 1, seed=6)
a = ps.smoothc(x, rect=[40, 40, 1], ctx=ctx); b = port.smooth3(x, [40, 40, 1]).reshape(x.shape, order="F")
print("smooth (40,40) bit-exact:", np.array_equal(a, b), float(np.abs(a - b).max()))
a = ps.smoothc(x, rect=[40, 1, 1], ctx=ctx); b = port.smooth3(x, [40, 1, 1]).reshape(x.shape, order="F")
print("smooth axis1 only:", np.array_equal(a, b))
a = ps.smoothc(x, rect=[1, 40, 1], ctx=ctx); b = port.smooth3(x, [1, 40, 1]).reshape(x.shape, order="F")
print("smooth axis2 only:", np.array_equal(a, b))
rel = lambda a, b: float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-30))
for (niter, liter) in [(1, 1), (1, 3), (1, 10), (2, 10)]:
    p = ps.dip2dc(x, niter, liter, 2, 0.01, 1, 1e-6, [40, 40, 1], verb=1, ctx=ctx)
    po = port.dip2dc(x, niter, liter, 2, rect=[40, 40, 1])
    print(f"dip2d niter={niter} liter={liter}: rel-L2 {rel(p, po):.2e} max|d| {np.abs(p-po).max():.2e}")
