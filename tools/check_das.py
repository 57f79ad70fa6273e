"""Config-1 check (2-D DAS-style panel, reference demos/test_pyseistr_das_massive.py:197-198):
dip2dc(d,2,10,2,0.01,1,1e-6,[40,40,1]) + somf2dc(d,dip,8,2,0.01) on a 3000x860 panel, GPU vs oracle."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyseistr_b200 as ps
from pyseistr_b200 import synth
from oracle import port
n1, n2 = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3000, 860)
d = synth.erratic(synth.cube(n1, n2, 1, seed=5), ntraces=20)
ctx = ps.default_context(0)
for rep in range(2):
    t = time.perf_counter(); p = ps.dip2dc(d, 2, 10, 2, 0.01, 1, 1e-6, [40, 40, 1], verb=0, ctx=ctx); t1 = time.perf_counter() - t
    t = time.perf_counter(); f = ps.somf2dc(d, p, 8, 2, 0.01, verb=0, ctx=ctx); t2 = time.perf_counter() - t
print(f"GPU  dip2d {t1*1e3:.1f} ms  somf2d {t2*1e3:.1f} ms  -> {n1*n2/(t1+t2)/1e6:.2f} Mvox/s (host buffers)")
t = time.perf_counter(); po = port.dip2dc(d, 2, 10, 2, rect=[40, 40, 1]); t3 = time.perf_counter() - t
t = time.perf_counter(); fo = port.somf2dc(d, po, 8, 2, 0.01); t4 = time.perf_counter() - t
rel = lambda a, b: float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-30))
print(f"CPU  dip2d {t3:.2f} s  somf2d {t4:.2f} s  -> {n1*n2/(t3+t4)/1e6:.3f} Mvox/s (1 core)")
print(f"dip2d rel-L2 {rel(p, po):.2e}; somf2d on oracle dips bit-exact: {np.array_equal(ps.somf2dc(d, po, 8, 2, 0.01, verb=0, ctx=ctx), fo)}")
