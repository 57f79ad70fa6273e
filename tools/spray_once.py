"""One somf3dc call on a 1000x256x96 cube: the target of the ncu capture of the prediction / slot-median kernels."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyseistr_b200 as ps  # noqa: E402
from pyseistr_b200 import synth  # noqa: E402

n1, n2, n3 = 1000, 256, 96
d = synth.cube(n1, n2, n3, seed=3)
di, dx = synth.smooth_dips(n1, n2, n3, seed=3)
ctx = ps.default_context(0)
ps.somf3dc(d, di, dx, 2, 2, 0.01, 2, verb=0, ctx=ctx)
