"""paint2dcfun: cpaint2d / cpaint3d (reference paint_cfuns.c:1861-2024, "OOiiiifi"; the two entries are the same code)."""
import numpy as np

from _common import check, ctx, f32, ptr

__all__ = ["cpaint2d", "cpaint3d"]


def cpaint2d(dip, trace, n1, n2, order, i0, eps, verb):
    d, t = f32(dip), f32(trace)
    c = ctx()
    out = np.empty_like(d)
    check(c.lib.pst_paint2d(c.handle, ptr(d), ptr(t), int(n1), int(n2), int(order), int(i0), float(eps), int(verb), ptr(out)))
    return out


cpaint3d = cpaint2d
