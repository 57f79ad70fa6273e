"""sof3dcfun: csomean3d (reference sof3d_cfuns.c:1355, "OOOiiiiiifi") and csomf3d (:1554, "OOOiiiiiiiifi")."""
import numpy as np

from _common import check, ctx, f32, ptr

__all__ = ["csomean3d", "csomf3d"]


def csomean3d(dn, dipi, dipx, n1, n2, n3, r1, r2, order, eps, verb):
    d, a, b = f32(dn), f32(dipi), f32(dipx)
    c = ctx()
    out = np.empty_like(d)
    check(c.lib.pst_somean3d(c.handle, ptr(d), ptr(a), ptr(b), int(n1), int(n2), int(n3), int(r1), int(r2), int(order),
                             float(eps), int(verb), ptr(out)))
    return out


def csomf3d(dn, dipi, dipx, n1, n2, n3, r1, r2, rmf, option, order, eps, verb):
    d, a, b = f32(dn), f32(dipi), f32(dipx)
    c = ctx()
    out = np.empty_like(d)
    check(c.lib.pst_somf3d(c.handle, ptr(d), ptr(a), ptr(b), int(n1), int(n2), int(n3), int(r1), int(r2), int(rmf),
                           int(option), int(order), float(eps), int(verb), ptr(out)))
    return out
