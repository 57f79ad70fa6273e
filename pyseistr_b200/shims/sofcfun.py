"""sofcfun: csomean2d (reference sof_cfuns.c:1433, "OOiiiiiifi") and csomf2d (:1534, "OOiiiiiiifi")."""
import numpy as np

from _common import check, ctx, f32, ptr

__all__ = ["csomean2d", "csomf2d"]


def csomean2d(dn, dip, n1, n2, n3, ns, order, adj, eps, verb):
    d, a = f32(dn), f32(dip)
    c = ctx()
    out = np.empty_like(d)
    check(c.lib.pst_somean2d(c.handle, ptr(d), ptr(a), int(n1), int(n2), int(n3), int(ns), int(order), int(adj), float(eps),
                             int(verb), ptr(out)))
    return out


def csomf2d(dn, dip, n1, n2, n3, ns, nmf, option, order, eps, verb):
    d, a = f32(dn), f32(dip)
    c = ctx()
    out = np.empty_like(d)
    check(c.lib.pst_somf2d(c.handle, ptr(d), ptr(a), int(n1), int(n2), int(n3), int(ns), int(nmf), int(option), int(order),
                           float(eps), int(verb), ptr(out)))
    return out
