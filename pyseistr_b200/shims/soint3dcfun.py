"""soint3dcfun: csoint3d (reference soint3d_cfuns.c:2405, "OOOOiiiiiiiiiifi") and csint3d (:2510, "OOOOiiiiiiiiif")."""
import numpy as np

from _common import check, ctx, f32, ptr

__all__ = ["csoint3d", "csint3d"]


def csoint3d(din, mask, dipi, dipx, n1, n2, n3, nw, nj1, nj2, niter, drift, seed, hasmask, var, verb):
    d, a, b = f32(din), f32(dipi), f32(dipx)
    m = f32(mask) if hasmask else None
    c = ctx()
    out = np.empty_like(d)
    check(c.lib.pst_soint3d(c.handle, ptr(d), ptr(m) if m is not None else None, ptr(a), ptr(b), int(n1), int(n2), int(n3),
                            int(nw), int(nj1), int(nj2), int(niter), int(drift), int(seed), int(hasmask), float(var),
                            int(verb), ptr(out)))
    return out


def csint3d(din, dipi, dipx, mask, n1, n2, n3, niter, ns1, ns2, order1, order2, verb, eps):
    d, a, b, m = f32(din), f32(dipi), f32(dipx), f32(mask)
    c = ctx()
    out = np.empty_like(d)
    check(c.lib.pst_sint3d(c.handle, ptr(d), ptr(a), ptr(b), ptr(m), int(n1), int(n2), int(n3), int(niter), int(ns1), int(ns2),
                           int(order1), int(order2), int(verb), float(eps), ptr(out)))
    return out
