"""dipcfun: dipc (reference pyseistr/src/dip_cfuns.c:1694, "Oiiiiiifffiiiii") and smoothcf (:2006, "Oiiiiiiiiiiiiii")."""
import numpy as np

from _common import check, ctx, f32, ptr

__all__ = ["dipc", "smoothcf"]


def dipc(din, n1, n2, n3, niter, liter, order, eps_dv, eps_cg, tol_cg, r1, r2, r3, hasmask, verb):
    n = int(n1) * int(n2) * int(n3)
    d = f32(din)                                       # data, followed by the mask when hasmask (dip3d.py:106)
    c = ctx()
    out = np.empty(n if int(n3) == 1 else 2 * n, np.float32)
    mask = ptr(d[n:]) if hasmask else None
    check(c.lib.pst_dip(c.handle, ptr(d), mask, int(n1), int(n2), int(n3), int(niter), int(liter), int(order),
                        float(eps_dv), float(eps_cg), float(tol_cg), int(r1), int(r2), int(r3), int(verb), ptr(out)))
    return out


def smoothcf(din, n1, n2, n3, repeat, adj, r1, r2, r3, diff1, diff2, diff3, box1, box2, box3):
    d = f32(din)
    c = ctx()
    out = np.empty_like(d)
    check(c.lib.pst_smoothcf(c.handle, ptr(d), int(n1), int(n2), int(n3), int(repeat), int(adj), int(r1), int(r2), int(r3),
                             int(diff1), int(diff2), int(diff3), int(box1), int(box2), int(box3), ptr(out)))
    return out
