"""soint2dcfun: csoint2d (reference soint2d_cfuns.c:2260, "OOOOiiiiiiiiiii"; the default path: one slope field, no
preconditioner, no drift) and csint2d (:2420, "OOOiiiiiif").  Both are their 3-D counterparts on an (n1, n2, 1) volume,
bit for bit on the compiled reference (CPU test suite)."""
import numpy as np

from _common import check, ctx, f32, ptr

__all__ = ["csoint2d", "csint2d"]


def csoint2d(din, mask, dip1, dip2, n1, n2, nw, nj1, nj2, niter, drift, hasmask, twoplane, prec, verb):
    if twoplane and not prec:
        # the reference's two-plane solver call without preconditioner is commented out (soint2d_cfuns.c:2354-2356,
        # :2389-2391): the input comes back unchanged
        ctx()                           # (raises without the library / a device, like every other entry point)
        return np.array(f32(din), copy=True)
    if twoplane or prec or drift:
        raise NotImplementedError("csoint2d on GPU: prec=0, drift=0 only")
    d, a = f32(din), f32(dip1)
    m = f32(mask) if hasmask else None
    c = ctx()
    out = np.empty_like(d)
    check(c.lib.pst_soint3d(c.handle, ptr(d), ptr(m) if m is not None else None, ptr(a), ptr(a), int(n1), int(n2), 1,
                            int(nw), int(nj1), int(nj2), int(niter), 0, 202223, int(hasmask), 0.0, int(verb), ptr(out)))
    return out


def csint2d(din, dip, mask, n1, n2, niter, ns, nw, verb, eps):
    d, a, m = f32(din), f32(dip), f32(mask)
    z = np.zeros_like(d)
    c = ctx()
    out = np.empty_like(d)
    check(c.lib.pst_sint3d(c.handle, ptr(d), ptr(a), ptr(z), ptr(m), int(n1), int(n2), 1, int(niter), int(ns), 0, int(nw), int(nw),
                           int(verb), float(eps), ptr(out)))
    return out
