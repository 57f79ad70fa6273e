"""Drop-in Python entry points of the reference's C variants, running on B200.

Same names, argument order, defaults, return layouts and quirks as the reference's
``*c`` functions (re-exported by reference pyseistr/__init__.py:81-96); each flattens to
float32 Fortran order exactly as the reference wrapper does and calls the C-ABI
(include/pst_b200.h) through ctypes.  No CPU fallback: without the CUDA library or a
B200 these raise.
"""
import ctypes

import numpy as np

from . import _lib

_fp = ctypes.POINTER(ctypes.c_float)


def _F(a):
    """float32 1-D array in Fortran order (the reference's ``np.float32(x).flatten(order='F')``).  An array that already
    is float32 and Fortran-contiguous is passed as a VIEW (no host copy)."""
    a = np.asarray(a)
    if a.dtype == np.float32 and a.flags.f_contiguous:
        return a.reshape(-1, order="F")
    return np.ascontiguousarray(np.float32(a).flatten(order="F"))


def _p(a):
    return a.ctypes.data_as(_fp)


def _shape3(d):
    if d.ndim == 2:
        return d.shape[0], d.shape[1], 1
    if d.ndim != 3:
        raise ValueError("expected a 2-D or 3-D array")
    return d.shape


def _ctx(ctx):
    return ctx if ctx is not None else _lib.default_context()


def dip3dc(din, niter=5, liter=10, order=2, eps_dv=0.01, eps_cg=1, tol_cg=0.000001,
           rect=[5, 5, 5], verb=1, runc=1, mask=None, ctx=None):
    """3-D local slopes by shaping-regularised PWD (reference pyseistr/dip3d.py:59-116 ->
    dipcfun.dipc, dip_cfuns.c:1694).  Returns (dip_i, dip_x), float32 (n1,n2,n3) views of one
    F-ordered (n1,n2,n3,2) array, like the reference.  eps_dv, eps_cg, tol_cg are accepted and
    ignored exactly as the reference's C ignores them."""
    din = np.asarray(din)
    if din.ndim != 3:
        raise ValueError("dip3dc expects a 3-D array (n1,n2,n3)")
    n1, n2, n3 = din.shape
    c = _ctx(ctx)
    d = _F(din)
    m = None
    if mask is not None:
        mask = np.asarray(mask)
        if mask.size != din.size:
            raise ValueError("Mask and Data should have the same dimension")
        m = _F(mask)
    out = np.empty(2 * d.size if n3 != 1 else d.size, np.float32)
    _lib.check(c.lib.pst_dip(c.handle, _p(d), _p(m) if m is not None else None, n1, n2, n3,
                             int(niter), int(liter), int(order), float(eps_dv), float(eps_cg),
                             float(tol_cg), int(rect[0]), int(rect[1]), int(rect[2]), int(verb),
                             _p(out)))
    if n3 == 1:
        # the reference's reshape(n1,n2,n3,2) fails for n3 == 1 (dipc returns N floats); we
        # return the inline dip and a zero xline dip instead of raising.
        dip = out.reshape(n1, n2, 1, order="F")
        return dip, np.zeros_like(dip)
    dip = out.reshape(n1, n2, n3, 2, order="F")
    return dip[:, :, :, 0], dip[:, :, :, 1]


def dip2dc(din, niter=5, liter=20, order=2, eps_dv=0.01, eps_cg=1, tol_cg=0.000001,
           rect=[10, 10, 1], verb=1, mask=None, ctx=None):
    """2-D local slope (reference pyseistr/dip2d.py:115-221: the same dipc entry with n3=1).
    Returns float32 (n1,n2), F-ordered."""
    din = np.asarray(din)
    if din.ndim != 2:
        raise ValueError("dip2dc expects a 2-D array (n1,n2)")
    n1, n2 = din.shape
    c = _ctx(ctx)
    d = _F(din)
    m = None
    if mask is not None:
        mask = np.asarray(mask)
        if mask.size != din.size:
            raise ValueError("Mask and Data should have the same dimension")
        m = _F(mask)
    out = np.empty(d.size, np.float32)
    _lib.check(c.lib.pst_dip(c.handle, _p(d), _p(m) if m is not None else None, n1, n2, 1,
                             int(niter), int(liter), int(order), float(eps_dv), float(eps_cg),
                             float(tol_cg), int(rect[0]), int(rect[1]), int(rect[2]), int(verb),
                             _p(out)))
    return out.reshape(n1, n2, order="F")


def somean3dc(dn, dipi, dipx, r1, r2, eps, order, verb=0, ctx=None):
    """3-D structure-oriented mean (reference pyseistr/somean3d.py:36-73 -> csomean3d,
    sof3d_cfuns.c:1355).  eps is ignored like the reference (overridden by 0.01);
    slots falling outside the cube count as zeros and the divisor is always np (Q4)."""
    dn = np.asarray(dn)
    n1, n2, n3 = _shape3(dn)
    c = _ctx(ctx)
    d, a, b = _F(dn), _F(dipi), _F(dipx)
    if a.size != d.size or b.size != d.size:
        raise ValueError("data and slope volumes must have the same size")
    out = np.empty_like(d)
    _lib.check(c.lib.pst_somean3d(c.handle, _p(d), _p(a), _p(b), n1, n2, n3, int(r1), int(r2),
                                  int(order), float(eps), int(verb), _p(out)))
    return out.reshape([n1, n2, n3], order="F")


def somf3dc(dn, dipi, dipx, r1, r2, eps, order, option=1, verb=1, ctx=None):
    """3-D structure-oriented median (reference pyseistr/somf3d.py:54-97 -> csomf3d,
    sof3d_cfuns.c:1554).  The median runs over nmf = 2*r1*r2+1 slots along the flattened slot
    axis around the centre slot.  option=1: median filter (MF); option=2: space-varying median filter
    (SVMF, sof3d_cfuns.c:1254-1352) in the defined-behaviour variant of the reference's one-row over-read (DESIGN.md
    section 1; needs nmf >= 5)."""
    dn = np.asarray(dn)
    n1, n2, n3 = _shape3(dn)
    c = _ctx(ctx)
    d, a, b = _F(dn), _F(dipi), _F(dipx)
    if a.size != d.size or b.size != d.size:
        raise ValueError("data and slope volumes must have the same size")
    out = np.empty_like(d)
    rmf = 2 * int(r1) * int(r2) + 1
    _lib.check(c.lib.pst_somf3d(c.handle, _p(d), _p(a), _p(b), n1, n2, n3, int(r1), int(r2), rmf,
                                int(option), int(order), float(eps), int(verb), _p(out)))
    return out.reshape([n1, n2, n3], order="F")


def somean2dc(dn, dip, ns, order, eps, adj=0, verb=1, ctx=None):
    """2-D structure-oriented (triangle-weighted, normalised) smoothing (reference
    pyseistr/somean2d.py:36-74 -> csomean2d, sof_cfuns.c:1433).  eps is honoured."""
    dn = np.asarray(dn)
    n1, n2, n3 = _shape3(dn)
    c = _ctx(ctx)
    d, a = _F(dn), _F(dip)
    if a.size != d.size:
        raise ValueError("data and slope must have the same size")
    out = np.empty_like(d)
    _lib.check(c.lib.pst_somean2d(c.handle, _p(d), _p(a), n1, n2, n3, int(ns), int(order), int(adj),
                                  float(eps), int(verb), _p(out)))
    return np.squeeze(out.reshape(n1, n2, n3, order="F"))


def somf2dc(dn, dip, ns, order, eps, option=1, verb=1, ctx=None):
    """2-D structure-oriented median over the 2*ns+1 sprayed slots (reference
    pyseistr/somf2d.py:60-105 -> csomf2d, sof_cfuns.c:1534).  option=1: MF, option=2: SVMF (see somf3dc)."""
    dn = np.asarray(dn)
    n1, n2, n3 = _shape3(dn)
    c = _ctx(ctx)
    d, a = _F(dn), _F(dip)
    if a.size != d.size:
        raise ValueError("data and slope must have the same size")
    out = np.empty_like(d)
    _lib.check(c.lib.pst_somf2d(c.handle, _p(d), _p(a), n1, n2, n3, int(ns), 2 * int(ns) + 1,
                                int(option), int(order), float(eps), int(verb), _p(out)))
    return np.squeeze(out.reshape(n1, n2, n3, order="F"))


def soint3dc(din, mask, dipi, dipx, order=1, niter=100, njs=[1, 1], drift=0, seed=202223, hasmask=1, var=0,
             verb=1, ctx=None):
    """3-D structure-oriented interpolation: CG on the inline + xline PWD residual with the known
    samples held fixed (reference pyseistr/soint3d.py:65-108 -> csoint3d, soint3d_cfuns.c:2405).
    GPU path: any njs >= 1, drift=0 (drift != 0 raises PST_EUNSUP).  var > 0 adds the reference's
    MT19937(seed) + Box-Muller noise to the right-hand side (drawn on the host, same stream of numbers)."""
    din = np.asarray(din)
    n1, n2, n3 = _shape3(din)
    c = _ctx(ctx)
    d, a, b = _F(din), _F(dipi), _F(dipx)
    m = _F(mask) if mask is not None else None
    if a.size != d.size or b.size != d.size or (m is not None and m.size != d.size):
        raise ValueError("data, mask and slope volumes must have the same size")
    if hasmask and m is None:
        raise ValueError("hasmask=1 needs a mask")
    out = np.empty_like(d)
    _lib.check(c.lib.pst_soint3d(c.handle, _p(d), _p(m) if m is not None else None, _p(a), _p(b), n1, n2, n3,
                                 int(order), int(njs[0]), int(njs[1]), int(niter), int(drift), int(seed),
                                 int(hasmask), float(var), int(verb), _p(out)))
    return out.reshape(n1, n2, n3, order="F")


def soint2dc(din, mask, dip, order=1, niter=100, njs=[1, 1], drift=0, hasmask=1, twoplane=0, prec=0, verb=1, ctx=None):
    """2-D structure-oriented interpolation (reference pyseistr/soint2d.py:92-141 -> csoint2d, soint2d_cfuns.c:2260).
    The default path of csoint2d (one slope field, no preconditioner: ps_solver on allpass21_lop :296-336) is the
    inline half of allpass3_lop, and is bit-identical to csoint3d on an (n1, n2, 1) volume -- checked on the compiled
    reference by the CPU test suite -- so it runs through pst_soint3d.
    twoplane=1, prec=0: the reference's solver call for two slope fields without preconditioner is commented out
    (soint2d_cfuns.c:2354-2356, :2389-2391), so csoint2d hands the input back unchanged; so does this (quirk Q7, pinned on
    the compiled reference by the CPU test suite).  prec=1 (ps_solver_prec over predict_lop / predict2_lop) and drift raise."""
    din = np.asarray(din)
    if din.ndim != 2:
        raise ValueError("soint2dc expects a 2-D panel")
    n1, n2 = din.shape
    if twoplane and not prec:
        dip = np.asarray(dip)
        if dip.ndim != 3 or dip.shape != (n1, n2, 2):
            raise ValueError("soint2dc(twoplane=1) expects the two slope fields as an (n1, n2, 2) array")
        if np.asarray(mask).size != din.size:
            raise ValueError("data and mask must have the same size")
        _ctx(ctx)                       # no CPU-only mode anywhere in this package: without the library / a device this raises too
        return np.array(din, dtype=np.float32, order="F", copy=True)
    if twoplane or prec or drift:
        raise NotImplementedError("soint2dc on GPU: prec=0, drift=0 only")
    slope = np.float32(dip).reshape(n1, n2, 1)
    m3 = np.float32(mask).reshape(n1, n2, 1) if mask is not None else None
    out = soint3dc(din.reshape(n1, n2, 1), m3, slope, slope, order=order, niter=niter, njs=njs, drift=0, hasmask=hasmask,
                   var=0, verb=verb, ctx=ctx)
    return out.reshape(n1, n2)


def sint3dc(din, mask, dipi, dipx, niter=100, eps=0.01, ns1=1, ns2=1, order1=1, order2=1, verb=1, ctx=None):
    """3-D structure-oriented interpolation by shaping-regularised CG, the plane-wave smoother
    (inline then xline, radii ns1/ns2, PWD orders order1/order2) being the shaping operator
    (reference pyseistr/sint.py:97-131 -> csint3d, soint3d_cfuns.c:2510-2640)."""
    din = np.asarray(din)
    n1, n2, n3 = _shape3(din)
    c = _ctx(ctx)
    d, a, b, m = _F(din), _F(dipi), _F(dipx), _F(mask)
    if a.size != d.size or b.size != d.size or m.size != d.size:
        raise ValueError("data, mask and slope volumes must have the same size")
    out = np.empty_like(d)
    _lib.check(c.lib.pst_sint3d(c.handle, _p(d), _p(a), _p(b), _p(m), n1, n2, n3, int(niter), int(ns1), int(ns2),
                                int(order1), int(order2), int(verb), float(eps), _p(out)))
    return out.reshape(n1, n2, n3, order="F")


def sint2dc(din, mask, dip, niter=100, eps=0.01, ns=1, order=1, verb=1, ctx=None):
    """2-D interpolation of sparse data by shaping-regularised CG with the 2-D plane-wave smoother (reference
    pyseistr/sint.py:61-94 -> csint2d, soint2d_cfuns.c).  csint2d is csint3d on an (n1, n2, 1) volume with no xline
    smoothing (ns2 = 0, zero xline slope), bit for bit on the compiled reference (CPU test suite:
    test_ref_sint2d_is_sint3d_with_one_plane), so it runs through pst_sint3d."""
    din = np.asarray(din)
    if din.ndim != 2:
        raise ValueError("sint2dc expects a 2-D panel")
    n1, n2 = din.shape
    r3 = lambda a: np.float32(a).reshape(n1, n2, 1)
    out = sint3dc(r3(din), r3(mask), r3(dip), np.zeros((n1, n2, 1), np.float32), niter=niter, eps=eps, ns1=ns, ns2=0,
                  order1=order, order2=order, verb=verb, ctx=ctx)
    return out.reshape(n1, n2)


def pwpaintc(dip, trace, order=1, i0=0, eps=0.01, verb=False, ctx=None):
    """Plane-wave painting: the seed `trace` placed at trace i0 is spread along the slope field `dip` (n1, n2) by
    successive plane-wave predictions (reference pyseistr/rgt.py:pwpaintc -> paint2dcfun.cpaint2d, paint_cfuns.c:1861).
    Returns float32 (n1, n2)."""
    dip = np.asarray(dip)
    if dip.ndim != 2:
        raise ValueError("pwpaintc expects a 2-D slope field (n1, n2)")
    n1, n2 = dip.shape
    c = _ctx(ctx)
    d, t = _F(dip), _F(trace)
    if t.size != n1:
        raise ValueError("the seed trace must have n1 samples")
    out = np.empty_like(d)
    _lib.check(c.lib.pst_paint2d(c.handle, _p(d), _p(t), n1, n2, int(order), int(i0), float(eps), int(bool(verb)), _p(out)))
    return out.reshape(n1, n2, order="F")


def rgt(dip, o1=0, d1=0.004, order=1, i0=0, eps=0.01, verb=False, ctx=None):
    """Relative geological time: the time axis painted along the slopes (reference pyseistr/rgt.py:rgt)."""
    n1 = np.asarray(dip).shape[0]
    trace = np.linspace(0, d1 * (n1 - 1), n1) + o1
    return pwpaintc(dip, trace, order=order, i0=i0, eps=eps, verb=verb, ctx=ctx)


def smoothc(din, rect=[1, 1, 1], diff=[0, 0, 0], box=[0, 0, 0], repeat=1, adj=1, ctx=None):
    """N-D triangle / box smoothing (reference pyseistr/smooth.py:115-183 -> dipcfun.smoothcf, dip_cfuns.c:2006-2123),
    same defaults as the reference (note adj=1).  adj=0 is ps_smooth2 (the operator inside dip3d's shaping CG, here the
    streaming kernel), adj=1 is ps_smooth; diff / box select single integration / box weights per axis."""
    din = np.asarray(din)
    n1, n2, n3 = _shape3(din)
    if int(repeat) < 1:
        raise ValueError("repeat must be >= 1")
    c = _ctx(ctx)
    d = _F(din)
    out = np.empty_like(d)
    _lib.check(c.lib.pst_smoothcf(c.handle, _p(d), n1, n2, n3, int(repeat), int(bool(adj)), int(rect[0]), int(rect[1]),
                                  int(rect[2]), int(bool(diff[0])), int(bool(diff[1])), int(bool(diff[2])),
                                  int(bool(box[0])), int(bool(box[1])), int(bool(box[2])), _p(out)))
    return np.squeeze(out.reshape(n1, n2, n3, order="F"))
