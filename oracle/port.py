"""TEST INFRASTRUCTURE ONLY — ctypes binding of oracle/libpst_oracle.so (pst_oracle.c, our
plain-C restatement of the reference hot path).  Never imported by pyseistr_b200/.
Signatures mirror the reference's ``*c`` wrappers so parity tests read like its demos."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_fp = ctypes.POINTER(ctypes.c_float)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "port"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libpst_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
    return _LIB


def _F(a):
    return np.ascontiguousarray(np.float32(a).flatten(order="F"))


def _p(a):
    return a.ctypes.data_as(_fp)


def _shape3(d):
    return (d.shape[0], d.shape[1], 1) if d.ndim == 2 else d.shape


def set_dot_mode(mode):
    """0 = reference (sequential double dots); 1 = blocked association (sensitivity probe)."""
    lib().pso_set_dot_mode(int(mode))


def passfilter(nw, sigma):
    a = np.zeros(2 * nw + 1, np.float32)
    lib().pso_passfilter(ctypes.c_int(nw), ctypes.c_float(sigma), _p(a))
    return a


def aderfilter(nw, sigma):
    a = np.zeros(2 * nw + 1, np.float32)
    lib().pso_aderfilter(ctypes.c_int(nw), ctypes.c_float(sigma), _p(a))
    return a


def allpass(u, sigma, nw, xline, der):
    n1, n2, n3 = _shape3(u)
    uu, ss = _F(u), _F(sigma)
    y = np.zeros_like(uu)
    lib().pso_allpass(_p(uu), _p(ss), n1, n2, n3, nw, int(xline), int(der), _p(y))
    return y.reshape(n1, n2, n3, order="F")


def smooth3(x, rect, repeat=1, adj=0):
    n1, n2, n3 = _shape3(x)
    xx = _F(x).copy()
    fn = lib().pso_smooth3_fwd if adj else lib().pso_smooth3_rep
    fn(_p(xx), n1, n2, n3, int(rect[0]), int(rect[1]), int(rect[2]), int(repeat))
    return xx.reshape(n1, n2, n3, order="F")


def smoothc(x, rect, diff=(0, 0, 0), box=(0, 0, 0), repeat=1, adj=1):
    """smoothcf with every option (reference pyseistr/smooth.py:115-183; note the default adj=1)."""
    n1, n2, n3 = _shape3(x)
    xx = _F(x).copy()
    I3 = ctypes.c_int * 3
    lib().pso_smoothcf(_p(xx), n1, n2, n3, int(repeat), int(adj), I3(*[int(v) for v in rect]), I3(*[int(v) for v in diff]),
                       I3(*[int(v) for v in box]))
    return xx.reshape(n1, n2, n3, order="F")


def divne(num, den, rect, liter, eps=1.0):
    n1, n2, n3 = _shape3(num)
    a, b = _F(num).copy(), _F(den).copy()
    rat = np.zeros_like(a)
    it = lib().pso_divne(_p(a), _p(b), _p(rat), n1, n2, n3, int(rect[0]), int(rect[1]), int(rect[2]),
                         int(liter), ctypes.c_float(eps))
    return rat.reshape(n1, n2, n3, order="F"), it


def dip3dc(din, niter=5, liter=10, order=2, eps_dv=0.01, eps_cg=1, tol_cg=0.000001,
           rect=(5, 5, 5), verb=0, mask=None):
    n1, n2, n3 = din.shape
    d = _F(din)
    m = _F(mask) if mask is not None else None
    out = np.zeros(2 * d.size if n3 != 1 else d.size, np.float32)
    lib().pso_dip(_p(d), _p(m) if m is not None else None, n1, n2, n3, niter, liter, order,
                  int(rect[0]), int(rect[1]), int(rect[2]), _p(out))
    if n3 == 1:
        return out.reshape(n1, n2, 1, order="F"), None
    out = out.reshape(n1, n2, n3, 2, order="F")
    return out[:, :, :, 0], out[:, :, :, 1]


def dip_counts():
    """{CG iterations, line-search evaluations, GN iterations} executed by the last dip3dc / dip2dc call."""
    c = (ctypes.c_longlong * 3)()
    lib().pso_get_counts(c)
    return {"cg_iterations": int(c[0]), "linesearch_evals": int(c[1]), "gn_iterations": int(c[2])}


def dip2dc(din, niter=5, liter=20, order=2, eps_dv=0.01, eps_cg=1, tol_cg=0.000001,
           rect=(10, 10, 1), verb=0, mask=None):
    n1, n2 = din.shape
    d = _F(din)
    m = _F(mask) if mask is not None else None
    out = np.zeros(d.size, np.float32)
    lib().pso_dip(_p(d), _p(m) if m is not None else None, n1, n2, 1, niter, liter, order,
                  int(rect[0]), int(rect[1]), int(rect[2]), _p(out))
    return out.reshape(n1, n2, order="F")


def predict(trace1, sig1, nw, forw1, eps=1e-4, trace2=None, sig2=None, forw2=0):
    n1 = trace1.size
    t1, s1 = _F(trace1), _F(sig1)
    two = trace2 is not None
    t2 = _F(trace2) if two else t1
    s2 = _F(sig2) if two else s1
    out = np.zeros(n1, np.float32)
    lib().pso_predict(n1, nw, ctypes.c_float(eps), int(two), int(forw1), int(forw2),
                      _p(t1), _p(t2), _p(s1), _p(s2), _p(out))
    return out


def somean3dc(dn, dipi, dipx, r1, r2, eps, order, verb=0):
    n1, n2, n3 = _shape3(dn)
    d, a, b = _F(dn), _F(dipi), _F(dipx)
    out = np.zeros_like(d)
    lib().pso_somean3d(_p(d), _p(a), _p(b), n1, n2, n3, r1, r2, order, _p(out))
    return out.reshape(n1, n2, n3, order="F")


def somf3dc(dn, dipi, dipx, r1, r2, eps, order, option=1, verb=0):
    n1, n2, n3 = _shape3(dn)
    d, a, b = _F(dn), _F(dipi), _F(dipx)
    out = np.zeros_like(d)
    rc = lib().pso_somf3d(_p(d), _p(a), _p(b), n1, n2, n3, r1, r2, 2 * r1 * r2 + 1, option, order, _p(out))
    if rc:
        raise ValueError("oracle: unsupported somf3d option")
    return out.reshape(n1, n2, n3, order="F")


def somean2dc(dn, dip, ns, order, eps, adj=0, verb=0):
    n1, n2, n3 = _shape3(dn)
    d, a = _F(dn), _F(dip)
    out = np.zeros_like(d)
    fn = lib().pso_somean2d_adj if adj else lib().pso_somean2d
    fn(_p(d), _p(a), n1, n2, n3, ns, order, ctypes.c_float(eps), _p(out))
    return np.squeeze(out.reshape(n1, n2, n3, order="F"))


def somf2dc(dn, dip, ns, order, eps, option=1, verb=0):
    n1, n2, n3 = _shape3(dn)
    d, a = _F(dn), _F(dip)
    out = np.zeros_like(d)
    rc = lib().pso_somf2d(_p(d), _p(a), n1, n2, n3, ns, 2 * ns + 1, option, order, ctypes.c_float(eps), _p(out))
    if rc:
        raise ValueError("oracle: unsupported somf2d option")
    return np.squeeze(out.reshape(n1, n2, n3, order="F"))


def soint3dc(din, mask, dipi, dipx, order=1, niter=100, njs=(1, 1), drift=0, seed=202223, hasmask=1, var=0, verb=0):
    n1, n2, n3 = _shape3(din)
    d, a, b = _F(din), _F(dipi), _F(dipx)
    m = _F(mask) if hasmask else None
    out = np.zeros_like(d)
    lib().pso_soint3d_full(_p(d), _p(m) if m is not None else None, _p(a), _p(b), n1, n2, n3, int(order),
                           int(njs[0]), int(njs[1]), int(niter), int(seed), ctypes.c_float(var), _p(out))
    return out.reshape(n1, n2, n3, order="F")


def sint3dc(din, mask, dipi, dipx, niter=100, eps=0.01, ns1=1, ns2=1, order1=1, order2=1, verb=0):
    n1, n2, n3 = _shape3(din)
    d, a, b, m = _F(din), _F(dipi), _F(dipx), _F(mask)
    out = np.zeros_like(d)
    lib().pso_sint3d(_p(d), _p(a), _p(b), _p(m), n1, n2, n3, int(niter), int(ns1), int(ns2), int(order1), int(order2),
                     ctypes.c_float(eps), _p(out))
    return out.reshape(n1, n2, n3, order="F")


def soint2dc(din, mask, dip, order=1, niter=100, njs=(1, 1), hasmask=1):
    """csoint2d default path = csoint3d on an (n1, n2, 1) volume (allpass21_lop is the inline half of allpass3_lop)."""
    n1, n2 = din.shape
    r3 = lambda a: np.float32(a).reshape(n1, n2, 1)
    return soint3dc(r3(din), r3(mask), r3(dip), r3(dip), order=order, niter=niter, njs=njs, hasmask=hasmask).reshape(n1, n2)


def sint2dc(din, mask, dip, niter=100, eps=0.01, ns=1, order=1, verb=0):
    n1, n2 = din.shape
    d, a, m = _F(din), _F(dip), _F(mask)
    out = np.zeros_like(d)
    lib().pso_sint2d(_p(d), _p(a), _p(m), n1, n2, int(niter), int(ns), int(order), ctypes.c_float(eps), _p(out))
    return out.reshape(n1, n2, order="F")


def predict_adj(trace, sig, nw, forw, eps=1e-4):
    t = _F(trace).copy()
    s = _F(sig)
    lib().pso_predict_adj(t.size, int(nw), ctypes.c_float(eps), int(forw), _p(t), _p(s))
    return t


def pwpaintc(dip, trace, order=1, i0=0, eps=0.01, verb=False):
    """plane-wave painting (reference pyseistr/rgt.py:pwpaintc -> cpaint2d, paint_cfuns.c:1861)"""
    n1, n2 = dip.shape
    d, t = _F(dip), _F(trace)
    out = np.zeros_like(d)
    rc = lib().pso_paint2d(_p(d), _p(t), n1, n2, int(order), int(i0), ctypes.c_float(eps), _p(out))
    if rc:
        raise ValueError("oracle: bad painting arguments")
    return out.reshape(n1, n2, order="F")
