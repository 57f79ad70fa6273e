"""TEST INFRASTRUCTURE ONLY — never imported by the product path (pyseistr_b200/).

Loader for ``oracle/_ref``: the UNMODIFIED reference C extension modules
(``dipcfun``, ``sof3dcfun``, ``sofcfun``, ``soint3dcfun``) compiled by
``oracle/Makefile`` from the sources where they lie under /root/reference.
The thin ``*c`` wrappers of the reference (pure reshaping glue) are restated here
because ``import pyseistr`` needs matplotlib (reference pyseistr/plot.py:1).

Each wrapper cites the reference wrapper it mirrors.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py`` (cpu_baseline / ``--impl reference``)
may import this module.
"""
import contextlib
import importlib
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")
_mods = {}


def available():
    """True when the compiled reference modules are present (built here or shipped)."""
    if not os.path.isdir(_REF_DIR):
        return False
    names = os.listdir(_REF_DIR)
    return all(any(n.startswith(m + ".") for n in names)
               for m in ("dipcfun", "sof3dcfun", "sofcfun", "soint3dcfun"))


def module(name):
    """Import one reference extension module from oracle/_ref by its top-level name."""
    if name not in _mods:
        if _REF_DIR not in sys.path:
            sys.path.insert(0, _REF_DIR)
        _mods[name] = importlib.import_module(name)
    return _mods[name]


@contextlib.contextmanager
def quiet():
    """The reference prints unconditionally from C: mute fd 1 around a call."""
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    try:
        os.dup2(devnull, 1)
        yield
    finally:
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)


def _F(a):
    return np.float32(a).flatten(order="F")


def dip3dc(din, niter=5, liter=10, order=2, eps_dv=0.01, eps_cg=1, tol_cg=0.000001,
           rect=(5, 5, 5), verb=0, mask=None):
    """reference pyseistr/dip3d.py:59-116 (dipc positional order, dip_cfuns.c:1708)."""
    n1, n2, n3 = din.shape
    if mask is None:
        d, hasmask = _F(din), 0
    else:
        d, hasmask = np.float32(np.concatenate([_F(din), _F(mask)])), 1
    with quiet():
        dip = module("dipcfun").dipc(d, n1, n2, n3, niter, liter, order, eps_dv, eps_cg, tol_cg,
                                     rect[0], rect[1], rect[2], hasmask, verb)
    dip = dip.reshape(n1, n2, n3, 2, order="F")
    return dip[:, :, :, 0], dip[:, :, :, 1]


def dip2dc(din, niter=5, liter=20, order=2, eps_dv=0.01, eps_cg=1, tol_cg=0.000001,
           rect=(10, 10, 1), verb=0, mask=None):
    """reference pyseistr/dip2d.py:115-221 (same dipc entry, n3=1)."""
    n1, n2 = din.shape
    if mask is None:
        d, hasmask = _F(din), 0
    else:
        d, hasmask = np.float32(np.concatenate([_F(din), _F(mask)])), 1
    with quiet():
        dip = module("dipcfun").dipc(d, n1, n2, 1, niter, liter, order, eps_dv, eps_cg, tol_cg,
                                     rect[0], rect[1], rect[2], hasmask, verb)
    return dip.reshape(n1, n2, order="F")


def _shape3(dn):
    if dn.ndim == 2:
        return dn.shape[0], dn.shape[1], 1
    return dn.shape


def somf3dc(dn, dipi, dipx, r1, r2, eps, order, option=1, verb=0):
    """reference pyseistr/somf3d.py:54-97 (csomf3d, sof3d_cfuns.c:1573)."""
    n1, n2, n3 = _shape3(dn)
    with quiet():
        ds = module("sof3dcfun").csomf3d(_F(dn), _F(dipi), _F(dipx), n1, n2, n3, r1, r2,
                                         2 * r1 * r2 + 1, option, order, eps, verb)
    return ds.reshape([n1, n2, n3], order="F")


def somean3dc(dn, dipi, dipx, r1, r2, eps, order, verb=0):
    """reference pyseistr/somean3d.py:36-73 (csomean3d, sof3d_cfuns.c:1372)."""
    n1, n2, n3 = _shape3(dn)
    with quiet():
        ds = module("sof3dcfun").csomean3d(_F(dn), _F(dipi), _F(dipx), n1, n2, n3, r1, r2,
                                           order, eps, verb)
    return ds.reshape([n1, n2, n3], order="F")


def somf2dc(dn, dip, ns, order, eps, option=1, verb=0):
    """reference pyseistr/somf2d.py:60-105 (csomf2d, sof_cfuns.c:1550)."""
    n1, n2, n3 = _shape3(dn)
    with quiet():
        ds = module("sofcfun").csomf2d(_F(dn), _F(dip), n1, n2, n3, ns, 2 * ns + 1, option,
                                       order, eps, verb)
    return np.squeeze(ds.reshape(n1, n2, n3, order="F"))


def somean2dc(dn, dip, ns, order, eps, adj=0, verb=0):
    """reference pyseistr/somean2d.py:36-74 (csomean2d, sof_cfuns.c:1448)."""
    n1, n2, n3 = _shape3(dn)
    with quiet():
        ds = module("sofcfun").csomean2d(_F(dn), _F(dip), n1, n2, n3, ns, order, adj, eps, verb)
    return np.squeeze(ds.reshape(n1, n2, n3, order="F"))


def soint3dc(din, mask, dipi, dipx, order=1, niter=100, njs=(1, 1), drift=0, seed=202223,
             hasmask=1, var=0, verb=0):
    """reference pyseistr/soint3d.py:65-108 (csoint3d, soint3d_cfuns.c:2423)."""
    n1, n2, n3 = _shape3(din)
    with quiet():
        d = module("soint3dcfun").csoint3d(_F(din), _F(mask), _F(dipi), _F(dipx), n1, n2, n3,
                                           order, njs[0], njs[1], niter, drift, seed, hasmask,
                                           var, verb)
    return d.reshape(n1, n2, n3, order="F")


def sint3dc(din, mask, dipi, dipx, niter=100, eps=0.01, ns1=1, ns2=1, order1=1, order2=1, verb=0):
    """reference pyseistr/sint.py:97-131 (csint3d, soint3d_cfuns.c:2528)."""
    n1, n2, n3 = _shape3(din)
    with quiet():
        d = module("soint3dcfun").csint3d(_F(din), _F(dipi), _F(dipx), _F(mask), n1, n2, n3,
                                          niter, ns1, ns2, order1, order2, verb, eps)
    return d.reshape(n1, n2, n3, order="F")


def soint2dc(din, mask, dip, order=1, niter=100, njs=(1, 1), drift=0, hasmask=1, twoplane=0, prec=0, verb=0):
    """reference pyseistr/soint2d.py:92-141 (csoint2d, soint2d_cfuns.c:2260), one slope field."""
    n1, n2 = din.shape
    with quiet():
        d = module("soint2dcfun").csoint2d(_F(din), _F(mask), _F(dip), _F(dip), n1, n2, order, njs[0], njs[1], niter,
                                           drift, hasmask, twoplane, prec, verb)
    return d.reshape(n1, n2, order="F")


def sint2dc(din, mask, dip, niter=100, eps=0.01, ns=1, order=1, verb=0):
    """reference pyseistr/sint.py:61-94 (csint2d, soint2d_cfuns.c:2421); needs the optional soint2dcfun module."""
    n1, n2 = din.shape
    with quiet():
        d = module("soint2dcfun").csint2d(_F(din), _F(dip), _F(mask), n1, n2, niter, ns, order, verb, eps)
    return d.reshape(n1, n2, order="F")


def smoothc(x, rect, adj=0, repeat=1, diff=(0, 0, 0), box=(0, 0, 0)):
    """reference pyseistr/smooth.py:115-183 via dipcfun.smoothcf (dip_cfuns.c:2006-2123);
    adj=0 selects ps_smooth2, the kernel dip3d uses (the reference wrapper's own default is adj=1)."""
    n1, n2, n3 = _shape3(x)
    with quiet():
        y = module("dipcfun").smoothcf(_F(x), n1, n2, n3, int(repeat), int(adj), rect[0], rect[1], rect[2],
                                       int(diff[0]), int(diff[1]), int(diff[2]), int(box[0]), int(box[1]), int(box[2]))
    return np.asarray(y, dtype=np.float32).reshape(n1, n2, n3, order="F")


def pwpaintc(dip, trace, order=1, i0=0, eps=0.01, verb=False):
    """reference pyseistr/rgt.py:pwpaintc (the reshaping glue around paint2dcfun.cpaint2d)"""
    n1, n2 = dip.shape
    with quiet():
        out = module("paint2dcfun").cpaint2d(_F(dip), _F(trace), n1, n2, int(order), int(i0), float(eps), int(verb))
    return np.asarray(out, np.float32).reshape(n1, n2, order="F")
