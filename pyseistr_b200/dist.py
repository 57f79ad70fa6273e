"""Multi-GPU host side: one process per GPU, the cube cut into n3-slabs (slowest axis).

The compute and all data-path communication (NCCL halos, carry planes, all-reduced CG scalars)
live in libpst_b200; this module only does the plumbing a launcher needs: which planes a rank
owns, getting rank 0's 128-byte NCCL id to every rank (torch.distributed, any backend), and slab
variants of the entry points (same arguments as the reference-facing ones, plus the global n3).
"""
import ctypes

import numpy as np

from . import _lib

_fp = ctypes.POINTER(ctypes.c_float)


def slab_bounds(n3, rank, world):
    """Global plane range [z0, z1) of `rank` — the same rule as pst_ctx_slab in the library."""
    return (n3 * rank) // world, (n3 * (rank + 1)) // world


def check_slabs(n3, world, r3=1, ns3=0):
    """Slabs must hold the axis-3 smoothing taps (2*r3 planes) and the xline spray radius."""
    need = max(2 * r3 if r3 > 1 else 1, ns3, 1)
    thin = min(slab_bounds(n3, r, world)[1] - slab_bounds(n3, r, world)[0] for r in range(world))
    if thin < need:
        raise ValueError(f"{n3} planes over {world} ranks leaves a slab of {thin} planes; need >= {need}")
    return thin


def broadcast_id(dist, make_id=None, src=0):
    """Create the communicator id on `src` and hand it to every rank through torch.distributed
    (works with gloo or nccl process groups).  `make_id` defaults to the library's NCCL id."""
    rank = dist.get_rank()
    payload = [None]
    if rank == src:
        payload[0] = (make_id or _lib.unique_id)()
    dist.broadcast_object_list(payload, src=src)
    ident = payload[0]
    if not isinstance(ident, (bytes, bytearray)) or len(ident) != 128:
        raise RuntimeError("communicator id must be 128 bytes")
    return bytes(ident)


def context_from_torch(dist, device):
    """One library context per rank of an initialised torch.distributed process group."""
    world, rank = dist.get_world_size(), dist.get_rank()
    if world == 1:
        return _lib.Context(device)
    return _lib.Context(device, rank, world, broadcast_id(dist))


def _F(a):
    return np.ascontiguousarray(np.float32(a).flatten(order="F"))


def _p(a):
    return a.ctypes.data_as(_fp)


def dip3dc_slab(ctx, slab, n3, niter=5, liter=10, order=2, rect=(5, 5, 5), verb=0, mask=None):
    """dip3dc on this rank's slab (n1, n2, z1-z0) of a cube with n3 planes in total (mask: the slab of the mask)."""
    n1, n2, nz = slab.shape
    d = _F(slab)
    m = _F(mask) if mask is not None else None
    out = np.empty(2 * d.size, np.float32)
    _lib.check(ctx.lib.pst_dip(ctx.handle, _p(d), _p(m) if m is not None else None, n1, n2, int(n3), int(niter), int(liter), int(order),
                               0.01, 1.0, 1e-6, int(rect[0]), int(rect[1]), int(rect[2]), int(verb), _p(out)))
    out = out.reshape(n1, n2, nz, 2, order="F")
    return out[:, :, :, 0], out[:, :, :, 1]


def somf3dc_slab(ctx, slab, dipi, dipx, n3, r1, r2, order, option=1):
    n1, n2, nz = slab.shape
    d, a, b = _F(slab), _F(dipi), _F(dipx)
    out = np.empty_like(d)
    _lib.check(ctx.lib.pst_somf3d(ctx.handle, _p(d), _p(a), _p(b), n1, n2, int(n3), int(r1), int(r2),
                                  2 * int(r1) * int(r2) + 1, int(option), int(order), 0.01, 0, _p(out)))
    return out.reshape(n1, n2, nz, order="F")


def somean3dc_slab(ctx, slab, dipi, dipx, n3, r1, r2, order):
    n1, n2, nz = slab.shape
    d, a, b = _F(slab), _F(dipi), _F(dipx)
    out = np.empty_like(d)
    _lib.check(ctx.lib.pst_somean3d(ctx.handle, _p(d), _p(a), _p(b), n1, n2, int(n3), int(r1), int(r2),
                                    int(order), 0.01, 0, _p(out)))
    return out.reshape(n1, n2, nz, order="F")


def soint3dc_slab(ctx, slab, mask, dipi, dipx, n3, order=1, niter=100, njs=(1, 1), hasmask=1, verb=0, var=0.0, seed=202223):
    """soint3dc on this rank's slab: the xline stencil reads one plane of the next rank, its adjoint one plane of
    the previous rank, the CG dots are all-reduced (drift = 0; var > 0: every rank draws the whole noise stream)."""
    n1, n2, nz = slab.shape
    d, a, b = _F(slab), _F(dipi), _F(dipx)
    m = _F(mask) if mask is not None else None
    out = np.empty_like(d)
    _lib.check(ctx.lib.pst_soint3d(ctx.handle, _p(d), _p(m) if m is not None else None, _p(a), _p(b), n1, n2, int(n3),
                                   int(order), int(njs[0]), int(njs[1]), int(niter), 0, int(seed), int(hasmask), float(var),
                                   int(verb), _p(out)))
    return out.reshape(n1, n2, nz, order="F")


def smoothc_slab(ctx, slab, n3, rect):
    """Plain N-D triangle smoothing (smoothcf with adj = 0, repeat = 1, no diff / box) of this rank's slab: the
    axis-3 running sums run across the ranks in the reference's order, so the result is the single-GPU one."""
    n1, n2, nz = slab.shape
    d = _F(slab)
    out = np.empty_like(d)
    _lib.check(ctx.lib.pst_smoothcf(ctx.handle, _p(d), n1, n2, int(n3), 1, 0, int(rect[0]), int(rect[1]), int(rect[2]),
                                    0, 0, 0, 0, 0, 0, _p(out)))
    return out.reshape(n1, n2, nz, order="F")


def sint3dc_slab(ctx, slab, mask, dipi, dipx, n3, niter=100, eps=0.01, ns1=1, ns2=1, order1=1, order2=1, verb=0):
    """sint3dc on this rank's slab: the inline plane-wave smoother is local, the xline smoother exchanges an ns2-plane
    halo of its input with both neighbours per application (forward and adjoint), the CG dots are all-reduced."""
    n1, n2, nz = slab.shape
    d, m, a, b = _F(slab), _F(mask), _F(dipi), _F(dipx)
    out = np.empty_like(d)
    _lib.check(ctx.lib.pst_sint3d(ctx.handle, _p(d), _p(a), _p(b), _p(m), n1, n2, int(n3), int(niter), int(ns1), int(ns2),
                                  int(order1), int(order2), int(verb), float(eps), _p(out)))
    return out.reshape(n1, n2, nz, order="F")


def somean2dc_slab(ctx, slab, dip, n3, ns, order, eps, adj=0):
    """somean2dc on a stack of (n1 x n2) panels cut into n3-slabs: the panels are independent."""
    n1, n2, nz = slab.shape
    d, a = _F(slab), _F(dip)
    out = np.empty_like(d)
    _lib.check(ctx.lib.pst_somean2d(ctx.handle, _p(d), _p(a), n1, n2, int(n3), int(ns), int(order), int(adj), float(eps), 0, _p(out)))
    return out.reshape(n1, n2, nz, order="F")
