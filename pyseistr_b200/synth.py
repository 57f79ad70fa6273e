"""Seeded synthetic inputs shaped like the reference demos.

``cube`` = sum of Ricker plane events w(t - t0 - px*i2 - py*i3) (one mildly curved),
normalised to max|d| = 1, plus Gaussian noise; ``erratic`` adds the spiky traces the
somf3d demo uses (reference demos/test_pyseistr_somf3d.py:20-31).  Pure NumPy; needs neither
the checker nor CUDA.
"""
import numpy as np


def ricker(tau, f):
    a = (np.pi * f * tau) ** 2
    return (1.0 - 2.0 * a) * np.exp(-a)


def cube(n1, n2, n3=1, seed=0, noise=0.05, nevents=4, dtype=np.float32):
    """float32 (n1,n2,n3) cube (or (n1,n2) panel when n3 == 1), Fortran-ordered."""
    rng = np.random.default_rng(seed)
    t = np.arange(n1, dtype=np.float64)[:, None, None]
    x = np.arange(n2, dtype=np.float64)[None, :, None]
    y = np.arange(n3, dtype=np.float64)[None, None, :]
    d = np.zeros((n1, n2, n3), dtype=np.float64)
    for k in range(nevents):
        f = rng.uniform(0.04, 0.1)
        px = rng.uniform(-0.5, 0.5)
        py = rng.uniform(-0.5, 0.5)
        t0 = rng.uniform(0.2, 0.8) * n1
        curv = 0.0 if k else rng.uniform(-0.2, 0.2) / max(n2, 2)
        amp = rng.uniform(0.5, 1.0) * rng.choice([-1.0, 1.0])
        tau = t - t0 - px * (x - n2 / 2) - py * (y - n3 / 2) - curv * (x - n2 / 2) ** 2
        d += amp * ricker(tau, f)
    d /= np.abs(d).max()
    if noise:
        d += noise * rng.standard_normal(d.shape)
    d = np.asfortranarray(d.astype(dtype))
    return d[:, :, 0] if n3 == 1 else d


def erratic(d, seed=202122, ntraces=6, amp=2.0):
    """Add uniform(-1,1)*amp spikes on a few random traces (erratic noise of the somf demos)."""
    rng = np.random.default_rng(seed)
    out = np.array(d, copy=True, order="F")
    n1 = out.shape[0]
    flat = out.reshape(n1, -1, order="F")
    idx = rng.choice(flat.shape[1], size=min(ntraces, flat.shape[1]), replace=False)
    flat[:, idx] += (amp * rng.uniform(-1.0, 1.0, size=(n1, idx.size))).astype(out.dtype)
    return flat.reshape(out.shape, order="F")


def smooth_dips(n1, n2, n3=1, seed=1, amp=0.6, dtype=np.float32):
    """A pair of smooth analytic slope fields |sigma| <= amp for dip-independent spray tests."""
    rng = np.random.default_rng(seed)
    t = np.linspace(0, 1, n1)[:, None, None]
    x = np.linspace(0, 1, n2)[None, :, None]
    y = np.linspace(0, 1, max(n3, 1))[None, None, :]
    ph = rng.uniform(0, 2 * np.pi, size=6)
    pi_ = amp * np.sin(2 * np.pi * (0.7 * t + 1.1 * x + 0.5 * y) + ph[0]) * np.cos(3 * x + ph[1])
    px_ = amp * np.cos(2 * np.pi * (0.4 * t - 0.8 * x + 0.9 * y) + ph[2]) * np.sin(2 * y + ph[3])
    pi_ = np.asfortranarray(np.broadcast_to(pi_, (n1, n2, max(n3, 1))).astype(dtype))
    px_ = np.asfortranarray(np.broadcast_to(px_, (n1, n2, max(n3, 1))).astype(dtype))
    if n3 == 1:
        return pi_[:, :, 0], px_[:, :, 0]
    return pi_, px_


def cube_big(n1, n2, n3, seed=0, noise=0.05, nevents=4, threads=None, out=None, z0=0, z1=None,
             normalise=True):
    """Same family of cubes as ``cube`` but generated plane-chunk by plane-chunk in float32 on a
    thread pool, for bench-sized volumes (1e9 voxels).  Writes into ``out`` (any float32
    Fortran-ordered array, e.g. a pinned buffer) when given.  ``z0:z1`` selects a slab of planes
    of the same global cube (multi-GPU ranks generate only their slab); returns (array, max|d|)
    when ``normalise`` is False so the caller can normalise with the global maximum."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    rng = np.random.default_rng(seed)
    ev = []
    for k in range(nevents):
        ev.append(dict(f=rng.uniform(0.04, 0.1), px=rng.uniform(-0.5, 0.5), py=rng.uniform(-0.5, 0.5),
                       t0=rng.uniform(0.2, 0.8) * n1,
                       curv=0.0 if k else rng.uniform(-0.2, 0.2) / max(n2, 2),
                       amp=rng.uniform(0.5, 1.0) * rng.choice([-1.0, 1.0])))
    z1 = n3 if z1 is None else z1
    if out is None:
        out = np.empty((n1, n2, z1 - z0), dtype=np.float32, order="F")
    t = np.arange(n1, dtype=np.float32)[:, None]
    x = (np.arange(n2, dtype=np.float32) - n2 / 2)[None, :]
    threads = threads or min(32, os.cpu_count() or 1)

    def plane(i3):
        acc = np.zeros((n1, n2), np.float32)
        for e in ev:
            tau = t - np.float32(e["t0"]) - np.float32(e["px"]) * x - np.float32(e["py"] * (i3 - n3 / 2)) \
                - np.float32(e["curv"]) * x * x
            a = (np.float32(np.pi * e["f"]) * tau) ** 2
            acc += np.float32(e["amp"]) * (1.0 - 2.0 * a) * np.exp(-a)
        if noise:
            acc += np.float32(noise) * np.random.default_rng([seed, i3]).standard_normal((n1, n2), dtype=np.float32)
        out[:, :, i3 - z0] = acc
        return float(np.abs(acc).max())

    with ThreadPoolExecutor(threads) as ex:
        mx = max(ex.map(plane, range(z0, z1)))
    if not normalise:
        return out, mx
    scale_by(out, mx, threads)
    return out


def scale_by(out, mx, threads=None):
    """out /= mx, plane by plane on a thread pool (second half of ``cube_big``)."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    if mx <= 0:
        return out
    scale = np.float32(1.0 / mx)

    def norm(i3):
        out[:, :, i3] *= scale

    with ThreadPoolExecutor(threads or min(32, os.cpu_count() or 1)) as ex:
        list(ex.map(norm, range(out.shape[2])))
    return out
