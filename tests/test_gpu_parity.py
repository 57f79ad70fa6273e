"""GPU suite (-m gpu): the CUDA path, called through the C-ABI, against (1) the golden
fixtures from the compiled reference, (2) the oracle port on fresh seeded inputs, and (3)
size-independent properties at larger sizes.

Tolerances.  Every float vector operation follows the reference's order without FMA, so the
building blocks (stencil, smoothing, spray, median) are required to be BIT-EXACT.  dip3d /
divne contain global sums that the reference accumulates sequentially (double for the CG
scalars, float for the line-search energies) and the GPU accumulates as double trees: their
results must agree to relative L2 <= 1e-5 (north-star tolerance); in practice they are
bit-identical or differ by ~1e-7.
"""
import ctypes

import numpy as np
import pytest

from conftest import golden, golden_names, rel_l2
from pyseistr_b200 import synth

pytestmark = pytest.mark.gpu

TOL = 1e-5
_vp = ctypes.c_void_p


def _F(a):
    return np.ascontiguousarray(np.float32(a).flatten(order="F"))


class Dev:
    """tiny RAII helper over pst_dev_alloc / pst_h2d / pst_d2h"""

    def __init__(self, ctx, arr=None, n=None):
        self.ctx = ctx
        self.n = arr.size if arr is not None else n
        self.p = ctx.alloc(4 * self.n)
        if arr is not None:
            ctx.h2d(self.p, _F(arr))

    def get(self, shape):
        out = np.empty(self.n, np.float32)
        self.ctx.d2h(out, self.p)
        return out.reshape(shape, order="F")

    def __del__(self):
        try:
            self.ctx.free(self.p)
        except Exception:
            pass


# ------------------------------------------------------------------ building blocks: bit-exact
@pytest.mark.parametrize("order", [1, 2])
@pytest.mark.parametrize("xline", [0, 1])
@pytest.mark.parametrize("der", [0, 1])
def test_allpass_bit_exact(ctx, port, order, xline, der):
    from pyseistr_b200 import _lib
    n1, n2, n3 = 37, 11, 6
    u = synth.cube(n1, n2, n3, seed=31)
    sg, _ = synth.smooth_dips(n1, n2, n3, seed=32, amp=1.5)
    du, ds, dy = Dev(ctx, u), Dev(ctx, sg), Dev(ctx, n=u.size)
    _lib.check(ctx.lib.pst_allpass_dev(ctx.handle, du.p, ds.p, n1, n2, n3, order, xline, der, dy.p))
    got = dy.get((n1, n2, n3))
    want = port.allpass(u, sg, order, xline, der)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("shape,rect", [((40, 12, 9), (5, 5, 5)), ((33, 7, 5), (3, 4, 2)),
                                        ((64, 20, 1), (10, 10, 1)), ((30, 12, 6), (2, 15, 9)),
                                        ((300, 9, 4), (40, 3, 1)), ((17, 5, 3), (1, 1, 1)),
                                        # streaming kernel (n1 % 4 == 0, radius <= 16): many tiles per CTA,
                                        # partial 32-line tiles, line ends inside / across pipeline stages
                                        ((256, 200, 150), (5, 5, 5)), ((100, 70, 37), (16, 2, 7)),
                                        ((1000, 40, 33), (8, 3, 10)), ((36, 1030, 9), (4, 6, 2)),
                                        ((12, 6, 1030), (3, 2, 13)), ((64, 64, 64), (9, 11, 12))])
def test_smooth3_bit_exact(ctx, port, shape, rect):
    import pyseistr_b200 as ps
    x = synth.cube(*shape, seed=33)
    got = ps.smoothc(x, rect=list(rect), adj=0, ctx=ctx)
    want = np.squeeze(port.smooth3(x, rect).reshape(x.shape, order="F"))
    assert np.array_equal(got, want)


@pytest.mark.parametrize("shape,r3,nranks", [((20, 13, 64), 5, 2), ((20, 13, 96), 5, 3), ((8, 40, 1024), 5, 8), ((37, 3, 224), 2, 7),
                                             ((16, 8, 256), 8, 2), ((12, 6, 128), 3, 4), ((12, 6, 160), 6, 5), ((9, 7, 384), 4, 3),
                                             # slabs of several chunks: 2 x 128 (4 ranks at n3 = 1024), 4 x 128 (2 ranks), 3 x 32, 2 x 32
                                             ((8, 24, 1024), 5, 4), ((8, 9, 1024), 5, 2), ((10, 7, 288), 3, 3), ((11, 6, 128), 6, 2),
                                             ((5, 13, 768), 8, 2)])
def test_multi_gpu_axis3_kernels_on_one_gpu(ctx, port, shape, r3, nranks):
    """The register kernels of the distributed axis-3 pass (first / interior / last rank variants, both tile orders),
    run rank after rank on one GPU over the n3-slabs of the volume: bit-identical to ps_smooth2 along axis 3."""
    n1, n2, n3 = shape
    x = synth.cube(n1, n2, n3, seed=41)
    d = Dev(ctx, x)
    from pyseistr_b200 import _lib
    _lib.check(ctx.lib.pst_selftest_axis3_slabs(ctx.handle, d.p, n1, n2, n3, r3, nranks))
    got = d.get(shape)
    want = port.smooth3(x, (1, 1, r3)).reshape(x.shape, order="F")
    assert np.array_equal(got, want)


@pytest.mark.parametrize("shape,rect,repeat", [((64, 20, 12), (5, 3, 4), 2), ((30, 12, 6), (2, 15, 9), 3), ((40, 9, 1), (4, 3, 1), 4)])
def test_smooth_repeat_bit_exact(ctx, port, shape, rect, repeat):
    """smoothc(repeat=k): every line of an axis is smoothed k times in a row (smoothcf dip_cfuns.c:2084-2098)."""
    import pyseistr_b200 as ps
    x = synth.cube(*shape, seed=34)
    got = ps.smoothc(x, rect=list(rect), repeat=repeat, adj=0, ctx=ctx)
    assert np.array_equal(got, np.squeeze(port.smooth3(x, rect, repeat).reshape(x.shape, order="F")))


@pytest.mark.parametrize("shape,rect,repeat", [((64, 20, 12), (5, 3, 4), 1), ((30, 12, 6), (2, 15, 9), 2), ((100, 70, 37), (7, 1, 3), 1)])
def test_smooth_adj1_bit_exact(ctx, port, shape, rect, repeat):
    """smoothc(adj=1) = ps_smooth: fold, backward + forward running sums, triple in double (dip_cfuns.c:591-603)."""
    import pyseistr_b200 as ps
    x = synth.cube(*shape, seed=35)
    got = ps.smoothc(x, rect=list(rect), repeat=repeat, ctx=ctx)                     # the reference's default IS adj=1
    assert np.array_equal(got, port.smooth3(x, rect, repeat, adj=1).reshape(x.shape, order="F"))


@pytest.mark.parametrize("adj", [0, 1])
@pytest.mark.parametrize("diff,box", [((1, 0, 0), (0, 0, 0)), ((0, 0, 0), (1, 1, 1)), ((0, 1, 1), (1, 0, 1))])
def test_smoothc_options_bit_exact(ctx, port, adj, diff, box):
    """smoothcf with derivative / box options on both operators (dip_cfuns.c:2006-2123)."""
    import pyseistr_b200 as ps
    x = synth.cube(48, 20, 9, seed=36)
    for rect, rep in (([5, 3, 4], 1), ([2, 25, 3], 2)):
        got = ps.smoothc(x, rect=rect, diff=list(diff), box=list(box), repeat=rep, adj=adj, ctx=ctx)
        assert np.array_equal(got, port.smoothc(x, rect, diff, box, rep, adj)), (rect, rep)


@pytest.mark.parametrize("name", golden_names("smooth_"))
def test_smooth_golden(ctx, name):
    import pyseistr_b200 as ps
    g = golden(name)
    assert np.array_equal(ps.smoothc(g["x"], rect=[int(v) for v in g["rect"]], adj=0, ctx=ctx), g["out"])


def test_divne_matches_oracle(ctx, port):
    from pyseistr_b200 import _lib
    n1, n2, n3 = 36, 10, 7
    num = synth.cube(n1, n2, n3, seed=41)
    den = synth.cube(n1, n2, n3, seed=42) + 0.5
    dn, dd, dr = Dev(ctx, num), Dev(ctx, den), Dev(ctx, n=num.size)
    its = ctypes.c_int(0)
    _lib.check(ctx.lib.pst_divne_dev(ctx.handle, dn.p, dd.p, dr.p, n1, n2, n3, 4, 3, 3, 12, ctypes.byref(its)))
    got = dr.get((n1, n2, n3))
    want, it_ref = port.divne(num, den, (4, 3, 3), 12)
    assert its.value == it_ref
    assert rel_l2(got, want) <= TOL


# ------------------------------------------------------------------ dip estimation
@pytest.mark.parametrize("name", golden_names("dip3d_"))
def test_dip3d_golden(ctx, name):
    import pyseistr_b200 as ps
    g = golden(name)
    di, dx = ps.dip3dc(g["din"], int(g["niter"]), int(g["liter"]), int(g["order"]),
                       rect=[int(v) for v in g["rect"]], verb=0, mask=g.get("mask"), ctx=ctx)
    assert di.shape == g["dipi"].shape and di.dtype == np.float32
    assert rel_l2(di, g["dipi"]) <= TOL, rel_l2(di, g["dipi"])
    assert rel_l2(dx, g["dipx"]) <= TOL, rel_l2(dx, g["dipx"])


def test_dip2d_golden(ctx):
    import pyseistr_b200 as ps
    g = golden("dip2d")
    p = ps.dip2dc(g["din"], int(g["niter"]), int(g["liter"]), int(g["order"]), rect=[int(v) for v in g["rect"]],
                  verb=0, ctx=ctx)
    assert p.shape == g["dip"].shape
    assert rel_l2(p, g["dip"]) <= TOL, rel_l2(p, g["dip"])


@pytest.mark.parametrize("shape,kw", [((100, 50, 10), dict()),
                                      ((60, 17, 12), dict(niter=3, liter=7, order=1, rect=[3, 6, 4])),
                                      ((50, 21, 3), dict(niter=2, liter=5, order=2, rect=[4, 4, 1]))])
def test_dip3d_vs_oracle(ctx, port, shape, kw):
    """C1-shaped cube of the reference demo (100x50x10, defaults) and two ragged shapes."""
    import pyseistr_b200 as ps
    d = synth.cube(*shape, seed=51)
    di, dx = ps.dip3dc(d, verb=0, ctx=ctx, **kw)
    oi, ox = port.dip3dc(d, **kw)
    assert rel_l2(di, oi) <= TOL, rel_l2(di, oi)
    assert rel_l2(dx, ox) <= TOL, rel_l2(dx, ox)
    st = ctx.stats()
    assert st["kernel_launches"] > 0 and st["cg_iterations"] > 0


def test_dip3d_ignores_eps_and_tol_like_reference(ctx):
    """eps_dv, eps_cg, tol_cg do not change the reference's C result."""
    import pyseistr_b200 as ps
    d = synth.cube(40, 12, 6, seed=52)
    a = ps.dip3dc(d, 2, 5, 2, 0.01, 1, 1e-6, [3, 3, 3], 0, ctx=ctx)
    b = ps.dip3dc(d, 2, 5, 2, 5.0, 7, 0.5, [3, 3, 3], 0, ctx=ctx)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_dip3d_constant_and_zero_input(ctx):
    import pyseistr_b200 as ps
    z = np.zeros((30, 8, 5), np.float32)
    di, dx = ps.dip3dc(z, verb=0, ctx=ctx)
    assert not di.any() and not dx.any()


# ------------------------------------------------------------------ spray: bit-exact
@pytest.mark.parametrize("name", golden_names("somean3d_") + golden_names("somf3d_"))
def test_spray3d_golden_bit_exact(ctx, name):
    import pyseistr_b200 as ps
    g = golden(name)
    fn = ps.somean3dc if name.startswith("somean") else ps.somf3dc
    out = fn(g["dn"], g["dipi"], g["dipx"], int(g["r1"]), int(g["r2"]), 0.01, int(g["order"]), verb=0, ctx=ctx)
    assert out.shape == g["out"].shape
    assert np.array_equal(out, g["out"]), rel_l2(out, g["out"])


@pytest.mark.parametrize("name", golden_names("somean2d_") + golden_names("somf2d_"))
def test_spray2d_golden_bit_exact(ctx, name):
    import pyseistr_b200 as ps
    g = golden(name)
    fn = ps.somean2dc if name.startswith("somean") else ps.somf2dc
    out = fn(g["dn"], g["dip"], int(g["ns"]), int(g["order"]), float(g["eps"]), verb=0, ctx=ctx)
    assert np.array_equal(out, g["out"]), rel_l2(out, g["out"])


@pytest.mark.parametrize("shape,r1,r2,order", [((100, 50, 10), 2, 2, 2), ((31, 5, 4), 2, 2, 1),
                                               ((40, 3, 9), 1, 2, 2), ((25, 140, 2), 2, 1, 2)])
def test_spray3d_vs_oracle_bit_exact(ctx, port, shape, r1, r2, order):
    """Edge cases: axes shorter than the spray diameter, one-plane halos, wide n2 (several
    thread blocks per plane)."""
    import pyseistr_b200 as ps
    d = synth.erratic(synth.cube(*shape, seed=61))
    pi, px = synth.smooth_dips(*shape, seed=62, amp=0.8)
    assert np.array_equal(ps.somf3dc(d, pi, px, r1, r2, 0.01, order, verb=0, ctx=ctx),
                          port.somf3dc(d, pi, px, r1, r2, 0.01, order))
    assert np.array_equal(ps.somean3dc(d, pi, px, r1, r2, 0.01, order, ctx=ctx),
                          port.somean3dc(d, pi, px, r1, r2, 0.01, order))


def test_spray_on_2d_input_like_reference(ctx, port):
    """somf3dc/somean3dc accept a 2-D panel (n3=1), reference somf3d.py:85-88."""
    import pyseistr_b200 as ps
    d = synth.cube(48, 20, 1, seed=63)
    p, _ = synth.smooth_dips(48, 20, 1, seed=64)
    z = np.zeros_like(p)
    got = ps.somf3dc(d, p, z, 2, 2, 0.01, 2, verb=0, ctx=ctx)
    assert got.shape == (48, 20, 1)
    assert np.array_equal(got, port.somf3dc(d, p, z, 2, 2, 0.01, 2))


@pytest.mark.parametrize("name", golden_names("svmf3d_") + golden_names("svmf2d_"))
def test_svmf_golden_bit_exact(ctx, name):
    """somf3dc / somf2dc with option=2 (space-varying median filter) against what the compiled reference returned, in
    the defined-behaviour variant of its one-row over-read."""
    import pyseistr_b200 as ps
    g = golden(name)
    if name.startswith("svmf3d"):
        out = ps.somf3dc(g["dn"], g["dipi"], g["dipx"], int(g["r1"]), int(g["r2"]), 0.01, int(g["order"]), option=2, verb=0, ctx=ctx)
    else:
        out = ps.somf2dc(g["dn"], g["dip"], int(g["ns"]), int(g["order"]), float(g["eps"]), option=2, verb=0, ctx=ctx)
    assert np.array_equal(out, g["out"]), rel_l2(out, g["out"])


def test_svmf_vs_oracle_and_short_windows_refused(ctx, port):
    import pyseistr_b200 as ps
    d = synth.erratic(synth.cube(60, 21, 7, seed=66), ntraces=8)
    di, dx = synth.smooth_dips(60, 21, 7, seed=66)
    for r1, r2, order in ((2, 2, 2), (1, 2, 1), (2, 3, 1)):
        got = ps.somf3dc(d, di, dx, r1, r2, 0.01, order, option=2, verb=0, ctx=ctx)
        assert np.array_equal(got, port.somf3dc(d, di, dx, r1, r2, 0.01, order, option=2)), (r1, r2, order)
    with pytest.raises(ps.PstError) as e:             # nmf = 3: the reference's shortest window would have length -1
        ps.somf3dc(d, di, dx, 1, 1, 0.01, 1, option=2, verb=0, ctx=ctx)
    assert e.value.code == -5
    with pytest.raises(ps.PstError) as e:
        ps.somf3dc(d, di, dx, 2, 2, 0.01, 2, option=3, verb=0, ctx=ctx)
    assert e.value.code == -1


@pytest.mark.parametrize("name", golden_names("somean2dadj_"))
def test_somean2d_adjoint_golden_bit_exact(ctx, name):
    """somean2dc(adj=1): adjoint chain spray, level by level, in the reference's accumulation order."""
    import pyseistr_b200 as ps
    g = golden(name)
    out = ps.somean2dc(g["dn"], g["dip"], int(g["ns"]), int(g["order"]), float(g["eps"]), adj=1, verb=0, ctx=ctx)
    assert np.array_equal(out, g["out"])


def test_somean2d_adjoint_dot_product(ctx):
    """<S x, y> = <x, S' y> (pwsmooth_lop is a linear operator for fixed slopes)."""
    import pyseistr_b200 as ps
    n1, n2 = 96, 70
    p, _ = synth.smooth_dips(n1, n2, 1, seed=66, amp=0.7)
    p = np.asarray(p).reshape(n1, n2)
    rng = np.random.default_rng(67)
    x = rng.standard_normal((n1, n2)).astype(np.float32)
    y = rng.standard_normal((n1, n2)).astype(np.float32)
    sx = ps.somean2dc(x, p, 4, 2, 0.01, adj=0, verb=0, ctx=ctx).astype(np.float64)
    sty = ps.somean2dc(y, p, 4, 2, 0.01, adj=1, verb=0, ctx=ctx).astype(np.float64)
    a, b = float((sx * y).sum()), float((x * sty).sum())
    assert abs(a - b) <= 2e-4 * max(abs(a), abs(b)), (a, b)


# ------------------------------------------------------------------ interpolation
@pytest.mark.parametrize("name", golden_names("soint3d_"))
def test_soint3d_golden(ctx, name):
    """Vectors are reference-ordered (bit-identical operators, gather adjoint in scatter order); the five
    CG dots are double trees instead of sequential doubles: relative L2 <= 1e-5."""
    import pyseistr_b200 as ps
    g = golden(name)
    out = ps.soint3dc(g["din"], g["mask"], g["dipi"], g["dipx"], order=int(g["order"]), niter=int(g["niter"]),
                      hasmask=int(g["hasmask"]), verb=0, ctx=ctx)
    assert out.shape == g["out"].shape
    assert rel_l2(out, g["out"]) <= TOL, rel_l2(out, g["out"])


def test_soint3d_vs_oracle_and_recovers_plane_waves(ctx, port):
    import pyseistr_b200 as ps
    d = synth.cube(64, 20, 12, seed=81, noise=0.0)
    pi, px = port.dip3dc(d, 3, 8, 2, rect=(4, 4, 3))
    keep = np.random.default_rng(82).random((20, 12)) > 0.5
    mask = np.zeros_like(d)
    mask[:, keep] = 1
    d0 = d * mask
    got = ps.soint3dc(d0, mask, pi, px, order=2, niter=25, verb=0, ctx=ctx)
    want = port.soint3dc(d0, mask, pi, px, order=2, niter=25)
    assert rel_l2(got, want) <= TOL, rel_l2(got, want)
    assert np.array_equal(got[:, keep], d0[:, keep])                      # known traces untouched
    assert np.linalg.norm(got - d) < 0.25 * np.linalg.norm(d0 - d)        # missing traces filled in


@pytest.mark.parametrize("name", golden_names("sint3d_"))
def test_sint3d_golden(ctx, name):
    """csint3d: forward and adjoint plane-wave smoothers are reference-ordered; the CG dots are double
    trees instead of sequential doubles: relative L2 <= 1e-5."""
    import pyseistr_b200 as ps
    g = golden(name)
    out = ps.sint3dc(g["din"], g["mask"], g["dipi"], g["dipx"], niter=int(g["niter"]), eps=float(g["eps"]),
                     ns1=int(g["ns1"]), ns2=int(g["ns2"]), order1=int(g["order1"]), order2=int(g["order2"]),
                     verb=0, ctx=ctx)
    assert out.shape == g["out"].shape
    assert rel_l2(out, g["out"]) <= TOL, rel_l2(out, g["out"])


def test_sint3d_vs_oracle_and_fills_gaps(ctx, port):
    import pyseistr_b200 as ps
    d = synth.cube(48, 40, 36, seed=84, noise=0.0)
    pi, px = port.dip3dc(d, 3, 8, 2, rect=(4, 4, 3))
    keep = np.random.default_rng(85).random((40, 36)) > 0.5
    mask = np.zeros_like(d)
    mask[:, keep] = 1
    d0 = d * mask
    got = ps.sint3dc(d0, mask, pi, px, niter=6, ns1=2, ns2=3, order1=2, order2=1, verb=0, ctx=ctx)
    want = port.sint3dc(d0, mask, pi, px, niter=6, ns1=2, ns2=3, order1=2, order2=1)
    assert rel_l2(got, want) <= TOL, rel_l2(got, want)
    assert np.linalg.norm(got - d) < 0.6 * np.linalg.norm(d0 - d)         # missing traces filled in
    # one iteration isolates the operators (x = S p, one adjoint, one forward): must be tighter still
    got1 = ps.sint3dc(d0, mask, pi, px, niter=1, ns1=1, ns2=2, order1=1, order2=2, verb=0, ctx=ctx)
    want1 = port.sint3dc(d0, mask, pi, px, niter=1, ns1=1, ns2=2, order1=1, order2=2)
    assert rel_l2(got1, want1) <= 1e-6, rel_l2(got1, want1)


def test_soint3d_noise_rhs_matches_oracle(ctx, port):
    """var > 0: the right-hand side is a*N(0,1) from MT19937(seed) + Box-Muller drawn like the reference
    (soint3d_cfuns.c:2304-2402); same seed, same interpolation (rel. L2 <= 1e-5)."""
    import pyseistr_b200 as ps
    g = golden("soint3d_o2n20")
    for var, seed in ((0.01, 202223), (0.3, 7)):
        got = ps.soint3dc(g["din"], g["mask"], g["dipi"], g["dipx"], order=2, niter=8, var=var, seed=seed, verb=0, ctx=ctx)
        want = port.soint3dc(g["din"], g["mask"], g["dipi"], g["dipx"], order=2, niter=8, var=var, seed=seed)
        assert rel_l2(got, want) <= TOL, (var, seed, rel_l2(got, want))
    other = ps.soint3dc(g["din"], g["mask"], g["dipi"], g["dipx"], order=2, niter=8, var=0.3, seed=8, verb=0, ctx=ctx)
    assert rel_l2(other, want) > 1e-3                                   # a different seed is a different answer


@pytest.mark.parametrize("njs,order", [((2, 1), 2), ((1, 3), 1), ((2, 2), 2)])
def test_soint3d_dealiasing_strides(ctx, port, njs, order):
    """njs != 1: stencil shifts (w - nw) * nj on rows [nw*nj, n1 - nw*nj) (allpass3_lop soint3d_cfuns.c:640-706)."""
    import pyseistr_b200 as ps
    g = golden("soint3d_o2n20")
    got = ps.soint3dc(g["din"], g["mask"], g["dipi"], g["dipx"], order=order, niter=8, njs=list(njs), verb=0, ctx=ctx)
    want = port.soint3dc(g["din"], g["mask"], g["dipi"], g["dipx"], order=order, niter=8, njs=njs)
    assert rel_l2(got, want) <= TOL, rel_l2(got, want)


def test_soint3d_unsupported_options_refused(ctx):
    import pyseistr_b200 as ps
    d = synth.cube(20, 6, 4, seed=83)
    for kw in (dict(drift=1),):
        with pytest.raises(ps.PstError) as e:
            ps.soint3dc(d, np.ones_like(d), 0 * d, 0 * d, niter=2, verb=0, ctx=ctx, **kw)
        assert e.value.code == -5


# ------------------------------------------------------------------ config 1: 2-D DAS-style panel
def test_das_panel_config1(ctx, port):
    """3000x860 panel with the demo's parameters (reference demos/test_pyseistr_das_massive.py:197-198).
    somf2d must be bit-exact.  dip2d with rect=[40,40] is an ill-conditioned shaping CG: the reference's
    OWN result moves by ~6e-5 when only the association of its double dot products changes (measured with
    the oracle's probe, pso_set_dot_mode), so the 1e-5 tolerance is widened to 3x that measured self-noise
    for this configuration — and only for it."""
    import pyseistr_b200 as ps
    n1, n2 = 3000, 860
    d = synth.erratic(synth.cube(n1, n2, 1, seed=5), ntraces=20)
    kw = dict(niter=2, liter=10, order=2, rect=[40, 40, 1])
    p = ps.dip2dc(d, verb=0, ctx=ctx, **kw)
    po = port.dip2dc(d, **kw)
    port.set_dot_mode(1)
    try:
        pb = port.dip2dc(d, **kw)
    finally:
        port.set_dot_mode(0)
    self_noise = rel_l2(pb, po)
    assert rel_l2(p, po) <= max(TOL, 3.0 * self_noise), (rel_l2(p, po), self_noise)
    f = ps.somf2dc(d, po, 8, 2, 0.01, verb=0, ctx=ctx)
    assert np.array_equal(f, port.somf2dc(d, po, 8, 2, 0.01))


def test_interpolators_at_scale(ctx):
    """200x128x64, half of the traces missing: soint3dc and sint3dc run through their chunked paths, return finite
    volumes and fill the gaps (properties instead of the oracle, which needs minutes here)."""
    import pyseistr_b200 as ps
    n1, n2, n3 = 200, 128, 64
    d = synth.cube(n1, n2, n3, seed=91, noise=0.0)
    di, dx = ps.dip3dc(d, verb=0, ctx=ctx)
    keep = np.random.default_rng(92).random((n2, n3)) > 0.5
    mask = np.zeros_like(d)
    mask[:, keep] = 1
    d0 = d * mask
    gap = np.linalg.norm(d0 - d)
    a = ps.soint3dc(d0, mask, di, dx, order=2, niter=20, verb=0, ctx=ctx)
    assert np.isfinite(a).all() and np.array_equal(a[:, keep], d0[:, keep])
    assert np.linalg.norm(a - d) < 0.35 * gap
    b = ps.sint3dc(d0, mask, di, dx, niter=8, ns1=2, ns2=2, order1=2, order2=2, verb=0, ctx=ctx)
    assert np.isfinite(b).all() and np.linalg.norm(b - d) < 0.6 * gap
    b2 = ps.sint3dc(d0, mask, di, dx, niter=8, ns1=2, ns2=2, order1=2, order2=2, verb=0, ctx=ctx)
    assert np.array_equal(b, b2)                                        # run-to-run reproducible


# ------------------------------------------------------------------ properties at scale
def test_pipeline_properties_at_scale(ctx):
    """200x128x64 (the survey's proxy cube; the oracle needs ~30 s there so properties are used
    instead): determinism, median of a constant-dip plane wave keeps the plane wave, r=1 median
    is independent of dipx (Q3), all-ones mean gives 1 / 6/9 / 4/9 (Q4)."""
    import pyseistr_b200 as ps
    n1, n2, n3 = 200, 128, 64
    d = synth.cube(n1, n2, n3, seed=71)
    di, dx = ps.dip3dc(d, verb=0, ctx=ctx)
    di2, dx2 = ps.dip3dc(d, verb=0, ctx=ctx)
    assert np.array_equal(di, di2) and np.array_equal(dx, dx2)          # run-to-run reproducible
    assert np.isfinite(di).all() and np.abs(di).max() < 4.0
    a = ps.somf3dc(d, di, dx, 1, 1, 0.01, 2, verb=0, ctx=ctx)
    b = ps.somf3dc(d, di, 0 * dx + 0.25, 1, 1, 0.01, 2, verb=0, ctx=ctx)
    assert np.array_equal(a, b)
    ones = np.ones((n1, 16, 8), np.float32)
    z = np.zeros_like(ones)
    m = ps.somean3dc(ones, z, z, 1, 1, 0.01, 1, ctx=ctx)
    assert np.allclose(m[5:-5, 5, 4], 1.0, atol=1e-4)
    assert np.allclose(m[5:-5, 0, 4], 6.0 / 9.0, atol=1e-4)
    assert np.allclose(m[5:-5, 0, 0], 4.0 / 9.0, atol=1e-4)
    # filtering reduces the erratic noise energy
    de = synth.erratic(d, ntraces=40)
    f = ps.somf3dc(de, di, dx, 2, 2, 0.01, 2, verb=0, ctx=ctx)
    assert np.linalg.norm(f - d) < 0.7 * np.linalg.norm(de - d)


# ------------------------------------------------------------------ added last in round 1 (kept at the end of the file)
def test_soint2d_default_path(ctx, port):
    """soint2dc (one slope field, no preconditioner) runs through pst_soint3d on an (n1, n2, 1) volume; the equivalence
    csoint2d == csoint3d(n3 = 1) is pinned on the compiled reference in tests/test_oracle.py."""
    import pyseistr_b200 as ps
    d = np.asarray(synth.cube(96, 40, 1, seed=95, noise=0.0)).reshape(96, 40)
    p2 = port.dip2dc(d, 2, 10, 2, 0.01, 1, 1e-6, [7, 7, 1])
    keep = np.random.default_rng(96).random(40) > 0.5
    mask = np.zeros_like(d)
    mask[:, keep] = 1
    got = ps.soint2dc(d * mask, mask, p2, order=2, niter=15, verb=0, ctx=ctx)
    want = port.soint2dc(d * mask, mask, p2, order=2, niter=15)
    assert got.shape == (96, 40)
    assert rel_l2(got, want) <= TOL, rel_l2(got, want)
    for name in golden_names("soint2d_"):                      # csoint2d outputs of the compiled reference
        g = golden(name)
        got = ps.soint2dc(g["din"], g["mask"], g["dip"], order=int(g["order"]), niter=int(g["niter"]),
                          njs=[int(v) for v in g["njs"]], hasmask=int(g["hasmask"]), verb=0, ctx=ctx)
        assert rel_l2(got, g["out"]) <= TOL, (name, rel_l2(got, g["out"]))


# ------------------------------------------------------------------ round 2: oracle comparisons at realistic sizes
def test_pipeline_vs_oracle_200x128x64(ctx, port):
    """The survey's proxy cube (1.6 M voxels, ~30 s of oracle time per pass).  At this size the reference's OWN dips
    move by 1e-5 .. 3e-5 when nothing but the association of its double-precision dot products changes (the oracle's
    probe pso_set_dot_mode(1): sums in blocks of 4096; beta = sr.sr + eps (sp.sp - sx.sx) cancels, so last-bit changes
    of the sums flip roundings of the float step lengths), i.e. the 1e-5 tolerance is below the reference's arithmetic
    noise floor and no parallel reduction can meet it against the SEQUENTIAL sums.  Contract checked here:
      * dips within 1e-5 of the reference with re-associated dot sums (observed: identical to 14 digits),
      * within max(1e-5, 3 x the reference's self-noise) of the sequential reference,
      * the SAME data-dependent control flow (CG iterations incl. early exits, line-search evaluations),
      * somf3dc / somean3dc on the oracle's dips bit-exact."""
    import pyseistr_b200 as ps
    n1, n2, n3 = 200, 128, 64
    d = synth.erratic(synth.cube(n1, n2, n3, seed=171), ntraces=60)
    oi, ox = port.dip3dc(d)
    want = port.dip_counts()
    port.set_dot_mode(1)
    try:
        bi, bx = port.dip3dc(d)
    finally:
        port.set_dot_mode(0)
    noise = max(rel_l2(bi, oi), rel_l2(bx, ox))
    di, dx = ps.dip3dc(d, verb=0, ctx=ctx)
    st = ctx.stats()
    e_seq, e_blk = max(rel_l2(di, oi), rel_l2(dx, ox)), max(rel_l2(di, bi), rel_l2(dx, bx))
    print(f"[200x128x64] dip rel-L2 vs sequential reference {e_seq:.3e}, vs re-associated reference {e_blk:.3e}; "
          f"reference self-noise {noise:.3e}; bit-identical to the re-associated reference: "
          f"{bool(np.array_equal(di, bi) and np.array_equal(dx, bx))}")
    assert e_blk <= TOL, e_blk
    assert e_seq <= max(TOL, 3.0 * noise), (e_seq, noise)
    got = {k: int(st[k]) for k in want}
    assert got == want, (got, want)
    f = ps.somf3dc(d, oi, ox, 2, 2, 0.01, 2, verb=0, ctx=ctx)
    assert np.array_equal(f, port.somf3dc(d, oi, ox, 2, 2, 0.01, 2))
    m = ps.somean3dc(d, oi, ox, 2, 2, 0.01, 2, ctx=ctx)
    assert np.array_equal(m, port.somean3dc(d, oi, ox, 2, 2, 0.01, 2))


def test_das_panel_upscaled_30000x1280(ctx, port):
    """BASELINE.json config 2 at its upscaled size: somf2dc (ns 8) bit-exact on the oracle's slopes; dip2dc with the
    demo's rect=[40,40,1] is the ill-conditioned shaping CG of test_das_panel_config1: its deviation is REPORTED next
    to the reference's own re-association noise and bounded by max(1e-5, 3 x that noise)."""
    import pyseistr_b200 as ps
    n1, n2 = 30000, 1280
    d = synth.erratic(synth.cube(n1, n2, 1, seed=15, nevents=6), ntraces=25)
    kw = dict(niter=2, liter=10, order=2, rect=[40, 40, 1])
    po = port.dip2dc(d, **kw)
    p = ps.dip2dc(d, verb=0, ctx=ctx, **kw)
    dev, noise = rel_l2(p, po), None
    if dev > TOL:                                # the second oracle pass (~80 s) only when the plain tolerance is missed
        port.set_dot_mode(1)
        try:
            pb = port.dip2dc(d, **kw)
        finally:
            port.set_dot_mode(0)
        noise = rel_l2(pb, po)
    print(f"[30000x1280 panel] dip2dc rel-L2 vs oracle {dev:.3e}; oracle self-noise (dot re-association) {noise}")
    assert dev <= max(TOL, 3.0 * (noise or 0.0)), (dev, noise)
    f = ps.somf2dc(d, po, 8, 2, 0.01, verb=0, ctx=ctx)
    assert np.array_equal(f, port.somf2dc(d, po, 8, 2, 0.01))


def test_sint2d_through_one_plane_sint3d(ctx, port):
    """sint2dc (reference sint.py:61-94) = sint3dc on an (n1, n2, 1) volume without xline smoothing: the one-trace-per-panel
    spray path, against the oracle's csint2d restatement (itself pinned to the compiled reference)."""
    import pyseistr_b200 as ps
    d = np.asarray(synth.cube(96, 40, 1, seed=97, noise=0.0)).reshape(96, 40)
    p2 = port.dip2dc(d, 2, 10, 2, 0.01, 1, 1e-6, [7, 7, 1])
    keep = np.random.default_rng(98).random(40) > 0.5
    mask = np.zeros_like(d)
    mask[:, keep] = 1
    for ns, order, niter in ((1, 1, 6), (2, 2, 8), (3, 1, 5)):
        got = ps.sint2dc(d * mask, mask, p2, niter=niter, eps=0.01, ns=ns, order=order, verb=0, ctx=ctx)
        want = port.sint2dc(d * mask, mask, p2, niter=niter, eps=0.01, ns=ns, order=order)
        assert got.shape == (96, 40)
        assert rel_l2(got, want) <= TOL, (ns, order, rel_l2(got, want))


@pytest.mark.parametrize("name", golden_names("paint2d_"))
def test_paint2d_golden_bit_exact(ctx, name):
    """pwpaintc (plane-wave painting, reference rgt.py -> cpaint2d): one dependent chain of trace predictions."""
    import pyseistr_b200 as ps
    g = golden(name)
    out = ps.pwpaintc(g["dip"], g["trace"], order=int(g["order"]), i0=int(g["i0"]), eps=float(g["eps"]), ctx=ctx)
    assert np.array_equal(out, g["out"]), rel_l2(out, g["out"])


def test_rgt_vs_oracle(ctx, port):
    import pyseistr_b200 as ps
    p = synth.smooth_dips(150, 37, 1, seed=41)[0]
    for order, i0 in ((1, 0), (2, 18), (2, 36)):
        t = ps.rgt(p, o1=-0.2, d1=0.004, order=order, i0=i0, eps=0.1, ctx=ctx)
        seed = np.linspace(0, 0.004 * 149, 150) - 0.2
        assert np.array_equal(t, port.pwpaintc(p, seed, order, i0, 0.1)), (order, i0)


def test_soint2dc_twoplane_without_preconditioner_returns_its_input(ctx):
    """csoint2d(twoplane=1, prec=0) is a no-op in the reference (its solver call is commented out, soint2d_cfuns.c:2354-2356,
    :2389-2391; pinned on the compiled reference by the CPU suite): the drop-in hands a copy of the input back."""
    import pyseistr_b200 as ps
    n1, n2 = 60, 24
    clean = np.asarray(synth.cube(n1, n2, 1, seed=5, noise=0.0)).reshape(n1, n2)
    mask = np.zeros_like(clean)
    mask[:, ::2] = 1
    gaps = np.float32(clean * mask)
    dip = np.asarray(synth.smooth_dips(n1, n2, 1, seed=5)[0]).reshape(n1, n2)
    two = np.stack([dip, 0.5 * dip], axis=2)
    got = ps.soint2dc(gaps, mask, two, order=1, niter=20, twoplane=1, verb=0, ctx=ctx)
    assert np.array_equal(got, gaps) and got is not gaps

