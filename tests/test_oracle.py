"""CPU suite: the oracle restatement (oracle/pst_oracle.c) is pinned bit-for-bit against the
golden fixtures generated from the unmodified reference C, and — when oracle/_ref is present
(build container, or shipped prebuilt to the GPU box) — against the compiled reference on
fresh seeded inputs."""
import numpy as np
import pytest

from conftest import golden, golden_names
from pyseistr_b200 import synth


@pytest.mark.parametrize("name", golden_names("dip3d_"))
def test_port_dip3d_matches_golden(port, name):
    g = golden(name)
    mask = g.get("mask")
    di, dx = port.dip3dc(g["din"], int(g["niter"]), int(g["liter"]), int(g["order"]),
                         rect=[int(v) for v in g["rect"]], mask=mask)
    assert np.array_equal(di, g["dipi"])
    assert np.array_equal(dx, g["dipx"])


def test_port_dip2d_matches_golden(port):
    g = golden("dip2d")
    p = port.dip2dc(g["din"], int(g["niter"]), int(g["liter"]), int(g["order"]), rect=[int(v) for v in g["rect"]])
    assert np.array_equal(p, g["dip"])


@pytest.mark.parametrize("name", golden_names("somean3d_") + golden_names("somf3d_"))
def test_port_spray3d_matches_golden(port, name):
    g = golden(name)
    fn = port.somean3dc if name.startswith("somean") else port.somf3dc
    out = fn(g["dn"], g["dipi"], g["dipx"], int(g["r1"]), int(g["r2"]), 0.01, int(g["order"]))
    assert np.array_equal(out, g["out"])


@pytest.mark.parametrize("name", golden_names("somean2d_") + golden_names("somf2d_"))
def test_port_spray2d_matches_golden(port, name):
    g = golden(name)
    fn = port.somean2dc if name.startswith("somean") else port.somf2dc
    out = fn(g["dn"], g["dip"], int(g["ns"]), int(g["order"]), float(g["eps"]))
    assert np.array_equal(out, g["out"])


@pytest.mark.parametrize("name", golden_names("somean2dadj_"))
def test_port_somean2d_adjoint_matches_golden(port, name):
    g = golden(name)
    out = port.somean2dc(g["dn"], g["dip"], int(g["ns"]), int(g["order"]), float(g["eps"]), adj=1)
    assert np.array_equal(out, g["out"])


@pytest.mark.parametrize("name", golden_names("soint3d_"))
def test_port_soint3d_matches_golden(port, name):
    g = golden(name)
    out = port.soint3dc(g["din"], g["mask"], g["dipi"], g["dipx"], order=int(g["order"]), niter=int(g["niter"]),
                        hasmask=int(g["hasmask"]))
    assert np.array_equal(out, g["out"])


@pytest.mark.parametrize("name", golden_names("soint2d_"))
def test_port_soint2d_matches_golden(port, name):
    """csoint2d of the reference (default path) against the one-plane soint3d restatement."""
    g = golden(name)
    out = port.soint2dc(g["din"], g["mask"], g["dip"], order=int(g["order"]), niter=int(g["niter"]),
                        njs=[int(v) for v in g["njs"]], hasmask=int(g["hasmask"]))
    assert np.array_equal(out, g["out"])


def test_port_soint3d_noise_matches_compiled_reference(port):
    """var > 0 (MT19937 + Box-Muller right-hand side): bit-identical to the compiled reference."""
    ref = _ref_or_skip()
    g = golden("soint3d_o2n20")
    for var, seed in ((0.02, 202223), (0.5, 11)):
        a = port.soint3dc(g["din"], g["mask"], g["dipi"], g["dipx"], order=2, niter=6, var=var, seed=seed)
        b = ref.soint3dc(g["din"], g["mask"], g["dipi"], g["dipx"], order=2, niter=6, var=var, seed=seed)
        assert np.array_equal(a, b)


def test_port_soint3d_strides_match_compiled_reference(port):
    ref = _ref_or_skip()
    g = golden("soint3d_o2n20")
    for njs, order in (((2, 1), 2), ((1, 3), 1), ((2, 2), 2)):
        a = port.soint3dc(g["din"], g["mask"], g["dipi"], g["dipx"], order=order, niter=6, njs=njs)
        b = ref.soint3dc(g["din"], g["mask"], g["dipi"], g["dipx"], order=order, niter=6, njs=njs)
        assert np.array_equal(a, b)


@pytest.mark.parametrize("name", golden_names("sint3d_"))
def test_port_sint3d_matches_golden(port, name):
    g = golden(name)
    out = port.sint3dc(g["din"], g["mask"], g["dipi"], g["dipx"], niter=int(g["niter"]), eps=float(g["eps"]),
                       ns1=int(g["ns1"]), ns2=int(g["ns2"]), order1=int(g["order1"]), order2=int(g["order2"]))
    assert np.array_equal(out, g["out"])


@pytest.mark.parametrize("name", golden_names("smooth_"))
def test_port_smooth_matches_golden(port, name):
    g = golden(name)
    assert np.array_equal(port.smooth3(g["x"], [int(v) for v in g["rect"]]), g["out"])


def test_port_smooth_repeat_matches_compiled_reference(port):
    ref = _ref_or_skip()
    x = synth.cube(30, 12, 6, seed=13)
    for rect, rep in (([5, 3, 4], 2), ([2, 15, 9], 3), ([3, 3, 1], 4)):
        assert np.array_equal(port.smooth3(x, rect, rep), ref.smoothc(x, rect, repeat=rep))


def test_port_smoothcf_options_match_compiled_reference(port):
    """every option of smoothcf (adj, repeat, diff, box per axis): bit-identical to the compiled reference."""
    ref = _ref_or_skip()
    x = synth.cube(30, 12, 6, seed=13)
    for rect in ([5, 3, 4], [2, 15, 9]):
        for adj in (0, 1):
            for diff in ((0, 0, 0), (1, 0, 0), (0, 1, 1)):
                for box in ((0, 0, 0), (1, 1, 1), (0, 0, 1)):
                    a = port.smoothc(x, rect, diff, box, 2, adj)
                    b = ref.smoothc(x, rect, adj=adj, repeat=2, diff=diff, box=box)
                    assert np.array_equal(a, b), (rect, adj, diff, box)


def test_port_smooth_adj1_matches_compiled_reference(port):
    ref = _ref_or_skip()
    x = synth.cube(30, 12, 6, seed=13)
    for rect, rep in (([5, 3, 4], 1), ([2, 15, 9], 2), ([7, 2, 2], 1)):
        assert np.array_equal(port.smooth3(x, rect, rep, adj=1), ref.smoothc(x, rect, adj=1, repeat=rep))


def test_port_filters_closed_form(port):
    """nw=1 taps have the closed form [(1-p)(2-p)/12, (2+p)(2-p)/6, (1+p)(2+p)/12]."""
    for p in (-0.7, 0.0, 0.3, 1.2):
        a = port.passfilter(1, p)
        want = np.array([(1 - p) * (2 - p) / 12, (2 + p) * (2 - p) / 6, (1 + p) * (2 + p) / 12])
        assert np.allclose(a, want, rtol=1e-6)
        assert abs(a.sum() - 1.0) < 1e-6          # allpass: taps sum to one
        d = port.aderfilter(1, p)
        h = 1e-3                                   # aderfilter = -(d/dp) passfilter
        fd = (port.passfilter(1, p + h).astype(np.float64) - port.passfilter(1, p - h)) / (2 * h)
        assert np.allclose(d, -fd, atol=2e-4)


def test_reference_quirks(port):
    """Q3/Q4 of SURVEY: r=1 median ignores dipx; all-ones + zero dip gives 1, 6/9, 4/9."""
    d = synth.cube(24, 9, 7, seed=4)
    pi, px = synth.smooth_dips(24, 9, 7, seed=2)
    a = port.somf3dc(d, pi, px, 1, 1, 0.01, 1)
    b = port.somf3dc(d, pi, 0 * px + 0.3, 1, 1, 0.01, 1)
    assert np.array_equal(a, b)
    ones = np.ones((12, 5, 5), np.float32)
    z = np.zeros_like(ones)
    m = port.somean3dc(ones, z, z, 1, 1, 0.01, 1)
    assert np.allclose(m[:, 2, 2], 1.0, atol=1e-4)
    assert np.allclose(m[:, 0, 2], 6.0 / 9.0, atol=1e-4)
    assert np.allclose(m[:, 0, 0], 4.0 / 9.0, atol=1e-4)


def _ref_or_skip():
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref (compiled reference) not present")
    return ref


def test_port_vs_compiled_reference_3d(port):
    ref = _ref_or_skip()
    d = synth.cube(48, 14, 9, seed=21)
    ri, rx = ref.dip3dc(d, niter=3, liter=8, order=2, rect=[4, 5, 3])
    pi, px = port.dip3dc(d, niter=3, liter=8, order=2, rect=[4, 5, 3])
    assert np.array_equal(ri, pi) and np.array_equal(rx, px)
    de = synth.erratic(d)
    assert np.array_equal(ref.somf3dc(de, ri, rx, 2, 2, 0.01, 2), port.somf3dc(de, ri, rx, 2, 2, 0.01, 2))
    assert np.array_equal(ref.somean3dc(de, ri, rx, 1, 2, 0.01, 2), port.somean3dc(de, ri, rx, 1, 2, 0.01, 2))


def test_port_vs_compiled_reference_2d(port):
    ref = _ref_or_skip()
    d = synth.cube(70, 30, 1, seed=22)
    rp = ref.dip2dc(d, 2, 8, 1, 0.01, 1, 1e-6, [6, 9, 1])
    assert np.array_equal(rp, port.dip2dc(d, 2, 8, 1, 0.01, 1, 1e-6, [6, 9, 1]))
    assert np.array_equal(ref.somf2dc(d, rp, 4, 2, 0.01), port.somf2dc(d, rp, 4, 2, 0.01))
    assert np.array_equal(ref.somean2dc(d, rp, 4, 1, 0.02), port.somean2dc(d, rp, 4, 1, 0.02))


def test_dot_association_probe_is_inert_on_small_cases(port):
    """The sensitivity probe re-associates the double dot products; on the small well-conditioned
    fixtures it must not change a single bit (so tolerances widened by it stay honest)."""
    g = golden("dip3d_default")
    port.set_dot_mode(1)
    try:
        di, dx = port.dip3dc(g["din"], int(g["niter"]), int(g["liter"]), int(g["order"]), rect=[int(v) for v in g["rect"]])
    finally:
        port.set_dot_mode(0)
    assert np.array_equal(di, g["dipi"]) and np.array_equal(dx, g["dipx"])


def test_port_sint2d_matches_compiled_reference(port):
    """csint2d: oracle groundwork for the next round, pinned to the compiled reference."""
    ref = _ref_or_skip()
    try:
        ref.module("soint2dcfun")
    except ImportError:
        pytest.skip("oracle/_ref/soint2dcfun not built")
    d = np.asarray(synth.cube(48, 24, 1, seed=3, noise=0.0)).reshape(48, 24)
    p2 = port.dip2dc(d, 2, 10, 2, 0.01, 1, 1e-6, [7, 7, 1])
    keep = np.random.default_rng(5).random(24) > 0.5
    mask = np.zeros_like(d)
    mask[:, keep] = 1
    for ns, order, niter in ((1, 1, 5), (2, 2, 6)):
        a = port.sint2dc(d * mask, mask, p2, niter=niter, ns=ns, order=order)
        b = ref.sint2dc(d * mask, mask, p2, niter=niter, ns=ns, order=order)
        assert np.array_equal(a, b)


def test_port_soint2d_default_path_is_soint3d_with_one_plane(port):
    """csoint2d (twoplane=0, prec=0) == csoint3d on (n1, n2, 1), bit for bit, on the compiled reference."""
    ref = _ref_or_skip()
    try:
        ref.module("soint2dcfun")
    except ImportError:
        pytest.skip("oracle/_ref/soint2dcfun not built")
    d = np.asarray(synth.cube(48, 24, 1, seed=3, noise=0.0)).reshape(48, 24)
    p2 = port.dip2dc(d, 2, 10, 2, 0.01, 1, 1e-6, [7, 7, 1])
    keep = np.random.default_rng(5).random(24) > 0.5
    mask = np.zeros_like(d)
    mask[:, keep] = 1
    for order, niter, njs, hasmask in ((1, 10, (1, 1), 1), (2, 8, (2, 1), 1), (1, 9, (1, 1), 0)):
        a = ref.soint2dc(d * mask, mask, p2, order=order, niter=niter, njs=njs, hasmask=hasmask)
        b = port.soint2dc(d * mask, mask, p2, order=order, niter=niter, njs=njs, hasmask=hasmask)
        assert np.array_equal(a, b)


def test_ref_sint2d_is_sint3d_with_one_plane():
    """Groundwork for a GPU sint2dc: csint2d == csint3d on (n1, n2, 1) with ns2 = 0 or 1, bit for bit (compiled reference)."""
    ref = _ref_or_skip()
    try:
        ref.module("soint2dcfun")
    except ImportError:
        pytest.skip("oracle/_ref/soint2dcfun not built")
    d = np.asarray(synth.cube(48, 24, 1, seed=3, noise=0.0)).reshape(48, 24)
    p2 = ref.dip2dc(d, 2, 10, 2, 0.01, 1, 1e-6, [7, 7, 1])
    keep = np.random.default_rng(5).random(24) > 0.5
    mask = np.zeros_like(d)
    mask[:, keep] = 1
    r3 = lambda a: np.float32(a).reshape(48, 24, 1)
    for niter, ns, order in ((5, 1, 1), (4, 2, 2)):
        a = ref.sint2dc(d * mask, mask, p2, niter=niter, eps=0.01, ns=ns, order=order)
        for ns2 in (0, 1):
            b = ref.sint3dc(r3(d * mask), r3(mask), r3(p2), r3(np.zeros_like(p2)), niter=niter, eps=0.01, ns1=ns, ns2=ns2,
                            order1=order, order2=order).reshape(48, 24)
            assert np.array_equal(a, b)


_SVMF_PROBE = r"""
import sys
sys.path.insert(0, {root!r})
import numpy as np
from oracle import port, ref
from pyseistr_b200 import synth
bad = tot = 0
for seed, shape, r in ((1, (64, 24, 10), (2, 2)), (3, (80, 20, 16), (2, 1)), (7, (96, 25, 11), (3, 1))):
    d = synth.erratic(synth.cube(*shape, seed=seed), ntraces=10)
    di, dx = synth.smooth_dips(*shape, seed=seed)
    a = port.somf3dc(d, di, dx, r[0], r[1], 0.01, 2, option=2)
    b = ref.somf3dc(d, di, dx, r[0], r[1], 0.01, 2, option=2)
    bad += int((a != b).sum()); tot += a.size
for seed, (n1, n2), ns in ((1, (200, 60), 3), (2, (300, 80), 8)):
    d = synth.erratic(synth.cube(n1, n2, 1, seed=seed), ntraces=5)
    p = synth.smooth_dips(n1, n2, 1, seed=seed)[0]
    a = port.somf2dc(d, p, ns, 2, 0.01, option=2)
    b = ref.somf2dc(d, p, ns, 2, 0.01, option=2)
    bad += int((a != b).sum()); tot += a.size
print("SVMFPROBE", bad, tot)
"""


def test_port_svmf_defined_behaviour_matches_compiled_reference(port):
    """option=2 (SVMF): the reference's first pass reads one row past its extended panel for the last slot row
    (sof3d_cfuns.c:1283), so heap contents enter its panel average and its output depends on what the process
    allocated before (measured: the same call differs in 2.8 % of the samples between a fresh process and this pytest
    process).  The oracle replicates the edge row there instead.  Pinned two ways: in a FRESH process the compiled
    reference returns exactly the oracle's result on five probes (1.0e5 samples), and the committed fixtures (generated
    from the compiled reference) are reproduced bit for bit."""
    import os
    import subprocess
    import sys
    _ref_or_skip()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # even a fresh process is not fully deterministic (seen once in ~10 runs under load: the over-read row is whatever
    # the allocator hands out), so the probe gets a few fresh processes; ONE clean-heap run that equals the
    # restatement on every sample is the pin, the fixtures below are the repeatable part
    seen = []
    for _ in range(4):
        r = subprocess.run([sys.executable, "-c", _SVMF_PROBE.format(root=root)], capture_output=True, text=True, timeout=600)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("SVMFPROBE")]
        assert r.returncode == 0 and line, r.stderr[-2000:]
        bad, tot = (int(v) for v in line[-1].split()[1:])
        seen.append((bad, tot))
        if bad == 0:
            break
    assert seen[-1][0] == 0 and seen[-1][1] > 100000, seen
    for name in golden_names("svmf3d_"):
        g = golden(name)
        assert np.array_equal(port.somf3dc(g["dn"], g["dipi"], g["dipx"], int(g["r1"]), int(g["r2"]), 0.01, int(g["order"]), option=2), g["out"])
    for name in golden_names("svmf2d_"):
        g = golden(name)
        assert np.array_equal(port.somf2dc(g["dn"], g["dip"], int(g["ns"]), int(g["order"]), float(g["eps"]), option=2), g["out"])


@pytest.mark.parametrize("name", golden_names("paint2d_"))
def test_port_paint2d_matches_golden(port, name):
    g = golden(name)
    out = port.pwpaintc(g["dip"], g["trace"], int(g["order"]), int(g["i0"]), float(g["eps"]))
    assert np.array_equal(out, g["out"])


def test_port_paint2d_matches_compiled_reference(port):
    ref = _ref_or_skip()
    try:
        ref.module("paint2dcfun")
    except ImportError:
        pytest.skip("oracle/_ref/paint2dcfun not built")
    for seed, (n1, n2), order, i0, eps in ((1, (120, 40), 1, 0, 0.01), (2, (200, 64), 2, 30, 0.1), (3, (64, 20), 2, 19, 0.01)):
        p = synth.smooth_dips(n1, n2, 1, seed=seed)[0]
        tr = np.linspace(0, 0.004 * (n1 - 1), n1).astype(np.float32)
        assert np.array_equal(port.pwpaintc(p, tr, order, i0, eps), ref.pwpaintc(p, tr, order, i0, eps))


def test_soint2dc_twoplane_without_preconditioner_returns_its_input():
    """csoint2d(twoplane=1, prec=0): the solver call is commented out in the reference (soint2d_cfuns.c:2354-2356,
    :2389-2391), the model comes back unchanged (the drop-in's side of it: tests/test_gpu_parity.py)."""
    ref = _ref_or_skip()
    try:
        m = ref.module("soint2dcfun")
    except ImportError:
        pytest.skip("oracle/_ref/soint2dcfun not built")
    import pyseistr_b200 as ps
    n1, n2 = 60, 24
    clean = np.asarray(synth.cube(n1, n2, 1, seed=5, noise=0.0)).reshape(n1, n2)
    keep = np.random.default_rng(6).random(n2) > 0.4
    mask = np.zeros_like(clean)
    mask[:, keep] = 1
    gaps = np.float32(clean * mask)
    dip = np.asarray(synth.smooth_dips(n1, n2, 1, seed=5)[0]).reshape(n1, n2)
    two = np.stack([dip, 0.5 * dip], axis=2)
    F = lambda a: np.asfortranarray(np.float32(a)).ravel(order="F")  # noqa: E731
    with ref.quiet():
        want = np.asarray(m.csoint2d(F(gaps), F(mask), F(two[:, :, 0]), F(two[:, :, 1]), n1, n2, 1, 1, 1, 20, 0, 1, 1, 0, 0))
    want = want.reshape(n1, n2, order="F")
    assert np.array_equal(want, gaps)
    with pytest.raises(NotImplementedError):
        ps.soint2dc(gaps, mask, two, twoplane=1, prec=1)
    with pytest.raises(ValueError):
        ps.soint2dc(gaps, mask, dip, twoplane=1)           # one slope field where two are announced
