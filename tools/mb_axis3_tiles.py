"""The tile kernels of the distributed axis-3 smoothing pass on ONE GPU (PST_TRI3_SOLO=1: whole axis, no neighbours):
time per pass and bit-exactness against the default single-GPU smoother.  Slab shape of 8 ranks by default.
    PST_TRI3_WMAX=64 python tools/mb_axis3_tiles.py [n1 n2 n3 r3 passes]"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def run(solo):
    import pyseistr_b200 as ps
    from pyseistr_b200 import _lib
    a = [int(v) for v in sys.argv[2:]] if solo is not None else []
    n1, n2, n3, r3, passes = (a + [1000, 1024, 128, 5, 20][len(a):])[:5]
    ctx = ps.Context(0)
    lib = ctx.lib
    n = n1 * n2 * n3
    x = np.random.default_rng(5).standard_normal(n).astype(np.float32)
    d = ctx.alloc(4 * n)
    ctx.h2d(d, x)
    _lib.check(lib.pst_smoothcf_dev(ctx.handle, d, n1, n2, n3, 1, 0, 1, 1, r3, 0, 0, 0, 0, 0, 0))
    y = np.empty_like(x)
    ctx.d2h(y, d)
    ctx.set_profile(True)
    ctx.reset_stats()
    for _ in range(passes):
        _lib.check(lib.pst_smoothcf_dev(ctx.handle, d, n1, n2, n3, 1, 0, 1, 1, r3, 0, 0, 0, 0, 0, 0))
    ctx.sync()
    st = ctx.stats()
    i = _lib.KERNEL_CLASSES.index("tri_axis3")
    ms, nl = st["class_ms"][i], st["class_launches"][i]
    print(f"solo={solo} {n1}x{n2}x{n3} r3={r3}: {ms / passes:.3f} ms per pass ({nl // passes} launches), "
          f"{8.0 * n / (ms / passes * 1e-3) / 1e9:.0f} GB/s algorithmic (8 B/voxel)", flush=True)
    np.save(f"/tmp/mb_axis3_{solo}.npy", y)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] in ("0", "1", "2"):
        run(int(sys.argv[1]))
    else:
        # 0: the default single-GPU smoother (reference bits); 1: shared-memory tile kernels (F staged through HBM) at three
        # tile widths; 2: register kernels as an interior rank runs them (TIMING ONLY)
        for solo, extra in (("0", {}), ("1", {"PST_TRI3_WMAX": "128"}), ("1", {"PST_TRI3_WMAX": "64"}), ("1", {"PST_TRI3_WMAX": "32"}),
                            ("1", {"PST_TRI3_WMAX": "64", "PST_TRI3_REV": "1"}), ("2", {}), ("2", {"PST_TRI3_REV": "1"})):
            env = dict(os.environ, PST_TRI3_SOLO=solo, **extra)
            print(f"-- PST_TRI3_SOLO={solo} {extra}", flush=True)
            subprocess.run([sys.executable, __file__, solo] + sys.argv[1:], env=env, check=True)
            if solo == "1":
                a, b = np.load("/tmp/mb_axis3_0.npy"), np.load("/tmp/mb_axis3_1.npy")
                print("   tile kernels == default smoother, bit for bit:", bool(np.array_equal(a, b)), flush=True)
