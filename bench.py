#!/usr/bin/env python
"""bench.py — throughput (Mvoxels/s) of pyseistr's structure-oriented filtering hot path on B200, next to the
reference C path on the box's host cores.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--shape n1,n2,n3] [--impl ours|reference]

Workloads (BASELINE.json configs; the default is the headline one):

  dip3d_somf3d   dip3dc(defaults: niter 5, liter 10, order 2, rect 5,5,5) + somf3dc(r1=r2=2, order 2, MF) using the
                 dips just estimated, on 1000x1024x1024 float32 (configs[4]; fits one B200; 1/2/4/8 GPUs)
  dip2d_somf2d   dip2dc(2,10,2,..,[40,40,1]) + somf2dc(ns 8, order 2) on a DAS-style panel 3000x860
                 (demos/test_pyseistr_das_massive.py:197-198; --shape 30000,1280 = the upscaled panel), 1 GPU
  somean3d       somean3dc(r=2, order 2) on 500x512x512 with dips estimated once in the set-up (1/2 GPUs)
  soint3d        soint3dc(order 2, niter 20) on a 500x512x512 cube with 50 % of its traces removed (1/2/4/8 GPUs)
  sint3d         sint3dc(niter 30, ns1=ns2=2) on the same decimated cube (spray-operator shaping CG)

One "step" = one pass of the workload over one synthetic volume.

  value     whole-job Mvoxels/s with the volume resident in HBM (CUDA events on the library's stream around exactly K
            steps, max over ranks);
  e2e       the same through the reference-facing C-ABI calls with HOST buffers (pinned host -> device copies of the
            inputs and device -> host copies of the results inside the timed region);
  roofline  the dominant kernel class, timed live with CUDA events around each of its launches during the timed steps:
            algorithmic bytes against the measured HBM copy peak (MEASURED_PEAKS.json), or, for the plane-wave
            prediction kernels, algorithmic flops against the FP32 peak at the SM clock sampled during the run;
  parity_vs_n1  (N > 1) the slabs of every output are gathered on rank 0 and compared with ONE single-GPU run of the
            same workload on the same input in the same process: bit-exactness and relative L2 error;
  cpu_baseline  the unmodified reference C (oracle/_ref) on the box's host cores on a bounded sample of the workload.

`--impl reference` times the reference's own CPU implementation (oracle/_ref, all host cores as independent processes —
the reference is single-threaded and non-re-entrant) on a bounded sample whose shape is named in config.workload.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

UNIT = "Mvoxels/s"
DIP_KW = dict(niter=5, liter=10, order=2, rect=(5, 5, 5))
SOMF_KW = dict(r1=2, r2=2, order=2, option=1)
DIP2_KW = dict(niter=2, liter=10, order=2, rect=(40, 40, 1))      # demos/test_pyseistr_das_massive.py:197
SOMF2_KW = dict(ns=8, order=2, eps=0.01, option=1)                 # :198
SOINT_KW = dict(order=2, niter=20)                                 # demos/test_pyseistr_passive_recon3d.py:27
SINT_KW = dict(niter=30, ns1=2, ns2=2, order1=1, order2=1, eps=0.01)
SETUP_DIP_KW = dict(niter=2, liter=5, order=2, rect=(5, 5, 5))     # dips of the spray / interpolation workloads (set-up)
FP32_LANES = 148 * 128                                             # B200: 148 SMs x 128 FP32 lanes


# --------------------------------------------------------------------------- helpers
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


def pinned_array(lib, n_floats):
    """float32 numpy array over cudaMallocHost memory (pinned), via the C-ABI."""
    from pyseistr_b200 import _lib
    p = ctypes.c_void_p()
    _lib.check(lib.pst_host_alloc_pinned(int(n_floats) * 4, ctypes.byref(p)))
    buf = (ctypes.c_float * int(n_floats)).from_address(p.value)
    a = np.frombuffer(buf, dtype=np.float32)
    return a, p


FP = ctypes.POINTER(ctypes.c_float)


def hp(a, off=0):
    """ctypes float* into a host numpy array (element offset)."""
    return ctypes.cast(a.ctypes.data + 4 * off, FP)


def dp(p, off=0):
    """device pointer + element offset"""
    return ctypes.c_void_p(p.value + 4 * off)


# --------------------------------------------------------------------------- workloads
class Workload:
    """A BASELINE.json configuration: synthetic inputs, the device-resident step, the host-buffer step, the reference."""
    name = ""
    metric = ""
    default_shape = (0, 0, 0)
    max_gpus = 8
    inputs = ()            # per-voxel input volumes (names)
    outputs = ()           # (name, components): output volumes of `components` x N floats
    ref_shape = (128, 64, 32)
    cpu_shape = (160, 128, 64)

    def label(self, shape):
        raise NotImplementedError

    def make_inputs(self, env, host):
        """Fill host[name] (pinned float32 arrays of this rank's slab) for every input."""
        raise NotImplementedError

    def run_dev(self, lib, h, n1, n2, n3g, n, d):
        """One step on device pointers d[name] (context handle h; n = floats per volume on this rank)."""
        raise NotImplementedError

    def run_host(self, lib, h, n1, n2, n3g, n, a):
        """One step through the host-pointer entry points on numpy arrays a[name]."""
        raise NotImplementedError

    def copies(self, n):
        """(h2d, d2h) bytes one host step moves for n voxels."""
        raise NotImplementedError

    ref_code = ""          # body of the reference worker: defines step(d...) on a cube `d` of the sample shape


def _dip_dev(lib, h, d_in, n1, n2, n3g, kw, d_out):
    from pyseistr_b200 import _lib
    _lib.check(lib.pst_dip_dev(h, d_in, None, n1, n2, n3g, kw["niter"], kw["liter"], kw["order"], *kw["rect"], 0, d_out))


class Dip3dSomf3d(Workload):
    name = "dip3d_somf3d"
    metric = "Mvoxels/s dip3d+somf3d"
    default_shape = (1000, 1024, 1024)
    inputs = ("din",)
    outputs = (("dip", 2), ("out", 1))

    def label(self, s):
        return f"dip3dc(niter5,liter10,order2,rect5x5x5)+somf3dc(r2x2,order2,MF) on {s[0]}x{s[1]}x{s[2]} f32"

    def make_inputs(self, env, host):
        from pyseistr_b200 import synth
        n1, n2, n3g, z0, z1 = env.n1, env.n2, env.n3g, env.z0, env.z1
        cube = host["din"].reshape((n1, n2, z1 - z0), order="F")
        _, mx = synth.cube_big(n1, n2, n3g, seed=7, out=cube, z0=z0, z1=z1, normalise=False)
        synth.scale_by(cube, env.allmax(mx))

    def run_dev(self, lib, h, n1, n2, n3g, n, d):
        from pyseistr_b200 import _lib
        _dip_dev(lib, h, d["din"], n1, n2, n3g, DIP_KW, d["dip"])
        _lib.check(lib.pst_somf3d_dev(h, d["din"], d["dip"], dp(d["dip"], n), n1, n2, n3g, SOMF_KW["r1"], SOMF_KW["r2"],
                                      2 * SOMF_KW["r1"] * SOMF_KW["r2"] + 1, SOMF_KW["option"], SOMF_KW["order"], d["out"]))

    def run_host(self, lib, h, n1, n2, n3g, n, a):
        from pyseistr_b200 import _lib
        _lib.check(lib.pst_dip(h, hp(a["din"]), None, n1, n2, n3g, DIP_KW["niter"], DIP_KW["liter"], DIP_KW["order"],
                               0.01, 1.0, 1e-6, *DIP_KW["rect"], 0, hp(a["dip"])))
        _lib.check(lib.pst_somf3d(h, hp(a["din"]), hp(a["dip"]), hp(a["dip"], n), n1, n2, n3g, SOMF_KW["r1"],
                                  SOMF_KW["r2"], 2 * SOMF_KW["r1"] * SOMF_KW["r2"] + 1, SOMF_KW["option"],
                                  SOMF_KW["order"], 0.01, 0, hp(a["out"])))

    def copies(self, n):
        return 4 * n * 4, 4 * n * 3

    e2e_path = "pst_dip(host)->pst_somf3d(host): H2D din; D2H dipi,dipx; H2D din,dipi,dipx; D2H out"
    ref_code = """
d = synth.erratic(synth.cube(n1, n2, n3, seed=seed), ntraces=max(4, n2 * n3 // 200))
def step():
    di, dx = impl.dip3dc(d, 5, 10, 2, 0.01, 1, 1e-6, (5, 5, 5), 0)
    impl.somf3dc(d, di, dx, 2, 2, 0.01, 2, 1, 0)
"""
    ref_what = "dip3dc(defaults)+somf3dc(2,2,order 2)"


class Dip2dSomf2d(Workload):
    name = "dip2d_somf2d"
    metric = "Mvoxels/s dip2d+somf2d"
    default_shape = (3000, 860, 1)
    max_gpus = 1
    inputs = ("din",)
    outputs = (("dip", 1), ("out", 1))
    ref_shape = (750, 215, 1)
    cpu_shape = (1000, 430, 1)

    def label(self, s):
        return (f"dip2dc(niter2,liter10,order2,rect40x40)+somf2dc(ns8,order2,MF) on a {s[0]}x{s[1]} f32 panel")

    def make_inputs(self, env, host):
        from pyseistr_b200 import synth
        n1, n2 = env.n1, env.n2
        host["din"][:] = synth.erratic(synth.cube(n1, n2, 1, seed=11, nevents=6), ntraces=max(4, n2 // 50)).flatten(order="F")

    def run_dev(self, lib, h, n1, n2, n3g, n, d):
        from pyseistr_b200 import _lib
        _dip_dev(lib, h, d["din"], n1, n2, 1, DIP2_KW, d["dip"])
        _lib.check(lib.pst_somf2d_dev(h, d["din"], d["dip"], n1, n2, 1, SOMF2_KW["ns"], 2 * SOMF2_KW["ns"] + 1,
                                      SOMF2_KW["option"], SOMF2_KW["order"], SOMF2_KW["eps"], d["out"]))

    def run_host(self, lib, h, n1, n2, n3g, n, a):
        from pyseistr_b200 import _lib
        _lib.check(lib.pst_dip(h, hp(a["din"]), None, n1, n2, 1, DIP2_KW["niter"], DIP2_KW["liter"], DIP2_KW["order"],
                               0.01, 1.0, 1e-6, *DIP2_KW["rect"], 0, hp(a["dip"])))
        _lib.check(lib.pst_somf2d(h, hp(a["din"]), hp(a["dip"]), n1, n2, 1, SOMF2_KW["ns"], 2 * SOMF2_KW["ns"] + 1,
                                  SOMF2_KW["option"], SOMF2_KW["order"], SOMF2_KW["eps"], 0, hp(a["out"])))

    def copies(self, n):
        return 4 * n * 3, 4 * n * 2

    e2e_path = "pst_dip(host, n3=1)->pst_somf2d(host): H2D din; D2H dip; H2D din,dip; D2H out"
    ref_code = """
d = synth.erratic(synth.cube(n1, n2, 1, seed=seed, nevents=6), ntraces=max(4, n2 // 50))
def step():
    pp = impl.dip2dc(d, 2, 10, 2, 0.01, 1, 1e-6, (40, 40, 1), 0)
    impl.somf2dc(d, pp, 8, 2, 0.01, 1, 0)
"""
    ref_what = "dip2dc(2,10,2,..,[40,40,1])+somf2dc(8,2)"


class _WithSetupDips(Workload):
    """Workloads whose slope fields are inputs: estimated ONCE in the (untimed) set-up with pst_dip_dev."""
    decimate = False

    def make_inputs(self, env, host):
        from pyseistr_b200 import synth
        n1, n2, n3g, z0, z1 = env.n1, env.n2, env.n3g, env.z0, env.z1
        nz = z1 - z0
        n = n1 * n2 * nz
        cube = host["din"].reshape((n1, n2, nz), order="F")
        _, mx = synth.cube_big(n1, n2, n3g, seed=9, out=cube, z0=z0, z1=z1, normalise=False,
                               noise=0.0 if self.decimate else 0.05)
        synth.scale_by(cube, env.allmax(mx))
        # slopes of the complete cube
        d_in, d_dip = env.ctx.alloc(4 * n), env.ctx.alloc(8 * n)
        env.ctx.h2d(d_in, host["din"])
        _dip_dev(env.lib, env.ctx.handle, d_in, n1, n2, n3g, SETUP_DIP_KW, d_dip)
        env.ctx.d2h(host["dipi"], d_dip)
        env.ctx.d2h(host["dipx"], dp(d_dip, n))
        env.ctx.free(d_in); env.ctx.free(d_dip)
        if self.decimate:                       # 50 % of the traces removed; the same mask on every rank
            keep = np.random.default_rng(79).random((n2, n3g)) > 0.5
            m = host["mask"].reshape((n1, n2, nz), order="F")
            m[:] = keep[None, :, z0:z1].astype(np.float32)
            cube *= m


class Somean3d(_WithSetupDips):
    name = "somean3d"
    metric = "Mvoxels/s somean3d"
    default_shape = (500, 512, 512)
    max_gpus = 2
    inputs = ("din", "dipi", "dipx")
    outputs = (("out", 1),)

    def label(self, s):
        return f"somean3dc(r2x2,order2) on {s[0]}x{s[1]}x{s[2]} f32 (slopes from dip3dc, estimated in the set-up)"

    def run_dev(self, lib, h, n1, n2, n3g, n, d):
        from pyseistr_b200 import _lib
        _lib.check(lib.pst_somean3d_dev(h, d["din"], d["dipi"], d["dipx"], n1, n2, n3g, 2, 2, 2, d["out"]))

    def run_host(self, lib, h, n1, n2, n3g, n, a):
        from pyseistr_b200 import _lib
        _lib.check(lib.pst_somean3d(h, hp(a["din"]), hp(a["dipi"]), hp(a["dipx"]), n1, n2, n3g, 2, 2, 2, 0.01, 0, hp(a["out"])))

    def copies(self, n):
        return 4 * n * 3, 4 * n

    e2e_path = "pst_somean3d(host): H2D din,dipi,dipx; D2H out"
    ref_code = """
d = synth.cube(n1, n2, n3, seed=seed)
di, dx = synth.smooth_dips(n1, n2, n3, seed=seed)
def step():
    impl.somean3dc(d, di, dx, 2, 2, 0.01, 2, 0)
"""
    ref_what = "somean3dc(2,2,order 2), analytic slopes"


class Soint3d(_WithSetupDips):
    name = "soint3d"
    metric = "Mvoxels/s soint3d"
    default_shape = (500, 512, 512)
    decimate = True
    inputs = ("din", "mask", "dipi", "dipx")
    outputs = (("out", 1),)

    def label(self, s):
        return f"soint3dc(order2,niter20) on {s[0]}x{s[1]}x{s[2]} f32 with 50% of the traces removed"

    def run_dev(self, lib, h, n1, n2, n3g, n, d):
        from pyseistr_b200 import _lib
        _lib.check(lib.pst_soint3d_dev(h, d["din"], d["mask"], d["dipi"], d["dipx"], n1, n2, n3g, SOINT_KW["order"], 1, 1,
                                       SOINT_KW["niter"], 0, 202223, 1, 0.0, 0, d["out"]))

    def run_host(self, lib, h, n1, n2, n3g, n, a):
        from pyseistr_b200 import _lib
        _lib.check(lib.pst_soint3d(h, hp(a["din"]), hp(a["mask"]), hp(a["dipi"]), hp(a["dipx"]), n1, n2, n3g,
                                   SOINT_KW["order"], 1, 1, SOINT_KW["niter"], 0, 202223, 1, 0.0, 0, hp(a["out"])))

    def copies(self, n):
        return 4 * n * 4, 4 * n

    e2e_path = "pst_soint3d(host): H2D din,mask,dipi,dipx; D2H out"
    ref_code = """
c = synth.cube(n1, n2, n3, seed=seed, noise=0.0)
di, dx = synth.smooth_dips(n1, n2, n3, seed=seed, amp=0.3)
keep = np.random.default_rng(79).random((n2, n3)) > 0.5
m = np.zeros_like(c); m[:, keep] = 1
d = c * m
def step():
    impl.soint3dc(d, m, di, dx, order=2, niter=20)
"""
    ref_what = "soint3dc(order 2, niter 20), 50% traces removed, analytic slopes"


class Sint3d(Soint3d):
    name = "sint3d"
    metric = "Mvoxels/s sint3d"

    def label(self, s):
        return f"sint3dc(niter30,ns1=ns2=2,order1) on {s[0]}x{s[1]}x{s[2]} f32 with 50% of the traces removed"

    def run_dev(self, lib, h, n1, n2, n3g, n, d):
        from pyseistr_b200 import _lib
        k = SINT_KW
        _lib.check(lib.pst_sint3d_dev(h, d["din"], d["dipi"], d["dipx"], d["mask"], n1, n2, n3g, k["niter"], k["ns1"], k["ns2"],
                                      k["order1"], k["order2"], 0, k["eps"], d["out"]))

    def run_host(self, lib, h, n1, n2, n3g, n, a):
        from pyseistr_b200 import _lib
        k = SINT_KW
        _lib.check(lib.pst_sint3d(h, hp(a["din"]), hp(a["dipi"]), hp(a["dipx"]), hp(a["mask"]), n1, n2, n3g, k["niter"],
                                  k["ns1"], k["ns2"], k["order1"], k["order2"], 0, k["eps"], hp(a["out"])))

    e2e_path = "pst_sint3d(host): H2D din,dipi,dipx,mask; D2H out"
    ref_shape = (96, 48, 24)
    cpu_shape = (128, 64, 32)
    ref_code = """
c = synth.cube(n1, n2, n3, seed=seed, noise=0.0)
di, dx = synth.smooth_dips(n1, n2, n3, seed=seed, amp=0.3)
keep = np.random.default_rng(79).random((n2, n3)) > 0.5
m = np.zeros_like(c); m[:, keep] = 1
d = c * m
def step():
    impl.sint3dc(d, m, di, dx, niter=30, eps=0.01, ns1=2, ns2=2, order1=1, order2=1, verb=0)
"""
    ref_what = "sint3dc(niter 30, ns 2,2), 50% traces removed, analytic slopes"


WORKLOADS = {w.name: w for w in (Dip3dSomf3d(), Dip2dSomf2d(), Somean3d(), Soint3d(), Sint3d())}


# --------------------------------------------------------------------------- reference arm
REF_WORKER = r"""
import sys, time, json
sys.path.insert(0, {root!r})
import numpy as np
from pyseistr_b200 import synth
kind = {kind!r}
if kind == "reference":
    from oracle import ref as impl
else:
    from oracle import port as impl
n1, n2, n3, seed, steps = {n1}, {n2}, {n3}, {seed}, {steps}
{body}
times = []
for s in range(steps):
    t = time.perf_counter()
    step()
    times.append(time.perf_counter() - t)
print("PSTREF " + json.dumps(times))
"""


def reference_kind():
    from oracle import ref
    if ref.available():
        return "reference"
    from oracle import port
    port.build()
    return "port"


def run_reference_cpu(wl, shape, nproc, steps):
    """nproc independent single-thread processes, each one step of the workload on its own volume of `shape`.
    Returns (per-step aggregate Mvox/s list, kind)."""
    kind = reference_kind()
    n1, n2, n3 = shape
    procs = []
    for r in range(nproc):
        code = REF_WORKER.format(root=ROOT, kind=kind, n1=n1, n2=n2, n3=n3, seed=100 + r, steps=steps, body=wl.ref_code)
        env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, "-c", code], stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True, env=env))
    per = []
    for p in procs:
        out, err = p.communicate()
        got = [ln for ln in out.splitlines() if ln.startswith("PSTREF ")]
        if p.returncode != 0 or not got:
            raise RuntimeError("reference worker failed: " + err[-2000:])
        per.append(json.loads(got[-1][7:]))
    vox = float(n1) * n2 * n3 * nproc
    vals = []
    for s in range(steps):
        tmax = max(w[s] for w in per)
        vals.append(vox / tmax / 1e6)
    return vals, kind


def shape_str(s):
    s = list(s)
    while len(s) > 2 and s[-1] == 1:
        s.pop()
    return "x".join(str(v) for v in s)


def reference_arm(args, wl, rank, world, real_stdout):
    if rank != 0:
        return 0
    nproc = os.cpu_count() or 1
    shape = tuple(int(v) for v in args.ref_shape.split(",")) if args.ref_shape else wl.ref_shape
    steps = args.warmup + args.steps
    t0 = time.perf_counter()
    vals, kind = run_reference_cpu(wl, shape, nproc, steps)
    timed = vals[args.warmup:]
    value = sum(timed) / len(timed)
    vox = float(np.prod(shape)) * nproc
    full = parse_shape(args, wl)
    sample = f"{nproc} independent single-thread processes x {shape_str(shape)} volume each, {wl.ref_what} per step"
    line = {
        "impl": "reference", "metric": wl.metric, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": vox / value / 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        # the workload this line TIMES is the sample (the GPU arm's full volume is named beside it)
        "config": {"workload": f"{wl.ref_what} on {nproc} x {shape_str(shape)} f32 (bounded sample of the GPU arm's "
                               f"{shape_str(full)} workload)",
                   "gpu_arm_workload": wl.label(full),
                   "reference_sample": sample,
                   "why_sample": "the reference is single-threaded C; the full volume needs hours and > 100 GB "
                                 "(dip3d+somf3d at 1000x1024x1024: ~4.7 h, 105 GB, SURVEY §6)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nproc, "kind": kind,
                         "sample": f"{nproc} x {shape_str(shape)}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    emit(real_stdout, line)
    return 0


# --------------------------------------------------------------------------- our arm
def parse_shape(args, wl):
    s = args.shape or os.environ.get("PST_BENCH_SHAPE")
    if s:
        v = tuple(int(x) for x in s.split(","))
        return v if len(v) == 3 else v + (1,)
    return wl.default_shape


def main():
    # everything but the final JSON line goes to stderr (NCCL prints its version to stdout)
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        return _main(real_stdout)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)


def emit(real_stdout, line):
    os.write(real_stdout, (json.dumps(line) + "\n").encode())


class Env:
    pass


class CudaView:
    """Expose a raw device pointer to torch (torch.as_tensor understands __cuda_array_interface__)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


def parity_vs_n1(env, wl, d_slab, torch, dist):
    """Gather every rank's input and output slabs on rank 0, run the workload ONCE on a single-GPU context over the
    whole volume there, and compare: bit-exactness and relative L2 per output."""
    n1, n2, n3g = env.n1, env.n2, env.n3g
    plane = n1 * n2
    rank, world = env.rank, env.world
    bounds = [((n3g * r) // world, (n3g * (r + 1)) // world) for r in range(world)]
    dev = torch.device("cuda", env.local)
    nfull = plane * n3g
    vols = [(nm, 1, True) for nm in wl.inputs] + [(nm, c, False) for nm, c in wl.outputs]
    full = {}
    for nm, comp, _ in vols:
        if rank == 0:
            full[nm] = torch.empty(comp * nfull, dtype=torch.float32, device=dev)
        mine = torch.as_tensor(CudaView(d_slab[nm].value, comp * env.n), device=dev)
        for c in range(comp):
            src = mine[c * env.n:(c + 1) * env.n]
            if rank == 0:
                z0, z1 = bounds[0]
                full[nm][c * nfull + z0 * plane: c * nfull + z1 * plane].copy_(src)
                for r in range(1, world):
                    z0, z1 = bounds[r]
                    dist.recv(full[nm][c * nfull + z0 * plane: c * nfull + z1 * plane], src=r)
            else:
                dist.send(src.contiguous(), dst=0)
    torch.cuda.synchronize()
    res = None
    if rank == 0:
        import pyseistr_b200 as ps
        from pyseistr_b200 import _lib
        c1 = ps.Context(env.local)
        d1 = {nm: ctypes.c_void_p(full[nm].data_ptr()) for nm in wl.inputs}
        ref = {}
        for nm, comp in wl.outputs:
            ref[nm] = torch.empty(comp * nfull, dtype=torch.float32, device=dev)
            d1[nm] = ctypes.c_void_p(ref[nm].data_ptr())
        torch.cuda.synchronize()
        wl.run_dev(c1.lib, c1.handle, n1, n2, n3g, nfull, d1)
        c1.sync()
        res = {"bit_exact": True, "rel_l2": {}, "how": f"outputs of the {world}-rank run gathered on rank 0 and compared "
               f"with one single-GPU run of the same workload on the same {n1}x{n2}x{n3g} input"}
        for nm, comp in wl.outputs:
            a, b = full[nm], ref[nm]
            exact = bool(torch.equal(a, b))
            num = den = 0.0
            step = 1 << 26
            for o in range(0, a.numel(), step):
                x = a[o:o + step].double(); y = b[o:o + step].double()
                num += float(((x - y) ** 2).sum()); den += float((y * y).sum())
            res["rel_l2"][nm] = (num / den) ** 0.5 if den > 0 else (0.0 if num == 0 else float("inf"))
            res["bit_exact"] = res["bit_exact"] and exact
        res["ok"] = all(v <= 1e-5 for v in res["rel_l2"].values())
        c1.close()
        del full, ref
        torch.cuda.empty_cache()
    return res


def _main(real_stdout):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="dip3d_somf3d", choices=sorted(WORKLOADS))
    ap.add_argument("--shape", default=None, help="n1,n2[,n3] (default: the workload's BASELINE.json shape)")
    ap.add_argument("--ref-shape", default=None, help="per-process sample volume of the reference arm")
    ap.add_argument("--cpu-shape", default=None, help="per-process sample volume of cpu_baseline (~10-25 s per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the comparison with a single-GPU run")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return reference_arm(args, wl, rank, world, real_stdout)
    if world > wl.max_gpus:
        raise SystemExit(f"workload {wl.name} runs on at most {wl.max_gpus} GPU(s)")

    dist = torch = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import pyseistr_b200 as ps
    from pyseistr_b200 import _lib
    from pyseistr_b200 import dist as pd

    n1, n2, n3g = parse_shape(args, wl)
    # N > 1: the cube is cut into n3-slabs, one per rank; the library exchanges halos, carry planes
    # of the axis-3 running sums and all-reduces the CG scalars over NCCL.
    if world > 1:
        pd.check_slabs(n3g, world, r3=DIP_KW["rect"][2], ns3=SOMF_KW["r2"])
        ctx = pd.context_from_torch(dist, local)
        z0, z1 = ctx.slab(n3g)
    else:
        ctx = ps.Context(local)
        z0, z1 = 0, n3g
    lib = ctx.lib
    n3 = z1 - z0
    N = n1 * n2 * n3
    Nglobal = n1 * n2 * n3g

    env = Env()
    env.n1, env.n2, env.n3g, env.z0, env.z1, env.n = n1, n2, n3g, z0, z1, N
    env.rank, env.world, env.local, env.ctx, env.lib = rank, world, local, ctx, lib

    def allmax(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    env.allmax = allmax

    # ---- synthetic inputs in pinned host memory (each rank generates its slab of the same volume)
    host, _keep = {}, []
    for nm in wl.inputs:
        host[nm], p = pinned_array(lib, N); _keep.append(p)
    for nm, comp in wl.outputs:
        host[nm], p = pinned_array(lib, comp * N); _keep.append(p)
    wl.make_inputs(env, host)

    # ---- device-resident leg
    dev = {}
    for nm in wl.inputs:
        dev[nm] = ctx.alloc(4 * N)
        ctx.h2d(dev[nm], host[nm])
    for nm, comp in wl.outputs:
        dev[nm] = ctx.alloc(4 * comp * N)

    def step_dev():
        wl.run_dev(lib, ctx.handle, n1, n2, n3g, N, dev)

    def barrier():
        ctx.sync()
        if dist is not None:
            torch.cuda.synchronize()
            dist.barrier()

    for _ in range(args.warmup):
        step_dev()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ctx.set_profile(True)
    ctx.reset_stats()
    ctx.timer_start()
    for _ in range(args.steps):
        step_dev()
    ms = ctx.timer_stop()
    barrier()
    st = ctx.stats()
    ctx.set_profile(False)
    clocks = sampler.stop()
    if os.environ.get("PST_BENCH_ALLRANKS"):          # debugging aid: every rank's per-class device time
        print(f"[rank {rank}] ms {ms:.1f} " + " ".join(f"{k}={v / args.steps:.1f}" for k, v in
              zip(_lib.KERNEL_CLASSES, st["class_ms"]) if v > 0), file=sys.stderr, flush=True)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = Nglobal / (ms_per_step * 1e-3) / 1e6

    # ---- N > 1: is the multi-GPU result the single-GPU result?  (after the timed region, on rank 0's GPU)
    parity = None
    if world > 1 and not args.no_parity:
        parity = parity_vs_n1(env, wl, dev, torch, dist)
        barrier()

    # ---- end-to-end leg: reference-facing C-ABI calls on host buffers
    e2e = None
    if not args.no_e2e:
        def step_e2e():
            wl.run_host(lib, ctx.handle, n1, n2, n3g, N, host)
        for p in dev.values():
            ctx.free(p)
        step_e2e()                                   # warm-up (allocations, page mapping)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        barrier()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        h2d, d2h = wl.copies(N)
        e2e = {"value": Nglobal / (dt / args.steps) / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
               "ms_per_step": dt / args.steps * 1e3, "path": wl.e2e_path}

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel class: CUDA events around every launch of the class
    # during the timed steps (on the library's stream); algorithmic bytes / flops counted by the library
    # per launch
    peak, peak_src = load_peaks()
    names = _lib.KERNEL_CLASSES
    cls_ms = dict(zip(names, st["class_ms"]))
    cls_n = dict(zip(names, st["class_launches"]))
    cls_b = dict(zip(names, st["class_bytes"]))
    cls_f = dict(zip(names, st["class_flops"]))
    # kernels, not classes: the three smoothing classes are ONE kernel (the classes only split its launches by axis),
    # every other class is one kernel family
    groups = {"tri_smooth": ("tri_axis1", "tri_axis2", "tri_axis3", "tri_axis3_bwd")}
    for k in names:
        if k not in groups["tri_smooth"] and k not in ("other", "slot_reduce"):
            groups[k] = (k,)
    g_ms = {g: sum(cls_ms[k] for k in ks) for g, ks in groups.items()}
    g_n = {g: sum(cls_n[k] for k in ks) for g, ks in groups.items()}
    g_b = {g: sum(cls_b[k] for k in ks) for g, ks in groups.items()}
    g_f = {g: sum(cls_f[k] for k in ks) for g, ks in groups.items()}
    live = [g for g in groups if g_n[g] > 0 and (g_b[g] > 0 or g_f[g] > 0)]
    dom = max(live, key=lambda g: g_ms[g])
    total_cls = sum(cls_ms.values()) or 1.0
    tri_kernel = os.environ.get("PST_TRI_KERNEL_NAME", "tri_l2_kernel<CONTIG,NB> (axis 1), tri_sys_kernel<CONTIG,NB,..> (axes 2-3); distributed axis 3: tri3_tile_fwd/bwd_kernel")
    kernel_names = {"tri_smooth": tri_kernel,
                    "cg_head": "cg_head4d_kernel", "cg_dir": "cg_dird_kernel", "cg_gp": "cg_gp4_kernel",
                    "allpass": "allpass_kernel", "cg_setup": "divne_prescale/scale_init_kernel / pwd3_*/cgstep kernels",
                    "predict": "predict_fast_kernel<NW,TWO> / predict_warp_kernel<NW,TWO> / predict_adj_kernel<NW>"}
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    fp32_peak = FP32_LANES * 2 * sm_mhz * 1e6 / 1e12           # TFLOP/s at the SM clock sampled during the run
    if dom == "predict":
        roof = {"bound": "alu", "kernel_class": dom, "kernel": kernel_names[dom], "peak": fp32_peak, "unit": "TFLOP/s",
                "peak_source": f"FP32 non-tensor: 148 SMs x 128 lanes x 2 flop x {sm_mhz:.0f} MHz (SM clock sampled during the run)",
                "achieved": g_f[dom] / (g_ms[dom] * 1e-3) / 1e12, "traffic": None,
                "avg_launch_ms": g_ms[dom] / g_n[dom], "launches": g_n[dom],
                "algorithmic_flops_per_launch": g_f[dom] / g_n[dom],
                "flop_model": "SURVEY 8d: per predicted sample predict1 47 (order 1) / 116 (order 2), predict2 78 / 187 flop",
                "share_of_step": g_ms[dom] / total_cls}
    else:
        roof = {"bound": "hbm", "kernel_class": dom, "kernel": kernel_names.get(dom, dom), "peak": peak, "unit": "GB/s",
                "peak_source": peak_src,
                "achieved": g_b[dom] / (g_ms[dom] * 1e-3) / 1e9,
                "traffic": None,
                "avg_launch_ms": g_ms[dom] / g_n[dom], "launches": g_n[dom],
                "algorithmic_bytes_per_launch": g_b[dom] / g_n[dom],
                "share_of_step": g_ms[dom] / total_cls}
    roof["frac"] = roof["achieved"] / roof["peak"]
    # DRAM traffic per launch: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of that
    # kernel, as a ratio to its algorithmic bytes (profiles/ncu_traffic.json: kernel, capture file, shape, commit)
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if dom in tr and roof["bound"] == "hbm":
            roof["traffic"] = tr[dom]["dram_bytes_per_algorithmic_byte"] * roof["algorithmic_bytes_per_launch"]
            roof["traffic_source"] = {k: tr[dom][k] for k in ("kernel", "capture", "shape", "commit") if k in tr[dom]}
    except Exception:
        pass
    roof["classes"] = {k: {"ms_per_step": cls_ms[k] / args.steps, "launches_per_step": cls_n[k] / args.steps,
                           "algorithmic_GBps": (cls_b[k] / (cls_ms[k] * 1e-3) / 1e9) if cls_ms[k] > 0 and cls_b[k] > 0 else None,
                           "algorithmic_TFLOPs": (cls_f[k] / (cls_ms[k] * 1e-3) / 1e12) if cls_ms[k] > 0 and cls_f[k] > 0 else None,
                           "share": cls_ms[k] / total_cls}
                       for k in names if cls_n[k] > 0}

    line = {
        "metric": wl.metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl.label((n1, n2, n3g)),
                   "parallelism": "single GPU" if world == 1 else
                   f"{world} n3-slabs; axis-3 running sums: carries and halo planes through peer memory over NVLink, "
                   f"inside the kernels; NCCL: halos of the xline stencil and the spray, all-reduce of the per-plane sum "
                   f"tables (partition-independent CG / line-search scalars)",
                   "l2_policy": "inputs (4 B x voxels per volume) are far larger than the 126 MB L2"
                   if 4 * N > 4 * 126e6 else "L2 flushed by the workload itself: every step streams several volumes "
                                              "through scratch buffers whose total exceeds the 126 MB L2",
                   "executed": {"cg_iterations_per_step": st["cg_iterations"] / args.steps,
                                "gn_iterations_per_step": st["gn_iterations"] / args.steps,
                                "linesearch_evals_per_step": st["linesearch_evals"] / args.steps,
                                "smooth_passes_per_step": st["smooth_passes"] / args.steps,
                                "predictions_per_step": st["predictions"] / args.steps}},
        "gpu_launches": int(st["kernel_launches"]),
        "clocks": clocks,
        "roofline": roof,
    }
    if e2e:
        line["e2e"] = e2e
    if world > 1:
        # the class timers above are rank 0's; N = its slab
        roof["note"] = "per-launch times and bytes are rank 0's slab"
        if parity is not None:
            line["parity_vs_n1"] = parity["ok"]
            line["parity_detail"] = parity
    if not args.no_cpu_baseline and world == 1:
        shape = tuple(int(v) for v in args.cpu_shape.split(",")) if args.cpu_shape else wl.cpu_shape
        nproc = os.cpu_count() or 1
        try:
            vals, kind = run_reference_cpu(wl, shape, nproc, 1)
            line["cpu_baseline"] = {"value": vals[0], "unit": UNIT, "cores": nproc, "kind": kind,
                                    "sample": f"{nproc} independent processes x {shape_str(shape)} volume, "
                                              f"{wl.ref_what}, one pass"}
        except Exception as e:      # the checker is optional for the number, never for the product
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": nproc, "kind": "unavailable",
                                    "sample": str(e)[:200]}
    emit(real_stdout, line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
