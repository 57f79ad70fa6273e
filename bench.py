#!/usr/bin/env python
"""bench.py — dip3d + somf3d throughput (Mvoxels/s) on B200, next to the reference C path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--shape n1,n2,n3] [--impl ours|reference]

One "step" = one pass of the hot path over one synthetic cube: dip3dc(defaults: niter 5,
liter 10, order 2, rect 5,5,5) followed by somf3dc(r1=r2=2, order 2, option 1) using the
dips just estimated — the BASELINE.json metric, on its headline configuration
(configs[4], 1000x1024x1024 float32), which fits one B200.

  value     whole-job Mvoxels/s with the cube resident in HBM (CUDA events on the library's
            stream around exactly K steps, max over ranks);
  e2e       the same through the reference-facing C-ABI calls with HOST buffers (pst_dip then
            pst_somf3d: pinned host -> device copies of the inputs and device -> host copies
            of the results inside the timed region);
  roofline  the dominant kernel class, timed live with CUDA events around each of its
            launches during the timed steps, against MEASURED_PEAKS.json;
  cpu_baseline  the unmodified reference C (oracle/_ref) on the box's host cores on a bounded
            sample of the same workload.

`--impl reference` times the reference's own CPU implementation (oracle/_ref, all host
cores as independent processes — the reference is single-threaded and non-re-entrant).
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

METRIC = "Mvoxels/s dip3d+somf3d"
UNIT = "Mvoxels/s"
DEFAULT_SHAPE = (1000, 1024, 1024)
DIP_KW = dict(niter=5, liter=10, order=2, rect=(5, 5, 5))
SOMF_KW = dict(r1=2, r2=2, order=2, option=1)


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
d = synth.erratic(synth.cube(n1, n2, n3, seed=seed), ntraces=max(4, n2 * n3 // 200))
times = []
for s in range(steps):
    t = time.perf_counter()
    di, dx = impl.dip3dc(d, 5, 10, 2, 0.01, 1, 1e-6, (5, 5, 5), 0)
    f = impl.somf3dc(d, di, dx, 2, 2, 0.01, 2, 1, 0)
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


def run_reference_cpu(shape, nproc, steps):
    """nproc independent single-thread processes, each dip3dc+somf3dc on its own cube of `shape`.
    Returns (per-step aggregate Mvox/s list, kind)."""
    kind = reference_kind()
    n1, n2, n3 = shape
    procs = []
    for r in range(nproc):
        code = REF_WORKER.format(root=ROOT, kind=kind, n1=n1, n2=n2, n3=n3, seed=100 + r, steps=steps)
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


def reference_arm(args, rank, world, real_stdout):
    if rank != 0:
        return 0
    nproc = os.cpu_count() or 1
    shape = tuple(int(v) for v in args.ref_shape.split(","))
    steps = args.warmup + args.steps
    t0 = time.perf_counter()
    vals, kind = run_reference_cpu(shape, nproc, steps)
    timed = vals[args.warmup:]
    value = sum(timed) / len(timed)
    vox = float(np.prod(shape)) * nproc
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": vox / value / 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(parse_shape(args)),
                   "reference_sample": f"{nproc} independent processes x ({shape[0]}x{shape[1]}x{shape[2]}) cube "
                                       f"each, dip3dc(defaults)+somf3dc(2,2,order 2) per step",
                   "why_sample": "the reference needs ~4.7 h and >105 GB for the full cube (SURVEY §6)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nproc, "kind": kind,
                         "sample": f"{nproc} x {shape[0]}x{shape[1]}x{shape[2]}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    emit(real_stdout, line)
    return 0


# --------------------------------------------------------------------------- our arm
def parse_shape(args):
    s = args.shape or os.environ.get("PST_BENCH_SHAPE")
    if s:
        return tuple(int(v) for v in s.split(","))
    return DEFAULT_SHAPE


def workload_name(shape):
    return f"dip3dc(niter5,liter10,order2,rect5x5x5)+somf3dc(r2x2,order2,MF) on {shape[0]}x{shape[1]}x{shape[2]} f32"


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


def _main(real_stdout):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default=None, help="n1,n2,n3 (default 1000,1024,1024)")
    ap.add_argument("--ref-shape", default="128,64,32", help="per-process sample cube of the reference arm")
    ap.add_argument("--cpu-shape", default="160,128,64", help="per-process sample cube of cpu_baseline (~10-25 s per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return reference_arm(args, rank, world, real_stdout)

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import pyseistr_b200 as ps
    from pyseistr_b200 import _lib, synth
    from pyseistr_b200 import dist as pd

    n1, n2, n3g = parse_shape(args)
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

    # ---- synthetic input in pinned host memory (each rank generates its slab of the same cube)
    h_in, _p1 = pinned_array(lib, N)
    cube = h_in.reshape((n1, n2, n3), order="F")
    _, mx = synth.cube_big(n1, n2, n3g, seed=7, out=cube, z0=z0, z1=z1, normalise=False)
    if dist is not None:
        import torch
        t = torch.tensor([mx], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        mx = float(t.item())
    synth.scale_by(cube, mx)
    h_dip, _p2 = pinned_array(lib, 2 * N)
    h_out, _p3 = pinned_array(lib, N)

    fp = ctypes.POINTER(ctypes.c_float)
    P = lambda a, off=0: ctypes.cast(a.ctypes.data + 4 * off, fp)

    # ---- device-resident leg
    d_in = ctx.alloc(4 * N)
    d_dip = ctx.alloc(8 * N)
    d_out = ctx.alloc(4 * N)
    ctx.h2d(d_in, h_in)
    d_dipx = ctypes.c_void_p(d_dip.value + 4 * N)
    rmf = 2 * SOMF_KW["r1"] * SOMF_KW["r2"] + 1

    def step_dev():
        _lib.check(lib.pst_dip_dev(ctx.handle, d_in, None, n1, n2, n3g, DIP_KW["niter"], DIP_KW["liter"],
                                   DIP_KW["order"], *DIP_KW["rect"], 0, d_dip))
        _lib.check(lib.pst_somf3d_dev(ctx.handle, d_in, d_dip, d_dipx, n1, n2, n3g, SOMF_KW["r1"], SOMF_KW["r2"],
                                      rmf, SOMF_KW["option"], SOMF_KW["order"], d_out))

    def barrier():
        ctx.sync()
        if dist is not None:
            import torch
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
        import torch
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = Nglobal / (ms_per_step * 1e-3) / 1e6

    # ---- end-to-end leg: reference-facing C-ABI calls on host buffers
    e2e = None
    if not args.no_e2e:
        def step_e2e():
            _lib.check(lib.pst_dip(ctx.handle, P(h_in), None, n1, n2, n3g, DIP_KW["niter"], DIP_KW["liter"],
                                   DIP_KW["order"], 0.01, 1.0, 1e-6, *DIP_KW["rect"], 0, P(h_dip)))
            _lib.check(lib.pst_somf3d(ctx.handle, P(h_in), P(h_dip), P(h_dip, N), n1, n2, n3g, SOMF_KW["r1"],
                                      SOMF_KW["r2"], rmf, SOMF_KW["option"], SOMF_KW["order"], 0.01, 0, P(h_out)))
        ctx.free(d_out); ctx.free(d_dip); ctx.free(d_in)
        step_e2e()                                   # warm-up (allocations, page mapping)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        barrier()
        dt = time.perf_counter() - t0
        if dist is not None:
            import torch
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": Nglobal / (dt / args.steps) / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": 4 * N * 4 * world, "d2h_bytes_per_step": 4 * N * 3 * world,
               "ms_per_step": dt / args.steps * 1e3,
               "path": "pst_dip(host)->pst_somf3d(host): H2D din; D2H dipi,dipx; H2D din,dipi,dipx; D2H out"}

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel class: CUDA events around every launch of the class
    # during the timed steps (on the library's stream); algorithmic bytes counted by the library
    # per launch
    peak, peak_src = load_peaks()
    names = _lib.KERNEL_CLASSES
    cls_ms = dict(zip(names, st["class_ms"]))
    cls_n = dict(zip(names, st["class_launches"]))
    cls_b = dict(zip(names, st["class_bytes"]))
    # kernels, not classes: the three smoothing classes are ONE kernel (tri_stream_kernel, pst_tri_stream.cu;
    # the classes only split its launches by axis), every other class is one kernel family
    groups = {"tri_smooth": ("tri_axis1", "tri_axis2", "tri_axis3")}
    for k in names:
        if k not in groups["tri_smooth"] and k not in ("predict", "other", "slot_reduce", "reserved"):
            groups[k] = (k,)
    g_ms = {g: sum(cls_ms[k] for k in ks) for g, ks in groups.items()}
    g_n = {g: sum(cls_n[k] for k in ks) for g, ks in groups.items()}
    g_b = {g: sum(cls_b[k] for k in ks) for g, ks in groups.items()}
    hbm = [g for g in groups if g_n[g] > 0 and g_b[g] > 0]
    dom = max(hbm, key=lambda g: g_ms[g])
    total_cls = sum(cls_ms.values()) or 1.0
    kernel_names = {"tri_smooth": "tri_stream_kernel<CONTIG,NB> (axes 1-3; distributed axis 3: tri3_tile_fwd/bwd_kernel)",
                    "cg_head": "cg_head4_kernel", "cg_dir": "cg_dir4_kernel", "cg_gp": "cg_gp4_kernel",
                    "allpass": "allpass_kernel", "cg_setup": "divne_prescale/scale_init_kernel"}
    roof = {"bound": "hbm", "kernel_class": dom, "kernel": kernel_names.get(dom, dom), "peak": peak, "unit": "GB/s",
            "peak_source": peak_src,
            "achieved": g_b[dom] / (g_ms[dom] * 1e-3) / 1e9,
            "traffic": None,
            "avg_launch_ms": g_ms[dom] / g_n[dom], "launches": g_n[dom],
            "algorithmic_bytes_per_launch": g_b[dom] / g_n[dom],
            "share_of_step": g_ms[dom] / total_cls}
    roof["frac"] = roof["achieved"] / peak
    # DRAM traffic per launch: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full`
    # capture of that kernel (profiles/r01b_ncu_summary.md, 500x512x512), as a ratio to its algorithmic bytes
    ncu_ratio = {"cg_head": (3.670041 + 2.054506) / (44 * 0.131072), "cg_dir": (3.145764 + 1.531844) / (36 * 0.131072),
                 "cg_gp": (1.048588 + 0.494331) / (12 * 0.131072),
                 "tri_smooth": (0.524381 + 0.470106) / (8 * 0.131072)}
    if dom in ncu_ratio:
        roof["traffic"] = ncu_ratio[dom] * roof["algorithmic_bytes_per_launch"]
        roof["traffic_source"] = "ncu --set full capture scaled by voxel count, see profiles/r01b_ncu_summary.md"
    roof["classes"] = {k: {"ms_per_step": cls_ms[k] / args.steps, "launches_per_step": cls_n[k] / args.steps,
                           "algorithmic_GBps": (cls_b[k] / (cls_ms[k] * 1e-3) / 1e9) if cls_ms[k] > 0 and cls_b[k] > 0 else None,
                           "share": cls_ms[k] / total_cls}
                       for k in names if cls_n[k] > 0}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name((n1, n2, n3g)),
                   "parallelism": "single GPU" if world == 1 else
                   f"{world} n3-slabs, NCCL: halos (xline stencil, spray), carry planes (axis-3 running sums), "
                   f"all-reduced CG scalars",
                   "l2_policy": "inputs (4 B x voxels per volume) are far larger than the 126 MB L2",
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
    if not args.no_cpu_baseline and world == 1:
        shape = tuple(int(v) for v in args.cpu_shape.split(","))
        nproc = os.cpu_count() or 1
        try:
            vals, kind = run_reference_cpu(shape, nproc, 1)
            line["cpu_baseline"] = {"value": vals[0], "unit": UNIT, "cores": nproc, "kind": kind,
                                    "sample": f"{nproc} independent processes x {shape[0]}x{shape[1]}x{shape[2]} cube, "
                                              f"dip3dc(defaults)+somf3dc(2,2,order 2), one pass"}
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
