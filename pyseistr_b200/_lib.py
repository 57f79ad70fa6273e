"""ctypes binding of libpst_b200.so (the C-ABI in include/pst_b200.h).

The shared library is built in-tree by ``pyseistr_b200/csrc/Makefile`` (or
``__graft_entry__.build()``).  There is no CPU fallback: if the library is missing, or no
sm_100 GPU is visible, the entry points raise.
"""
import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpst_b200.so")

_fp = ctypes.POINTER(ctypes.c_float)
_vp = ctypes.c_void_p
_i = ctypes.c_int
_f = ctypes.c_float


class PstError(RuntimeError):
    """A C-ABI call returned a non-zero status."""

    def __init__(self, code, msg):
        super().__init__(f"pst_b200 error {code}: {msg}")
        self.code = code


class Stats(ctypes.Structure):
    _fields_ = [("kernel_launches", ctypes.c_longlong), ("cg_iterations", ctypes.c_longlong),
                ("linesearch_evals", ctypes.c_longlong), ("gn_iterations", ctypes.c_longlong),
                ("smooth_passes", ctypes.c_longlong), ("predictions", ctypes.c_longlong),
                ("device_ms", ctypes.c_double), ("h2d_bytes", ctypes.c_double),
                ("d2h_bytes", ctypes.c_double), ("class_ms", ctypes.c_double * 12),
                ("class_launches", ctypes.c_longlong * 12), ("class_bytes", ctypes.c_double * 12),
                ("class_flops", ctypes.c_double * 12)]

    def as_dict(self):
        d = {}
        for k, _ in self._fields_:
            v = getattr(self, k)
            d[k] = list(v) if hasattr(v, "__len__") else v
        return d


KERNEL_CLASSES = ("allpass", "tri_axis1", "tri_axis2", "tri_axis3", "cg_setup", "predict", "slot_reduce", "other",
                  "cg_head", "cg_gp", "cg_dir", "tri_axis3_bwd")


# name -> (restype, argtypes); every symbol include/pst_b200.h declares
SIGNATURES = {
    "pst_last_error": (ctypes.c_char_p, []),
    "pst_version": (ctypes.c_char_p, []),
    "pst_device_count": (_i, []),
    "pst_ctx_create": (_i, [_i, ctypes.POINTER(_vp)]),
    "pst_ctx_destroy": (None, [_vp]),
    "pst_ctx_stats": (_i, [_vp, ctypes.POINTER(Stats)]),
    "pst_ctx_reset_stats": (_i, [_vp]),
    "pst_ctx_set_profile": (_i, [_vp, _i]),
    "pst_timer_start": (_i, [_vp]),
    "pst_timer_stop": (_i, [_vp, ctypes.POINTER(ctypes.c_double)]),
    "pst_comm_unique_id": (_i, [_vp]),
    "pst_ctx_create_dist": (_i, [_i, _i, _i, _vp, ctypes.POINTER(_vp)]),
    "pst_ctx_slab": (_i, [_vp, _i, ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    "pst_dip": (_i, [_vp, _fp, _fp, _i, _i, _i, _i, _i, _i, _f, _f, _f, _i, _i, _i, _i, _fp]),
    "pst_dip_dev": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "pst_somean3d": (_i, [_vp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _f, _i, _fp]),
    "pst_somf3d": (_i, [_vp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _i, _i, _f, _i, _fp]),
    "pst_somean3d_dev": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "pst_somf3d_dev": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "pst_somean2d": (_i, [_vp, _fp, _fp, _i, _i, _i, _i, _i, _i, _f, _i, _fp]),
    "pst_somf2d": (_i, [_vp, _fp, _fp, _i, _i, _i, _i, _i, _i, _i, _f, _i, _fp]),
    "pst_somean2d_dev": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp]),
    "pst_somf2d_dev": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _vp]),
    "pst_soint3d": (_i, [_vp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _i, _fp]),
    "pst_soint3d_dev": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _i, _vp]),
    "pst_sint3d": (_i, [_vp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _fp]),
    "pst_sint3d_dev": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _vp]),
    "pst_paint2d": (_i, [_vp, _fp, _fp, _i, _i, _i, _i, _f, _i, _fp]),
    "pst_paint2d_dev": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _vp]),
    "pst_allpass_dev": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "pst_smooth3_dev": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i]),
    "pst_selftest_axis3_slabs": (_i, [_vp, _vp, _i, _i, _i, _i, _i]),
    "pst_divne_dev": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, ctypes.POINTER(_i)]),
    "pst_smooth3": (_i, [_vp, _fp, _i, _i, _i, _i, _i, _i, _i, _i, _fp]),
    "pst_smoothcf": (_i, [_vp, _fp] + [_i] * 14 + [_fp]),
    "pst_smoothcf_dev": (_i, [_vp, _vp] + [_i] * 14),
    "pst_dev_alloc": (_i, [_vp, ctypes.c_size_t, ctypes.POINTER(_vp)]),
    "pst_dev_free": (_i, [_vp, _vp]),
    "pst_h2d": (_i, [_vp, _vp, _vp, ctypes.c_size_t]),
    "pst_d2h": (_i, [_vp, _vp, _vp, ctypes.c_size_t]),
    "pst_host_alloc_pinned": (_i, [ctypes.c_size_t, ctypes.POINTER(_vp)]),
    "pst_host_free_pinned": (_i, [_vp]),
    "pst_sync": (_i, [_vp]),
}

_lib = None
_lock = threading.Lock()


def build(verbose=False):
    """Compile libpst_b200.so in-tree (nvcc, sm_100a)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j4"]
    if not verbose:
        cmd.insert(1, "-s")
    subprocess.check_call(cmd)


def load():
    """dlopen the library and set the ctypes prototypes.  Raises if it is not built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(
                    f"{LIB_PATH} is missing: build it with `make -C pyseistr_b200/csrc` "
                    "(there is no CPU fallback)")
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        msg = load().pst_last_error()
        raise PstError(rc, msg.decode() if msg else "")


class Context:
    """One GPU context (stream, workspace arena, reduction buffers)."""

    def __init__(self, device=0, rank=0, nranks=1, nccl_id=None):
        """Single-GPU context, or (nranks > 1) one rank of an n3-slab decomposition; nccl_id is the
        128-byte id from ``unique_id()`` on rank 0, distributed by the caller."""
        self._h = _vp()
        self.lib = load()
        self.rank, self.nranks = int(rank), int(nranks)
        if self.nranks > 1:
            buf = ctypes.create_string_buffer(bytes(nccl_id), 128)
            check(self.lib.pst_ctx_create_dist(int(device), self.rank, self.nranks, buf, ctypes.byref(self._h)))
        else:
            check(self.lib.pst_ctx_create(int(device), ctypes.byref(self._h)))
        self.device = device

    def slab(self, n3):
        """Global plane range [z0, z1) this rank owns."""
        z0, z1 = _i(0), _i(0)
        check(self.lib.pst_ctx_slab(self._h, int(n3), ctypes.byref(z0), ctypes.byref(z1)))
        return z0.value, z1.value

    @property
    def handle(self):
        return self._h

    def stats(self):
        s = Stats()
        check(self.lib.pst_ctx_stats(self._h, ctypes.byref(s)))
        return s.as_dict()

    def reset_stats(self):
        check(self.lib.pst_ctx_reset_stats(self._h))

    def sync(self):
        check(self.lib.pst_sync(self._h))

    def set_profile(self, on):
        check(self.lib.pst_ctx_set_profile(self._h, int(bool(on))))

    def timer_start(self):
        check(self.lib.pst_timer_start(self._h))

    def timer_stop(self):
        ms = ctypes.c_double(0.0)
        check(self.lib.pst_timer_stop(self._h, ctypes.byref(ms)))
        return ms.value

    def alloc(self, nbytes):
        p = _vp()
        check(self.lib.pst_dev_alloc(self._h, int(nbytes), ctypes.byref(p)))
        return p

    def free(self, p):
        check(self.lib.pst_dev_free(self._h, p))

    def h2d(self, dptr, arr):
        check(self.lib.pst_h2d(self._h, dptr, arr.ctypes.data_as(_vp), arr.nbytes))

    def d2h(self, arr, dptr):
        check(self.lib.pst_d2h(self._h, arr.ctypes.data_as(_vp), dptr, arr.nbytes))

    def close(self):
        if self._h:
            self.lib.pst_ctx_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def unique_id():
    """128-byte NCCL id (call on rank 0, broadcast to the other ranks)."""
    buf = ctypes.create_string_buffer(128)
    check(load().pst_comm_unique_id(buf))
    return buf.raw


_default_ctx = {}


def default_context(device=None):
    """Process-wide context per device (PST_DEVICE or LOCAL_RANK selects the default GPU)."""
    if device is None:
        device = int(os.environ.get("PST_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]
