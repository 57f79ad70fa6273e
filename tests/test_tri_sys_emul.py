"""CPU suite: the systolic smoother kernel (pyseistr_b200/csrc/pst_tri_sys_kernels.cuh) run on the host, thread for
thread (tests/native/tri_sys_emul.cpp: mbarriers with the hardware's phase-parity semantics, TMA boxes with zero fill,
one std::thread per CUDA thread), against the oracle.  Several tiles per CTA so that both orientations, the mailbox
hand-offs and the loader's ring are exercised; one case runs with schedule fuzzing."""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("tri_sys_emul") / "tri_sys_emul.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-pthread", "-I",
                    os.path.join(ROOT, "pyseistr_b200", "csrc"), "-o", so,
                    os.path.join(ROOT, "tests", "native", "tri_sys_emul.cpp")], check=True)
    lib = ctypes.CDLL(so)
    lib.tri_sys_plan.restype = ctypes.c_int
    lib.tri_sys_plan.argtypes = [ctypes.c_int] * 5 + [ctypes.c_void_p]
    lib.path = so
    return lib


def run_cases(emul, cases, env=None):
    """The emulated kernels run in a child process: a protocol deadlock aborts the child, not pytest."""
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "native", "run_emul.py"), emul.path, "tri_sys_emul",
                        json.dumps(cases)], env=e, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


CASES = [  # shape, axis, radius, emulated SM count
    ((8, 150, 3), 1, 5, 2),
    ((36, 3, 530), 2, 10, 1),        # 132-step segments, 110 leading dummy steps, 4 tiles on one CTA
    ((100, 140, 7), 1, 8, 3),        # 28 tiles on 6 CTAs, partial tiles (100 = 3 * 32 + 4 lanes)
    ((12, 9, 260), 2, 2, 1),
    ((64, 1034, 2), 1, 5, 1),        # 8 segments of 132 (the 1024-sample case of the bench)
    ((40, 300, 2), 1, 3, 1),
    # contiguous axis: x box = 32 lines x samples, stores transposed through the warp's tile
    ((152, 9, 5), 0, 5, 1),          # 45 lines: a full and a 13-line tile
    ((1032, 5, 9), 0, 5, 1),         # 8 segments of 132
    ((528, 40, 3), 0, 10, 2),
    ((260, 70, 1), 0, 2, 1),
]


def test_systolic_kernel_emulated_on_host_matches_oracle(emul):
    plan = (ctypes.c_int * 4)()
    cases = []
    for shape, axis, nb, sm in CASES:
        assert emul.tri_sys_plan(*shape, axis, nb, plan) == 1, (shape, axis, nb)
        cases += [[list(shape), axis, nb, sm, 0], [list(shape), axis, nb, sm, 1]]
    run_cases(emul, cases)


def test_systolic_kernel_under_schedule_fuzzing(emul):
    run_cases(emul, [[[100, 140, 5], 1, 5, 1, 0], [[100, 140, 5], 1, 5, 1, 1], [[152, 9, 5], 0, 5, 1, 0]], {"PST_EMUL_JITTER": "300"})


def test_systolic_kernel_stores_inside_the_chain_variant(emul):
    """PST_TRI_SYS_ILS=1: last and interior segments store from inside the backward chain loop (strided axes)."""
    cases = []
    for shape, axis, nb, sm in (((100, 140, 7), 1, 8, 3), ((36, 3, 530), 2, 10, 1), ((64, 1034, 2), 1, 5, 1)):
        cases += [[list(shape), axis, nb, sm, 0], [list(shape), axis, nb, sm, 1]]
    run_cases(emul, cases, {"PST_TRI_SYS_ILS": "1"})


def test_systolic_kernel_prebuild_variant(emul):
    """PST_TRI_SYS_PRE=1: t of the next tile built in place in the x box while waiting for a carry; several tiles per
    CTA so that the hand-over between tiles is exercised, once with schedule fuzzing, once combined with ILS."""
    cases = []
    for shape, axis, nb, sm in (((100, 140, 7), 1, 8, 2), ((36, 3, 530), 2, 10, 1), ((64, 1034, 2), 1, 5, 1),
                                ((152, 9, 25), 0, 5, 1), ((528, 40, 3), 0, 10, 1)):
        cases += [[list(shape), axis, nb, sm, 0], [list(shape), axis, nb, sm, 1]]
    run_cases(emul, cases, {"PST_TRI_SYS_PRE": "1"})
    run_cases(emul, cases[:4], {"PST_TRI_SYS_PRE": "1", "PST_EMUL_JITTER": "200"})
    run_cases(emul, cases[:6], {"PST_TRI_SYS_PRE": "1", "PST_TRI_SYS_ILS": "1"})


def test_plan_refuses_what_the_kernel_cannot_do(emul):
    plan = (ctypes.c_int * 4)()
    assert emul.tri_sys_plan(62, 64, 64, 0, 5, plan) == 0          # n1 % 4 != 0 (TMA box start / stride)
    assert emul.tri_sys_plan(1000, 1024, 1024, 0, 5, plan) == 1 and list(plan)[:3] == [132, 8, 46]
    assert emul.tri_sys_plan(64, 40, 8, 1, 5, plan) == 0           # one segment only
    assert emul.tri_sys_plan(62, 300, 8, 1, 5, plan) == 0          # n1 % 4 != 0 (TMA stride)
    assert emul.tri_sys_plan(64, 2000, 8, 1, 5, plan) == 0         # line longer than 8 segments
    assert emul.tri_sys_plan(1000, 1024, 1024, 1, 5, plan) == 1 and list(plan)[:3] == [132, 8, 22]
    assert emul.tri_sys_plan(1000, 1024, 1024, 2, 5, plan) == 1
