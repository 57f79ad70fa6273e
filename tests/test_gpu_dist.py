"""GPU suite: n3-slab decomposition over every GPU of the box (up to 8) vs the oracle (skipped on single-GPU boxes;
`bench.py --gpus N` carries the same evidence for the driver's scaling runs as `parity_vs_n1`)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _world():
    from pyseistr_b200 import _lib
    return min(8, _lib.load().pst_device_count())


def _run(world, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0 and "[dist_check] PASS" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])


def test_slabs_over_all_gpus_match_oracle():
    """world = min(8, GPUs): 128-plane slabs (the 8-GPU geometry of the headline cube: tri3_tile_*<128>), tall slabs
    (<64>), n3 % world != 0, masks, soint3d, sint3d, smoothc -- all against the oracle on the whole cube."""
    world = _world()
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    _run(world, 29517)


def test_two_rank_slabs_match_oracle():
    """The 2-rank geometry (tall slabs: line kernels for axis 3) even when the box has more GPUs."""
    world = _world()
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    if world == 2:
        pytest.skip("covered by test_slabs_over_all_gpus_match_oracle")
    _run(2, 29519)
