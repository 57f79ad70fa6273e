"""GPU suite: 2-rank slab decomposition vs the oracle (skipped on single-GPU boxes)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_slabs_match_oracle():
    from pyseistr_b200 import _lib
    if _lib.load().pst_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "[dist_check] PASS" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])
