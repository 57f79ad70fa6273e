"""CPU suite: the host side of the N > 1 path, world_size 2 over gloo (no GPU): slab rule,
slab validation, and the communicator-id hand-off through torch.distributed."""
import os
import socket
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from pyseistr_b200 import dist as pd  # noqa: E402


def test_slab_rule_partitions_exactly():
    for n3 in (7, 10, 64, 1024, 1000):
        for world in (1, 2, 3, 4, 8):
            if world > n3:
                continue
            b = [pd.slab_bounds(n3, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n3
            for (a0, a1), (c0, c1) in zip(b[:-1], b[1:]):
                assert a1 == c0 and a1 > a0
            sizes = [z1 - z0 for z0, z1 in b]
            assert max(sizes) - min(sizes) <= 1


def test_slab_validation():
    assert pd.check_slabs(1024, 8, r3=5, ns3=2) == 128
    with pytest.raises(ValueError):
        pd.check_slabs(64, 8, r3=5, ns3=2)       # 8 planes < 2*r3
    with pytest.raises(ValueError):
        pd.check_slabs(8, 8, r3=1, ns3=2)        # 1 plane < ns3


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r})
    import torch.distributed as dist
    from pyseistr_b200 import dist as pd
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    # the id is made on rank 0 only (a fake one here: no GPU / NCCL on this box) and must reach
    # every rank unchanged
    ident = pd.broadcast_id(dist, make_id=lambda: bytes(range(128)))
    assert ident == bytes(range(128)), "id corrupted"
    z = pd.slab_bounds(1000, rank, world)
    allz = [None] * world
    dist.all_gather_object(allz, z)
    assert allz[0][0] == 0 and allz[-1][1] == 1000 and all(a[1] == b[0] for a, b in zip(allz[:-1], allz[1:]))
    dist.barrier()
    dist.destroy_process_group()
    print("OK", rank)
""")


def test_id_broadcast_and_slabs_over_gloo_world2(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "w.py"
    script.write_text(WORKER.format(root=ROOT))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    for p in procs:
        out, err = p.communicate(timeout=240)
        assert p.returncode == 0, err[-2000:]
        assert "OK" in out
