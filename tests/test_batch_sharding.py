"""Frame sharding across ranks: partition properties on one process, and the gather path with a real
world_size-2 gloo group on CPU (the compute is stubbed -- the GPU call itself is covered by the -m gpu tests)."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from omni_b200 import batch


@pytest.mark.parametrize("n,world", [(0, 1), (1, 4), (7, 2), (8, 8), (512, 8), (513, 8), (5, 3)])
def test_shard_range_partitions(n, world):
    seen = []
    for r in range(world):
        rg = batch.shard_range(n, world, r)
        assert len(rg) <= -(-n // world) if n else len(rg) == 0
        seen += list(rg)
    assert seen == list(range(n))                 # contiguous, ordered, no overlap, complete




def test_gather_single_process():
    local = batch.run_shard([np.full((2, 2), i) for i in range(3)], range(3), lambda f: np.full((4, 3), int(f[0, 0])))
    out = batch.gather_counts(local, 3, 4)
    assert out.shape == (3, 4, 3) and [int(out[i, 0, 0]) for i in range(3)] == [0, 1, 2]
    with pytest.raises(RuntimeError):
        batch.gather_counts({0: np.zeros((4, 3))}, 2, 4)


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, os.environ["OMNI_PKG"])
    from omni_b200 import batch
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n, K = 7, 3
    ids = batch.shard_range(n, world, rank)
    frames = [np.full((4, 5, 3), 10 * i, np.uint8) for i in ids]
    local = batch.run_shard(frames, ids, lambda f: np.full((K, 3), int(f[0, 0, 0]) + rank * 1000))
    out = batch.gather_counts(local, n, K, dist)
    if rank == 0:
        want = [10 * i + (0 if i < 4 else 1000) for i in range(n)]
        assert [int(out[i, 0, 0]) for i in range(n)] == want, out[:, 0, 0]
        print("GATHER_OK")
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()
""")


def test_gather_world2_gloo(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "omnirevolve-image-processor_b200")
    env = dict(os.environ, OMNI_PKG=pkg)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=240)
    assert r.returncode == 0 and "GATHER_OK" in r.stdout, r.stdout[-2000:]
