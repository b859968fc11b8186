"""CPU tests of the multi-GPU host logic: locus sharding and the gather / reduce of per-locus tables with
world_size 2 over gloo (the GPU path uses the same code over NCCL)."""
import os
import subprocess
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_locus_shard_partitions_contiguously():
    from trtools_b200.dist import locus_shard
    for L in (0, 1, 7, 8, 100000, 100003):
        for world in (1, 2, 3, 4, 8):
            ranges = [locus_shard(L, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == L
            for a, b in zip(ranges, ranges[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1


def test_gather_and_reduce_world2_gloo():
    port = 29500 + (os.getpid() % 500)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(REPO, "tests", "dist_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "RANK0 OK" in res.stdout and "RANK1 OK" in res.stdout


def test_single_process_passthrough():
    from trtools_b200 import dist as tdist
    rows = np.arange(12, dtype=float).reshape(4, 3)
    assert np.array_equal(tdist.gather_table(None, rows), rows)
    assert tdist.max_over_ranks(None, 3.5) == 3.5
