"""Worker of tests/test_dist_cpu.py (launched by torch.distributed.run with the gloo backend, 2 ranks, CPU)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trtools_b200 import dist as tdist  # noqa: E402


def main():
    rank, world, _ = tdist.env_rank_world()
    d = tdist.init(backend="gloo")
    L = 1001
    lo, hi = tdist.locus_shard(L, rank, world)
    # every rank "computes" the rows of its own loci: row i = [i, i^2, nan for i % 7 == 0]
    idx = np.arange(lo, hi, dtype=np.float64)
    rows = np.stack([idx, idx * idx, np.where(idx % 7 == 0, np.nan, idx / 3.0)], axis=1)
    table = tdist.gather_table(d, rows)
    counts = np.zeros(50, dtype=np.int64)
    counts[rank::world] = rank + 1
    total = tdist.allreduce_sum(d, counts)
    dp = np.zeros(4)
    if rank == 1:
        dp[2] = np.nan
    dp_total = tdist.allreduce_sum(d, dp + rank)
    t = tdist.max_over_ranks(d, 10.0 + rank)
    ok = True
    # the CLIs' block dealing: every rank walks all 7 blocks, keeps the text of its own, rank 0 restores file order
    sh = tdist.BlockSharder(d)
    for b in range(7):
        if sh.mine():
            sh.add(("rank%d-block%d\n" % (rank, b)) * (b + 1))
    merged = sh.finish()
    if rank == 0:
        ok = ok and merged == [(("rank%d-block%d\n" % (b % world, b)) * (b + 1)).encode() for b in range(7)]
    else:
        ok = ok and merged is None
    sizes = d.allgather_i64([rank, 10 * rank + 1])
    ok = ok and sizes.shape == (world, 2) and sizes[:, 1].tolist() == [10 * r + 1 for r in range(world)]
    if rank == 0:
        want = np.arange(L, dtype=np.float64)
        ok = ok and table.shape == (L, 3) and np.array_equal(table[:, 0], want) and np.array_equal(table[:, 1], want * want)
        ok = ok and np.array_equal(np.isnan(table[:, 2]), want % 7 == 0)
    else:
        ok = ok and table is None
    expect = np.zeros(50, dtype=np.int64)
    for r in range(world):
        expect[r::world] = r + 1
    ok = ok and np.array_equal(total, expect)
    ok = ok and np.isnan(dp_total[2]) and dp_total[0] == sum(range(world)) and t == 10.0 + world - 1
    d.close()
    print("RANK{} {}".format(rank, "OK" if ok else "FAIL"), flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
