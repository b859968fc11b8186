"""Worker of tests/test_gpu_dist.py: one process per GPU (or a single process: world 1), NCCL through the library's own
trt_dist_* entry points.  Each rank runs the statSTR / associaTR kernels on ITS block (different sizes per rank),
rank 0 gathers the result regions from DEVICE buffers and checks them against a local recomputation of every rank's
block; plus the sample-counter all-reduces (NaN poison), max, barrier and the ragged byte gather of the CLIs."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trtools_b200 import _lib, dist as tdist, synth  # noqa: E402

SEED = 99
S = 4100


def block_of(ctx, r):
    L = 300 + 37 * r
    sl = synth.make_loci(L, seed=SEED + r)
    ctx.block_begin(L, S, 2, "hipstr")
    ctx.synth_fill(SEED, 1000 * r, sl.cum_freq, sl.miss_thresh, sl.half_thresh, with_format=False)
    ctx.block_set_alleles(*synth.allele_tables(sl))
    ctx.check(ctx.lib.trt_harmonize(ctx.h))
    return L, ctx.nA


def design(ctx):
    rng = np.random.default_rng(SEED)
    traits = rng.standard_normal((S, 4))
    cov = np.hstack([np.full((S, 1), -1.0), traits])
    cov = (cov - cov.mean(axis=0)) / np.maximum(cov.std(axis=0), 1e-300)
    out = cov[:, 1].copy()
    cov[:, 1] = 1.0
    ctx.assoc_set_design(cov, out, np.arange(S, dtype=np.int32))


def main():
    rank, world, local_rank = tdist.env_rank_world()
    ctx = _lib.Context(local_rank)
    comm = tdist.NcclComm(ctx, rank, world)          # also at world 1: exercises the self-gather path
    ok = True
    design(ctx)
    L, nA = block_of(ctx, rank)
    sizes = comm.allgather_i64([L, nA])
    ok = ok and sizes[:, 0].tolist() == [300 + 37 * r for r in range(world)]
    # results stay on the device; rank 0 gathers three regions without blocking, then waits once
    ctx.locus_stats(False, None, 0.01, want=())
    host = None
    if rank == 0:
        host = [ctx.pinned_empty((int(sizes[:, 0].sum()) * 96,), np.uint8), ctx.pinned_empty((int(sizes[:, 1].sum()) * 4,), np.uint8),
                ctx.pinned_empty((int(sizes[:, 0].sum()) * 40,), np.uint8)]
    comm.gather_region(tdist.REGION_STATS, 0, 96 * L, sizes[:, 0] * 96, 0, None if host is None else host[0], wait=False)
    comm.gather_region(tdist.REGION_ALLELE_COUNTS, 0, 4 * nA, sizes[:, 1] * 4, 0, None if host is None else host[1], wait=False)
    ctx.assoc_ols(5.0, want=())
    comm.gather_region(tdist.REGION_ASSOC, 0, 40 * L, sizes[:, 0] * 40, 0, None if host is None else host[2], wait=False)
    comm.wait()
    if rank == 0:
        o96 = o4 = o40 = 0
        for r in range(world):
            Lr, nAr = block_of(ctx, r)
            st = ctx.locus_stats(False, None, 0.01)
            tab = np.frombuffer(host[0][o96:o96 + 96 * Lr].tobytes(), dtype=np.float64).reshape(12, Lr)
            for k, key in enumerate(("thresh", "het", "entropy", "mean", "mode", "var", "hwep")):
                ok = ok and np.array_equal(tab[k], st[key][0], equal_nan=True)
            itab = tab.view(np.int64)
            ok = ok and np.array_equal(itab[8], st["n_hom"][0]) and np.array_equal(itab[9], st["n_called"][0])
            ok = ok and np.array_equal(itab[10], st["n_called_nonstrict"][0]) and np.array_equal(itab[11], st["n_padded"][0])
            ok = ok and np.array_equal(np.frombuffer(host[0][o96 + 56 * Lr:o96 + 56 * Lr + 4 * Lr].tobytes(), dtype=np.int32), st["nalleles"][0])
            ok = ok and np.array_equal(np.frombuffer(host[1][o4:o4 + 4 * nAr].tobytes(), dtype=np.int32), st["ac"][0])
            res = ctx.assoc_ols(5.0)
            atab = np.frombuffer(host[2][o40:o40 + 40 * Lr].tobytes(), dtype=np.float64).reshape(5, Lr)
            for k, key in enumerate(("p", "coef", "se", "r2", "std_g")):
                ok = ok and np.array_equal(atab[k], res[key], equal_nan=True)
            ok = ok and int((res["filter_code"] == 0).sum()) > 10
            o96 += 96 * Lr
            o4 += 4 * nAr
            o40 += 40 * Lr
    # dumpSTR's per-sample accumulators: int64 sums and the NaN-poisoned float64 total depth (dumpSTR.py:710-713)
    counts = np.zeros(50, dtype=np.int64)
    counts[rank::world] = rank + 1
    total = comm.allreduce_sum(counts)
    expect = np.zeros(50, dtype=np.int64)
    for r in range(world):
        expect[r::world] = r + 1
    ok = ok and np.array_equal(total, expect)
    dp = np.zeros(4) + rank
    if rank == world - 1:
        dp[2] = np.nan
    dp_total = comm.allreduce_sum(dp)
    ok = ok and np.isnan(dp_total[2]) and dp_total[0] == sum(range(world))
    ok = ok and comm.max(10.0 + rank) == 10.0 + world - 1
    # the CLIs' block dealing + ragged text gather
    sh = tdist.BlockSharder(comm)
    for b in range(9):
        if sh.mine():
            sh.add(("rank%d-block%d\n" % (rank, b)) * (b + 1))
    merged = sh.finish()
    if rank == 0:
        ok = ok and merged == [(("rank%d-block%d\n" % (b % world, b)) * (b + 1)).encode() for b in range(9)]
    else:
        ok = ok and merged is None
    rows = np.arange(3 * (5 + rank), dtype=np.float64).reshape(-1, 3) + 1000 * rank
    table = comm.gather_table(rows)
    if rank == 0:
        want = np.concatenate([np.arange(3 * (5 + r), dtype=np.float64).reshape(-1, 3) + 1000 * r for r in range(world)])
        ok = ok and np.array_equal(table, want)
    comm.barrier()
    comm.close()
    ctx.close()
    print("RANK{} {}".format(rank, "OK" if ok else "FAIL"), flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
