"""Profiling driver (GPU): one resident synthetic HipSTR block, then statSTR / dumpSTR / associaTR passes.
usage: python tools/prof_tools.py [L] [S] [reps] [tools=stat,dump,assoc]   (wrap in ncu to profile single kernels)"""
import sys, os, time, numpy as np
sys.path.insert(0, '.')
from trtools_b200 import _lib, synth
L = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
tools = (sys.argv[4] if len(sys.argv) > 4 else "stat,dump,assoc").split(",")
SEED = 20261017
ctx = _lib.Context(0)
loci = synth.make_loci(L, seed=SEED)
ctx.block_begin(L, S, 2, "hipstr")
ctx.synth_fill(SEED, 0, loci.cum_freq, loci.miss_thresh, loci.half_thresh, with_format=5)
ctx.block_set_alleles(*synth.allele_tables(loci))
ctx.check(ctx.lib.trt_harmonize(ctx.h))
if "stat" in tools:
    for i in range(reps):
        t0 = time.perf_counter()
        ctx.check(ctx.lib.trt_harmonize(ctx.h))
        t1 = time.perf_counter()
        ctx.locus_stats(False, None, 0.01, pinned=True)
        t2 = time.perf_counter()
        print("statSTR harmonize wall %.3f ms, locus_stats wall %.3f ms (kernels %.3f, scan %.3f)" % (
            (t1 - t0) * 1e3, (t2 - t1) * 1e3, ctx.last_kernel_ms(), ctx.last_scan_ms()), flush=True)
if "pack" in tools:
    for i in range(reps):
        ctx.check(ctx.lib.trt_pack_length_genotypes(ctx.h))
        ctx.synchronize()
        print("pack: kernels %.3f ms -> %.0f GB/s of 10 B/call" % (ctx.last_kernel_ms(), 10.0 * L * S / ctx.last_kernel_ms() / 1e6), flush=True)
if "assoc" in tools:
    rng = np.random.default_rng(SEED)
    traits = np.hstack([rng.standard_normal((S, 1)), rng.standard_normal((S, 10))])
    covars = np.hstack([np.full((S, 1), -1.0), traits])
    covars = (covars - covars.mean(axis=0)) / np.maximum(covars.std(axis=0), 1e-300)
    outcome = covars[:, 1].copy()
    covars[:, 1] = 1.0
    ctx.assoc_set_design(covars, outcome, np.arange(S, dtype=np.int32))
    for i in range(reps):
        t0 = time.perf_counter()
        ctx.assoc_ols(20.0, pinned=True)
        print("associaTR wall %.3f ms (kernels %.3f, sample-axis kernels %.3f)" % (
            (time.perf_counter() - t0) * 1e3, ctx.last_kernel_ms(), ctx.last_scan_ms()), flush=True)
if "dump" in tools:
    cf = [(_lib.CF_RATIO_GT, _lib.FMT_DFLANKINDEL, 0.15), (_lib.CF_MIN, _lib.FMT_DP, 20)]
    counts = np.zeros((2, S), np.int64); numcalls = np.zeros(S, np.int64); totaldp = np.zeros(S)
    for i in range(reps):
        t0 = time.perf_counter()
        ctx.call_filters(cf, _lib.FMT_DP, counts, numcalls, totaldp, want_mask=False, want_trigger=False, want_gt=False)
        t1 = time.perf_counter(); k1 = ctx.last_scan_ms()
        ctx.locus_filters([(_lib.LF_HWE, 1e-4)], False, pinned=True)
        t2 = time.perf_counter()
        print("dumpSTR call_filters wall %.3f ms (kernel %.3f); locus_filters wall %.3f ms (scan %.3f)" % (
            (t1 - t0) * 1e3, k1, (t2 - t1) * 1e3, ctx.last_scan_ms()), flush=True)
