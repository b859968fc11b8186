"""Diagnostic (GPU): run the statSTR pass at several block sizes; on an out-of-range report, fetch the flagged
loci's GT rows and check on the host whether the DATA is out of range or the scan miscounted."""
import os, re, sys, io, contextlib
import numpy as np
sys.path.insert(0, '.')
os.environ["TRT_DEBUG_BAD"] = "1"
from trtools_b200 import _lib, synth
SEED = 20261017
ctx = _lib.Context(int(os.environ.get("DIAG_DEV", "0")))
print(ctx.device_info(), flush=True)
for L in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "20000,100000").split(",")]:
    S = 50000
    loci = synth.make_loci(L, seed=SEED)
    ctx.block_begin(L, S, 2, "hipstr")
    ctx.synth_fill(SEED, 0, loci.cum_freq, loci.miss_thresh, loci.half_thresh, with_format=False)
    ctx.block_set_alleles(*synth.allele_tables(loci))
    ctx.check(ctx.lib.trt_harmonize(ctx.h))
    for rep in range(3):
        try:
            st = ctx.locus_stats(False, None, 0.01)
            print("L=%d rep %d ok  (sum n_called %d, scan %.3f ms)" % (L, rep, int(st["n_called"].sum()), ctx.last_scan_ms()), flush=True)
        except Exception as e:
            print("L=%d rep %d FAILED: %s" % (L, rep, e), flush=True)
    # host-side truth for a sample of loci: any GT entry outside [-2, A)?
    nbad_true = 0
    for l0 in range(0, L, max(1, L // 64)):
        gt = ctx.block_get_gt(l0, 1)[0]
        A = int(loci.n_alleles[l0])
        nbad_true += int(((gt[:, :2] < -2) | (gt[:, :2] >= A)).sum())
    print("L=%d host check of 64 sampled loci: %d out-of-range entries" % (L, nbad_true), flush=True)
