"""Biobank-shape sanity check (GPU): one locus block of BASELINE config 5's sample count (S = 500 000).
statSTR pass twice (bit-identical), the allele-count checksum, a numpy recount of sampled rows, and the associaTR
fast path against the generic kernels."""
import os, sys, time
import numpy as np
sys.path.insert(0, '.')
from trtools_b200 import _lib, synth
L = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
S = int(sys.argv[2]) if len(sys.argv) > 2 else 500000
SEED = 7
ctx = _lib.Context(0)
loci = synth.make_loci(L, seed=SEED)
ctx.block_begin(L, S, 2, "hipstr")
ctx.synth_fill(SEED, 0, loci.cum_freq, loci.miss_thresh, loci.half_thresh, with_format=False)
ctx.block_set_alleles(*synth.allele_tables(loci))
ctx.check(ctx.lib.trt_harmonize(ctx.h))
a = ctx.locus_stats(False, None, 0.01)
a = {k: v.copy() for k, v in a.items()}
b = ctx.locus_stats(False, None, 0.01)
print("statSTR %d x %d: scan %.3f ms = %.0f GB/s" % (L, S, ctx.last_scan_ms(), 6.0 * L * S / ctx.last_scan_ms() / 1e6))
for k in ("ac", "n_called", "n_called_nonstrict", "n_hom", "n_padded"):
    assert np.array_equal(a[k], b[k]), k
per_locus = np.add.reduceat(a["ac"][0].astype(np.int64), ctx.locus_off[:-1].astype(np.int64))
assert np.array_equal(per_locus, a["n_called"][0] - a["n_padded"][0] + a["n_called_nonstrict"][0])
for l in (0, L // 2, L - 1):
    gt = ctx.block_get_gt(l, 1)[0][:, :2].astype(np.int64)
    A = int(loci.n_alleles[l])
    assert np.array_equal(np.bincount(gt[gt >= 0], minlength=A), a["ac"][0][ctx.locus_off[l]:ctx.locus_off[l + 1]]), l
    assert int((~(gt == -1).any(axis=1)).sum()) == int(a["n_called"][0][l])
print("statSTR integers: idempotent, checksum ok, sampled rows match numpy")
rng = np.random.default_rng(SEED)
traits = np.hstack([rng.standard_normal((S, 1)), rng.standard_normal((S, 10))])
covars = np.hstack([np.full((S, 1), -1.0), traits])
covars = (covars - covars.mean(axis=0)) / np.maximum(covars.std(axis=0), 1e-300)
outcome = covars[:, 1].copy()
covars[:, 1] = 1.0
ctx.assoc_set_design(covars, outcome, np.arange(S, dtype=np.int32))
fast = ctx.assoc_ols(20.0)
fast = {k: v.copy() for k, v in fast.items()}
t_fast = ctx.last_scan_ms()
os.environ["TRT_ASSOC_GENERIC"] = "1"
gen = ctx.assoc_ols(20.0)
t_gen = ctx.last_scan_ms()
del os.environ["TRT_ASSOC_GENERIC"]
assert np.array_equal(fast["filter_code"], gen["filter_code"]) and np.array_equal(fast["n_tested"], gen["n_tested"])
assert np.array_equal(fast["ac_len"], gen["ac_len"])
worst = 0.0
for k in ("p", "coef", "se", "r2"):
    x, y = fast[k], gen[k]
    ok = ~(np.isnan(x) & np.isnan(y))
    rel = np.abs(x[ok] - y[ok]) / np.maximum(np.maximum(np.abs(x[ok]), np.abs(y[ok])), 1e-300)
    worst = max(worst, float(rel.max()) if rel.size else 0.0)
print("associaTR %d x %d: fast path %.2f ms, generic kernels %.2f ms, worst relative difference %.2e, %d loci tested" % (
    L, S, t_fast, t_gen, worst, int((fast["filter_code"] == 0).sum())))
assert worst < 1e-6
print("OK")
