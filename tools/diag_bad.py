"""Diagnostic (GPU): repeat the statSTR pass on one resident block and report every locus whose integer outputs
differ between repetitions (cold first call vs warm calls)."""
import os, sys
import numpy as np
sys.path.insert(0, '.')
os.environ["TRT_DEBUG_BAD"] = "1"
from trtools_b200 import _lib, synth, _lib as L_
SEED = 20261017
ctx = _lib.Context(int(os.environ.get("DIAG_DEV", "0")))
L = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
S = 50000
loci = synth.make_loci(L, seed=SEED)
ctx.block_begin(L, S, 2, "hipstr")
ctx.synth_fill(SEED, 0, loci.cum_freq, loci.miss_thresh, loci.half_thresh, with_format=False)
ctx.block_set_alleles(*synth.allele_tables(loci))
ctx.check(ctx.lib.trt_harmonize(ctx.h))
runs = []
for rep in range(4):
    try:
        st = ctx.locus_stats(False, None, 0.01)
        runs.append({k: st[k].copy() for k in ("ac", "n_called", "n_called_nonstrict", "n_hom", "n_padded")})
        print("rep %d ok scan %.3f ms sum n_called %d" % (rep, ctx.last_scan_ms(), int(st["n_called"].sum())), flush=True)
    except Exception as e:
        runs.append(None)
        print("rep %d FAILED: %s" % (rep, e), flush=True)
ref = runs[-1]
off = ctx.locus_off
for rep, r in enumerate(runs[:-1]):
    if r is None or ref is None:
        continue
    bad = np.zeros(L, bool)
    for k in ("n_called", "n_called_nonstrict", "n_hom", "n_padded"):
        bad |= (r[k][0] != ref[k][0])
    dac = np.nonzero(r["ac"][0] != ref["ac"][0])[0]
    bad[np.searchsorted(off, dac, side="right") - 1] = True
    idx = np.nonzero(bad)[0]
    print("rep %d vs last: %d loci differ" % (rep, len(idx)), flush=True)
    for l in idx[:20]:
        A = int(loci.n_alleles[l])
        sl = slice(int(off[l]), int(off[l + 1]))
        print("  locus %d A=%d  n_called %d/%d nonstrict %d/%d hom %d/%d  ac %s / %s" % (
            l, A, r["n_called"][0][l], ref["n_called"][0][l], r["n_called_nonstrict"][0][l], ref["n_called_nonstrict"][0][l],
            r["n_hom"][0][l], ref["n_hom"][0][l], r["ac"][0][sl].tolist(), ref["ac"][0][sl].tolist()), flush=True)
