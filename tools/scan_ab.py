"""A/B of the GT scan on the bench block (GPU): ring depth, cell width, stream-only ceiling.
    python tools/scan_ab.py [L] [S]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trtools_b200 import _lib, synth
L = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
SEED = 20261017
ctx = _lib.Context(0)
loci = synth.make_loci(L, seed=SEED)
ctx.block_begin(L, S, 2, "hipstr")
ctx.synth_fill(SEED, 0, loci.cum_freq, loci.miss_thresh, loci.half_thresh, with_format=False)
ctx.block_set_alleles(*synth.allele_tables(loci))
ctx.check(ctx.lib.trt_harmonize(ctx.h))
ref = None
def run(tag, env):
    global ref
    for k in ("TRT_SCAN_CELLS16", "TRT_SCAN_STAGES", "TRT_SCAN_STREAM_ONLY"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ts = []
    for i in range(6):
        st = ctx.locus_stats(False, None, 0.01, want=("ac", "n_called", "n_hom", "hwep"))
        if i >= 2:
            ts.append(ctx.last_scan_ms())
    ok = ""
    if "TRT_SCAN_STREAM_ONLY" not in env:
        snap = {k: v.copy() for k, v in st.items()}
        if ref is None:
            ref = snap
        ok = "same" if all(np.array_equal(ref[k], snap[k], equal_nan=True) for k in ref) else "DIFFERENT"
    t = float(np.mean(ts))
    print("%-28s scan %.3f ms  %.0f GB/s  frac %.3f  total kernels %.3f ms  %s" % (tag, t, 6.0 * L * S / t / 1e6, 6.0 * L * S / t / 1e6 / 6458.4, ctx.last_kernel_ms(), ok), flush=True)
run("cells8 (default)", {})
run("cells16", {"TRT_SCAN_CELLS16": "1"})
for st in (3, 4, 5, 6, 7):
    run("cells8 stages<=%d" % st, {"TRT_SCAN_STAGES": str(st)})
run("cells16 stages<=3", {"TRT_SCAN_CELLS16": "1", "TRT_SCAN_STAGES": "3"})
run("stream only cells8", {"TRT_SCAN_STREAM_ONLY": "1"})
run("stream only cells16", {"TRT_SCAN_STREAM_ONLY": "1", "TRT_SCAN_CELLS16": "1"})
run("cells8 again", {})
