"""Calibration helper (GPU): scan kernel time vs the HBM-read ceiling of its own TMA ring, plus the
consumer-side cycle breakdown (TRT_SCAN_DEBUG)."""
import sys, os, numpy as np
sys.path.insert(0, '.')
from trtools_b200 import _lib, synth
ctx = _lib.Context(0)
L, S = int(os.environ.get("CAL_L", "40000")), 50000
loci = synth.make_loci(L, seed=1, max_alleles=int(os.environ.get("CAL_MAXA", "16")))
ctx.block_begin(L, S, 2, "hipstr")
ctx.synth_fill(1, 0, loci.cum_freq, loci.miss_thresh, loci.half_thresh, with_format=False)
ctx.block_set_alleles(*synth.allele_tables(loci))
ctx.check(ctx.lib.trt_harmonize(ctx.h))
for mode in ("normal", "debug", "stream"):
    os.environ.pop("TRT_SCAN_DEBUG", None)
    if mode == "stream": os.environ["TRT_SCAN_STREAM_ONLY"] = "1"
    if mode == "debug": os.environ["TRT_SCAN_DEBUG"] = "1"
    ms = []
    for i in range(5):
        try:
            ctx.locus_stats(False, None, 0.01, want=("het",))
        except Exception as e:
            pass
        ms.append(ctx.last_scan_ms())
    m = float(np.median(ms[2:]))
    print(mode, "scan ms", m, "GB/s", 6.0 * L * S / m / 1e6, flush=True)
