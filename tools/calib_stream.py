"""Calibration helper (GPU): GT scan time vs the HBM-read ceiling of its own TMA ring (TRT_SCAN_STREAM_ONLY: the
consumers only drain the ring), (a 2-3 chunks-per-stage variant was measured with this tool and dropped)."""
import sys, os, numpy as np
sys.path.insert(0, '.')
from trtools_b200 import _lib, synth
ctx = _lib.Context(0)
L, S = int(os.environ.get("CAL_L", "40000")), 50000
for maxa in [int(x) for x in os.environ.get("CAL_MAXA", "16,6").split(",")]:
    loci = synth.make_loci(L, seed=1, max_alleles=maxa)
    ctx.block_begin(L, S, 2, "hipstr")
    ctx.synth_fill(1, 0, loci.cum_freq, loci.miss_thresh, loci.half_thresh, with_format=False)
    ctx.block_set_alleles(*synth.allele_tables(loci))
    ctx.check(ctx.lib.trt_harmonize(ctx.h))
    ref = None
    for nsub in (1,):
        for mode in ("normal", "stream"):
            os.environ.pop("TRT_SCAN_STREAM_ONLY", None)
            if mode == "stream":
                os.environ["TRT_SCAN_STREAM_ONLY"] = "1"
            ms = []
            for i in range(5):
                try:
                    st = ctx.locus_stats(False, None, 0.01, want=("het", "ac", "n_called"))
                except Exception:
                    st = None
                ms.append(ctx.last_scan_ms())
            if mode == "normal" and st is not None:
                key = (st["ac"].tobytes(), st["n_called"].tobytes())
                ref = ref or key
                same = key == ref
            else:
                same = None
            m = float(np.median(ms[2:]))
            print("max_alleles %2d nsub %d %-6s scan %.3f ms  %.0f GB/s  results_equal_nsub1=%s" % (maxa, nsub, mode, m, 6.0 * L * S / m / 1e6, same), flush=True)
