"""A/B timing of two builds of the library on the SAME GPU box (run-to-run variance between boxes is ~5%).
usage: python tools/ab_scan.py libA.so libB.so"""
import os, subprocess, sys
code = r'''
import sys, os, numpy as np
sys.path.insert(0, '.')
from trtools_b200 import _lib, synth
ctx = _lib.Context(0)
L, S = 40000, 50000
loci = synth.make_loci(L, seed=20261017)
ctx.block_begin(L, S, 2, "hipstr")
ctx.synth_fill(20261017, 0, loci.cum_freq, loci.miss_thresh, loci.half_thresh, with_format=False)
ctx.block_set_alleles(*synth.allele_tables(loci))
ctx.check(ctx.lib.trt_harmonize(ctx.h))
ms = []
for i in range(8):
    ctx.locus_stats(False, None, 0.01, want=("het",))
    ms.append(ctx.last_scan_ms())
m = float(np.median(ms[3:]))
print(os.environ.get("TRTOOLS_B200_LIB", "default"), "scan ms %.3f  GB/s %.0f" % (m, 6.0 * L * S / m / 1e6), flush=True)
'''
for rep in range(2):
    for lib in sys.argv[1:]:
        env = dict(os.environ, TRTOOLS_B200_LIB=os.path.abspath(lib))
        subprocess.run([sys.executable, "-c", code], env=env)
