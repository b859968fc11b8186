"""Debug driver for the associaTR tensor path: small golden block, per-locus n_tested under the three paths."""
import os, sys
import numpy as np
sys.path.insert(0, '.')
from oracle.records import load_loci, LocusAsVariant
from oracle import assoc as oassoc
from trtools_b200 import _lib, block
loci, extra, _ = load_loci('tests/golden/synth_small.npz')
traits = np.array(extra["traits"], dtype=float)
S = loci[0].gt.shape[0]
design = oassoc.prepare_design([traits], S, None)
ctx = _lib.default_context()
def run():
    blk = block.build_block(ctx, "hipstr", [LocusAsVariant(l) for l in loci])
    ctx.assoc_set_design(design.covars, design.outcome, np.nonzero(design.sample_filter)[0].astype(np.int32))
    return ctx.assoc_ols(20)
a = run()
os.environ["TRT_ASSOC_NO_MMA"] = "1"
b = run()
del os.environ["TRT_ASSOC_NO_MMA"]
called = np.array([int(((l.gt[:, :2] >= 0).all(axis=1)).sum()) for l in loci])
print("K", design.covars.shape, "S", S)
print("mma  n:", a["n_tested"][:24].tolist())
print("fp64 n:", b["n_tested"][:24].tolist())
print("want n:", called[:24].tolist())
print("diff loci:", np.nonzero(a["n_tested"] != b["n_tested"])[0].tolist()[:40])
print("std_g mma ", a["std_g"][:6]); print("std_g fp64", b["std_g"][:6])
print("p mma ", a["p"][:6]); print("p fp64", b["p"][:6])
