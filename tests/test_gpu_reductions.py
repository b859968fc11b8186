"""
GPU parity tests of the qcSTR / compareSTR reductions (trt_qc_reduce, trt_compare) against the
outputs of the unmodified reference (tests/golden/reductions.json: qcSTR's record loop on many_samples.vcf.gz,
compareSTR.UpdateComparisonResults on the reference's own GangSTR pair) and against the oracle.  Integers exact; float
sums within 1e-6 relative (the reference's per-locus mean quality is a float32 mean: 1e-6 is also its own precision).
"""
import json
import os

import numpy as np
import pytest

from helpers import assert_close, assert_close_list

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from trtools_b200 import _lib
    return _lib.default_context()


@pytest.mark.parametrize("key,ignore", [("zero", False), ("ignore", True)])
def test_qc_reductions_match_reference(golden_dir, data_dir, ctx, key, ignore):
    from trtools_b200 import block, reductions
    from trtools_b200.vcf_ingest import NativeVCF
    want = json.load(open(os.path.join(golden_dir, "reductions.json")))
    w = want["runs"][key]
    idx = np.array(want["sample_index"], dtype=bool)
    v = NativeVCF(os.path.join(data_dir, "many_samples.vcf.gz"))
    v._prefetch = ("Q",)
    recs = [r for _, r in zip(range(150), v)]
    S = len(v.samples)
    sample_calls = np.zeros(S, np.int64)
    sample_qual = np.zeros(S, np.float64)
    locus_calls, locus_qual = [], []
    for i0 in (0, 64, 128):                                    # three blocks: the per-sample sums persist across them
        blk = block.build_block(ctx, "hipstr", recs[i0:i0 + 64], ("Q",))
        r = reductions.qc_block(blk, idx, sample_calls, sample_qual, "Q", ignore)
        locus_calls += r["locus_calls"].tolist()
        locus_qual += r["locus_quality"].tolist()
        # per-locus length allele counts of the selected samples == the scan with a group mask
        assert r["allele_counts"].shape[0] == int(blk.locus_off[-1])
    assert sample_calls[idx].tolist() == [int(x) for x in w["sample_calls"]]
    assert not sample_calls[~idx].any() and not sample_qual[~idx].any()
    assert locus_calls == w["locus_calls"]
    assert_close_list(sample_qual[idx].tolist(), w["per_sample_total_qual"], "per-sample quality", rel=1e-6)
    assert_close_list(locus_qual, w["per_locus"], "per-locus mean quality", rel=1e-6)


@pytest.mark.parametrize("key,ignore_phasing", [("phased", False), ("ignore_phasing", True)])
def test_compare_reductions_match_reference(golden_dir, data_dir, ctx, key, ignore_phasing):
    from trtools_b200 import block, reductions
    from trtools_b200.vcf_ingest import NativeVCF
    want = json.load(open(os.path.join(golden_dir, "reductions.json")))["compare"]
    w = want["runs"][key]
    v1 = NativeVCF(os.path.join(data_dir, "test_gangstr1.vcf.gz"))
    v2 = NativeVCF(os.path.join(data_dir, "test_gangstr2.vcf.gz"))
    shared = want["shared"]
    idxs = [np.array([v1.samples.index(s) for s in shared], np.int32), np.array([v2.samples.index(s) for s in shared], np.int32)]
    r2 = {(r.CHROM, r.POS): r for r in v2}
    pairs = [(r, r2[(r.CHROM, r.POS)]) for r in v1 if (r.CHROM, r.POS) in r2]
    assert len(pairs) == want["n_pairs"]
    sample = {k: np.zeros(len(shared), np.int64) for k in ("numcalls", "conc-seq-count", "conc-len-count")}
    numcalls, cs, cl = [], [], []
    tot = np.zeros(5)
    from trtools_b200 import _lib
    ctx2 = _lib.Context(0)
    for i0 in range(0, len(pairs), 300):
        chunk = pairs[i0:i0 + 300]
        blk2 = block.build_block(ctx2, "gangstr", [p[1] for p in chunk])
        blk1 = block.build_block(ctx, "gangstr", [p[0] for p in chunk])
        res = reductions.compare_blocks(blk1, blk2, idxs, ignore_phasing, sample)
        numcalls += res["numcalls"].tolist()
        cs += res["conc_seq"].tolist()
        cl += res["conc_len"].tolist()
        tot += res["len_sums"].sum(axis=0)
    ctx2.close()
    keep = [i for i, n in enumerate(numcalls) if n > 0]            # the reference skips loci nobody is called at in both
    assert [numcalls[i] for i in keep] == [int(x) for x in w["locus"]["numcalls"]]
    assert_close_list([cs[i] / numcalls[i] for i in keep], w["locus"]["metric-conc-seq"], "conc-seq", rel=1e-12)
    assert_close_list([cl[i] / numcalls[i] for i in keep], w["locus"]["metric-conc-len"], "conc-len", rel=1e-12)
    for k in sample:
        assert sample[k].tolist() == [int(x) for x in w["sample"][k]], k
    o = w["overall"]
    assert (sum(numcalls), sum(cs), sum(cl)) == (int(o["numcalls"]), int(o["conc_seq_count"]), int(o["conc_len_count"]))
    assert_close_list(tot.tolist(), [o["total_len_1"], o["total_len_2"], o["total_len_11"], o["total_len_12"], o["total_len_22"]],
                      "length sums", rel=1e-9)


def test_compare_flags_mixed_phasing_and_ploidy(ctx):
    """The two conditions the reference gives up on (compareSTR.py:573-586) come back as the reference's ValueError."""
    from oracle.records import Locus, LocusAsVariant
    from trtools_b200 import _lib, block, reductions

    def loc(gt):
        return Locus("hipstr", "1", 100, "ACACAC", ["ACAC", "ACACACAC"], {"START": 100, "END": 105, "PERIOD": 2},
                     np.array(gt, dtype=np.int16))
    idxs = [np.arange(3, dtype=np.int32), np.arange(3, dtype=np.int32)]
    ctx2 = _lib.Context(0)

    def run(g1, g2, ignore=False):
        sample = {k: np.zeros(3, np.int64) for k in ("numcalls", "conc-seq-count", "conc-len-count")}
        b2 = block.build_block(ctx2, "hipstr", [LocusAsVariant(loc(g2))])
        b1 = block.build_block(ctx, "hipstr", [LocusAsVariant(loc(g1))])
        return reductions.compare_blocks(b1, b2, idxs, ignore, sample), sample
    base = [[0, 1, 0], [1, 2, 0], [2, 2, 0]]
    res, sample = run(base, [[1, 0, 0], [1, 2, 0], [0, 2, 0]])
    assert res["numcalls"].tolist() == [3] and res["conc_seq"].tolist() == [2] and sample["conc-seq-count"].tolist() == [1, 1, 0]
    with pytest.raises(ValueError, match="different phasedness"):
        run(base, [[1, 0, 1], [1, 2, 0], [0, 2, 0]])
    res, _ = run(base, [[1, 0, 1], [1, 2, 0], [0, 2, 0]], ignore=True)
    assert res["conc_seq"].tolist() == [2]
    with pytest.raises(ValueError, match="different ploidy"):
        run(base, [[1, -2, 0], [1, 2, 0], [0, 2, 0]])
    ctx2.close()
