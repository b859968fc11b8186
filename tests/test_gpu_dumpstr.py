"""
GPU parity tests for the dumpSTR path (trt_call_filters / trt_locus_filters and the drop-in
``dumpSTR.main``) against golden outputs of the unmodified reference and against the oracle on seeded
synthetic blocks.  Integer counters, filter flags and masked genotypes are bit-exact; HET/HWEP within 1e-6.
"""
import argparse
import os

import numpy as np
import pytest

from helpers import assert_close

pytestmark = pytest.mark.gpu
_cache = {}


def fixture(golden_dir, name):
    from oracle.records import load_loci
    if name not in _cache:
        _cache[name] = load_loci(os.path.join(golden_dir, name + ".npz"))
    return _cache[name]


@pytest.fixture(scope="module")
def ctx():
    from trtools_b200 import _lib
    return _lib.default_context()


def dump_args(out, vcf, **kw):
    ns = argparse.Namespace(
        vcf=vcf, vcftype="auto", out=out, zip=False, min_locus_callrate=None, min_locus_hwep=None,
        min_locus_het=None, max_locus_het=None, use_length=False, filter_regions=None, filter_regions_names=None,
        filter_hrun=False, drop_filtered=False, hipstr_min_call_DP=None, hipstr_max_call_DP=None,
        hipstr_min_call_Q=None, hipstr_max_call_flank_indel=None, hipstr_max_call_stutter=None,
        hipstr_min_supp_reads=None, longtr_min_call_DP=None, longtr_max_call_DP=None, longtr_min_call_Q=None,
        longtr_max_call_flank_indel=None, longtr_min_supp_reads=None, gangstr_expansion_prob_het=None,
        gangstr_expansion_prob_hom=None, gangstr_expansion_prob_total=None, gangstr_filter_span_only=False,
        gangstr_filter_spanbound_only=False, gangstr_filter_badCI=None, gangstr_min_call_DP=None,
        gangstr_max_call_DP=None, gangstr_min_call_Q=None, advntr_min_call_DP=None, advntr_max_call_DP=None,
        advntr_min_spanning=None, advntr_min_flanking=None, advntr_min_ML=None, eh_min_ADFL=None, eh_min_ADIR=None,
        eh_min_ADSP=None, eh_min_call_LC=None, eh_max_call_LC=None, popstr_min_call_DP=None, popstr_max_call_DP=None,
        popstr_require_support=None, num_records=None, die_on_warning=False, verbose=False, block_size=700)
    for k, v in kw.items():
        assert hasattr(ns, k), k
        setattr(ns, k, v)
    return ns


CLI_CASES = [
    ("hipstr_many", "many_samples.vcf.gz", "dump_numeric", "hipstr"),
    ("hipstr_many", "many_samples.vcf.gz", "dump_uselength", "hipstr"),
    ("hipstr_trio", "trio_chr21_hipstr.sorted.vcf.gz", "dump_numeric", "hipstr"),
    ("hipstr_trio", "trio_chr21_hipstr.sorted.vcf.gz", "dump_hipstr_filters", "hipstr"),   # incl. host-side MinSuppReads
    ("hipstr_trio", "trio_chr21_hipstr.sorted.vcf.gz", "dump_locus_only", "hipstr"),
    ("gangstr", "test_gangstr_head.vcf", "dump", "gangstr"),
    ("popstr", "test_popstr.vcf", "dump", "popstr"),
    ("eh", "test_ExpansionHunter.vcf", "dump", "eh"),
    ("advntr", "test_advntr.vcf", "dump", "advntr"),
    ("longtr", "test_longtr.vcf", "dump", "longtr"),
]


@pytest.mark.parametrize("name,vcf,key,vcftype", CLI_CASES)
def test_dumpstr_cli_matches_reference(golden_dir, data_dir, tmp_path, name, vcf, key, vcftype):
    """dumpSTR.main: samplog / loclog byte-exact; FILTER column, per-call FILTER strings, masked genotypes and
    AC/REFAC exact; HET/HWEP to 1e-6 against what the unmodified reference produced for the same flags."""
    import warnings
    from trtools_b200 import dumpSTR, cyvcf2_compat
    _, extra, _ = fixture(golden_dir, name)
    ref = extra[key]
    out = str(tmp_path / "o")
    kw = dict(ref["args"])
    kw.setdefault("vcftype", vcftype)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert dumpSTR.main(dump_args(out, os.path.join(data_dir, vcf), **kw)) == 0
    assert open(out + ".samplog.tab").read() == ref["samplog"]
    assert open(out + ".loclog.tab").read() == ref["loclog"]
    outvcf = cyvcf2_compat.VCF(out + ".vcf")
    n = 0
    for i, rec in enumerate(outvcf):
        r = ref["per_locus"][i]
        assert rec._filter_raw == r["filter"], (i, rec._filter_raw, r["filter"])
        assert rec.INFO["HRUN"] == r["HRUN"], i
        ac = rec.INFO["AC"]
        ac = [] if (ac == 0 and len(r["AC"]) == 0) else ([ac] if isinstance(ac, int) else list(ac))
        assert ac == r["AC"] or (r["AC"] == [] and ac == [0]), (i, ac, r["AC"])
        assert rec.INFO["REFAC"] == r["REFAC"], i
        assert_close(rec.INFO["HET"], r["HET"], "HET %d" % i, rel=1e-5)      # INFO floats are written with %g, read as float32
        assert_close(rec.INFO["HWEP"], r["HWEP"], "HWEP %d" % i, rel=2e-5, abs_tol=1e-300)
        if i < len(ref["calls"]):
            assert [str(x) for x in rec.format("FILTER")] == ref["calls"][i]["filter_text"], i
            assert rec.genotype.array().astype(int).tolist() == ref["calls"][i]["gt"], i
        n += 1
    assert n == len(ref["per_locus"])


def _specs_from_args(argd, blk):
    from trtools_b200 import _lib
    specs, names = [], []
    if argd.get("hipstr_max_call_flank_indel") is not None:
        specs.append((_lib.CF_RATIO_GT, blk.fmt_slot["DFLANKINDEL"], argd["hipstr_max_call_flank_indel"]))
    if argd.get("hipstr_max_call_stutter") is not None:
        specs.append((_lib.CF_RATIO_GT, blk.fmt_slot["DSTUTTER"], argd["hipstr_max_call_stutter"]))
    if argd.get("hipstr_min_call_DP") is not None:
        specs.append((_lib.CF_MIN, blk.fmt_slot["DP"], argd["hipstr_min_call_DP"]))
    if argd.get("hipstr_max_call_DP") is not None:
        specs.append((_lib.CF_MAX, blk.fmt_slot["DP"], argd["hipstr_max_call_DP"]))
    if argd.get("hipstr_min_call_Q") is not None:
        specs.append((_lib.CF_MIN, blk.fmt_slot["Q"], argd["hipstr_min_call_Q"]))
    return specs


@pytest.mark.parametrize("name", ["synth_small", "synth_wide"])
@pytest.mark.parametrize("key", ["dump", "dump_all"])
def test_call_and_locus_filter_kernels_vs_reference_on_synthetic(golden_dir, ctx, name, key):
    """C-ABI level: trt_call_filters + trt_locus_filters on the synthetic golden blocks."""
    from oracle.records import LocusAsVariant
    from trtools_b200 import _lib, block
    loci, extra, _ = fixture(golden_dir, name)
    ref = extra[key]
    a = ref["args"]
    recs = [LocusAsVariant(l) for l in loci]
    blk = block.build_block(ctx, "hipstr", recs, ["DP", "DSTUTTER", "DFLANKINDEL", "Q"])
    specs = _specs_from_args(a, blk)
    assert len(specs) == len(ref["call_filter_names"])
    S = blk.S
    counts = np.zeros((len(specs), S), np.int64)
    numcalls = np.zeros(S, np.int64)
    totaldp = np.zeros(S)
    res = ctx.call_filters(specs, blk.fmt_slot["DP"], counts, numcalls, totaldp, want_mask=True, want_trigger=True)
    assert res["negative_dp_locus"] == -1
    # sample log columns (integers bit-exact)
    lines = ref["samplog"].splitlines()[1:]
    for s in range(S):
        cols = lines[s].split("\t")
        assert int(cols[1]) == numcalls[s], s
        for f in range(len(specs)):
            assert int(cols[3 + f]) == counts[f, s], (s, f)
        want_mean = float(cols[2])
        got_mean = totaldp[s] / numcalls[s] if numcalls[s] > 0 else 0.0
        assert_close(got_mean, want_mean, "meanDP %d" % s, rel=1e-12)
    # masked genotypes + per-call filter strings for the captured loci
    from trtools_b200.dumpSTR import _filter_text, _BlockFilterResult
    from trtools_b200 import filters as F
    for i, c in enumerate(ref["calls"]):
        assert res["gt_masked"][i].astype(int).tolist() == c["gt"], i
    # locus filters on the masked genotypes
    lspecs = []
    if a.get("min_locus_callrate") is not None: lspecs.append((_lib.LF_CALLRATE, a["min_locus_callrate"]))
    if a.get("min_locus_hwep") is not None: lspecs.append((_lib.LF_HWE, a["min_locus_hwep"]))
    if a.get("min_locus_het") is not None: lspecs.append((_lib.LF_HETLOW, a["min_locus_het"]))
    if a.get("max_locus_het") is not None: lspecs.append((_lib.LF_HETHIGH, a["max_locus_het"]))
    if a.get("filter_hrun"): lspecs.append((_lib.LF_HRUN, 0.0))
    lres = ctx.locus_filters(lspecs, bool(a.get("use_length", False)))
    names = ref["locus_filter_names"]
    for i, r in enumerate(ref["per_locus"]):
        got = [names[b] for b in range(len(lspecs)) if (int(lres["flags"][i]) >> b) & 1]
        if int(lres["flags"][i]) & 0x80000000:
            got.append("NO_CALLS_REMAINING")
        assert (";".join(got) if got else "PASS") == r["filter"], i
        sl = blk.allele_slice(i)
        assert lres["ac"][sl][1:].tolist() == r["AC"] and int(lres["ac"][sl][0]) == r["REFAC"], i
        assert int(lres["hrun"][i]) == r["HRUN"]
        assert_close(lres["het"][i], r["HET"], "HET %d" % i)
        assert_close(lres["hwep"][i], r["HWEP"], "HWEP %d" % i, abs_tol=1e-300)


@pytest.mark.parametrize("L,S", [(300, 3001), (40, 20480), (700, 4096), (130, 6152), (33, 2900)])
def test_call_filters_vs_oracle_large(ctx, L, S):
    """Device-generated FORMAT arrays (trt_synth_fill) through the call-filter kernel vs the oracle on sampled
    loci, and the per-sample accumulators vs a numpy restatement over the whole block."""
    from oracle import dumpstr as od
    from oracle.records import synth_to_loci
    from trtools_b200 import _lib, synth
    sl = synth.make_loci(L, seed=L)
    ctx.block_begin(L, S, 2, "hipstr")
    ctx.synth_fill(L, 0, sl.cum_freq, sl.miss_thresh, sl.half_thresh, with_format=True)
    ctx.block_set_alleles(*synth.allele_tables(sl))
    ctx._current_block = None
    ctx.harmonize()
    specs = [(_lib.CF_RATIO_GT, _lib.FMT_DFLANKINDEL, 0.15), (_lib.CF_MIN, _lib.FMT_DP, 20), (_lib.CF_MIN, _lib.FMT_Q, 0.97)]
    counts = np.zeros((3, S), np.int64)
    numcalls = np.zeros(S, np.int64)
    totaldp = np.zeros(S)
    res = ctx.call_filters(specs, _lib.FMT_DP, counts, numcalls, totaldp, want_mask=True, want_trigger=False)
    calls = synth.fill_calls(sl, S)
    # whole-block restatement of the accumulators (numpy, float32 compare for Q as numpy does)
    nocall = np.any(calls.gt[:, :, :2] == -1, axis=2)
    with np.errstate(divide="ignore", invalid="ignore"):
        f0 = (calls.dflankindel / calls.dp) > 0.15
    f1 = calls.dp < 20
    f2 = calls.q < np.float32(0.97)
    for f, arr in enumerate((f0, f1, f2)):
        assert np.array_equal(counts[f], np.sum(arr & ~nocall, axis=0)), f
    passed = ~(f0 | f1 | f2) & ~nocall
    assert np.array_equal(numcalls, passed.sum(axis=0))
    assert np.array_equal(totaldp, np.where(passed, calls.dp, 0).sum(axis=0).astype(float))
    want_mask = (f0.astype(np.uint32) | (f1.astype(np.uint32) << 1) | (f2.astype(np.uint32) << 2) |
                 (nocall.astype(np.uint32) << 31))
    assert np.array_equal(res["call_mask"], want_mask)
    filtered = (f0 | f1 | f2) & ~nocall
    want_gt = calls.gt.copy()
    want_gt[filtered] = (-1, -1, 0)
    assert np.array_equal(res["gt_masked"], want_gt)
    # oracle agreement on a few loci (sample-log increments)
    cf = [od.hipstr_flank_indels(0.15), od.min_value("HipSTRCallMinDepth", "DP", 20), od.min_value("HipSTRCallMinQ", "Q", 0.97)]
    loci = synth_to_loci(sl, calls)
    for j in (0, L // 2, L - 1):
        sinfo = od.new_sample_info(S, cf)
        r = od.apply_call_filters(loci[j], cf, sinfo)
        assert np.array_equal(r.gt, res["gt_masked"][j])


def test_filter_classes_are_drop_in(ctx, golden_dir):
    """The reference's per-record protocol: filter(record) -> float[S] (NaN = keep) / value-or-None."""
    from oracle.records import LocusAsVariant
    from oracle import dumpstr as od, trh as otrh
    from trtools_b200 import filters as F, tr_harmonizer as trh
    loci, _, _ = fixture(golden_dir, "synth_small")
    l = loci[3]
    rec = trh.HarmonizeRecord("hipstr", LocusAsVariant(l))
    for mine, theirs in [(F.CallFilterMinValue("HipSTRCallMinDepth", "DP", 25), od.min_value("HipSTRCallMinDepth", "DP", 25)),
                         (F.CallFilterMaxValue("HipSTRCallMaxDepth", "DP", 40), od.max_value("HipSTRCallMaxDepth", "DP", 40)),
                         (F.HipSTRCallFlankIndels(0.1), od.hipstr_flank_indels(0.1)),
                         (F.HipSTRCallStutter(0.1), od.hipstr_stutter(0.1)),
                         (F.CallFilterMinValue("HipSTRCallMinQ", "Q", 0.98), od.min_value("HipSTRCallMinQ", "Q", 0.98))]:
        assert mine.name == theirs.name
        got, want = mine(rec), theirs(l, l.gt)
        assert np.array_equal(np.isnan(got), np.isnan(want))
        assert np.allclose(got[~np.isnan(got)], want[~np.isnan(want)], rtol=0, atol=0)
    h = otrh.harmonize(l)
    for mine, theirs in [(F.Filter_MinLocusCallrate(0.99), od.LocusFilter("callrate", 0.99)),
                         (F.Filter_MinLocusHWEP(0.5, True), od.LocusFilter("hwe", 0.5, True)),
                         (F.Filter_MinLocusHet(0.9), od.LocusFilter("hetlow", 0.9)),
                         (F.Filter_MaxLocusHet(0.1), od.LocusFilter("hethigh", 0.1)),
                         (F.Filter_LocusHrun(), od.LocusFilter("hrun"))]:
        assert mine.filter_name() == theirs.filter_name()
        got, want = mine(rec), theirs(l, h, l.gt)
        assert (got is None) == (want is None), mine.filter_name()
        if got is not None:
            assert_close(got, want, mine.filter_name())


@pytest.mark.parametrize("L,S", [(97, 4104), (20, 2048)])
def test_call_filter_variants_tma_vs_legacy_vs_numpy(ctx, monkeypatch, L, S):
    """Every comparison variant of the TMA call-filter kernel (int/float min/max, stutter ratio with zero and missing
    depths, host-evaluated values) against the legacy kernel on the same block and against numpy."""
    from trtools_b200 import _lib, synth
    sl = synth.make_loci(L, seed=1000 + L)
    calls = synth.fill_calls(sl, S)
    rng = np.random.default_rng(L)
    dp = calls.dp.copy()
    dp[rng.random(dp.shape) < 0.01] = 0                      # 0/0 and x/0 ratios
    dp[rng.random(dp.shape) < 0.01] = np.iinfo(np.int32).min  # missing depth
    host = np.where(rng.random((L, S)) < 0.05, rng.random((L, S)) * 10, np.nan).astype(np.float32)
    specs = [(_lib.CF_MAX, _lib.FMT_DP, 45), (_lib.CF_RATIO_GT, _lib.FMT_DSTUTTER, 0.1), (_lib.CF_MAX, _lib.FMT_Q, 0.9999),
             (_lib.CF_MIN, _lib.FMT_DP, 12.5), (_lib.CF_HOST_VALUE, _lib.FMT_AUX0, 0.0)]

    def run():
        ctx.block_begin(L, S, 2, "hipstr")
        ctx.block_set_gt(calls.gt)
        ctx.block_set_format(_lib.FMT_DP, dp)
        ctx.block_set_format(_lib.FMT_DSTUTTER, calls.dstutter)
        ctx.block_set_format(_lib.FMT_Q, calls.q)
        ctx.block_set_format(_lib.FMT_AUX0, host)
        ctx.block_set_alleles(*synth.allele_tables(sl))
        counts = np.zeros((len(specs), S), np.int64)
        numcalls = np.zeros(S, np.int64)
        totaldp = np.zeros(S)
        res = ctx.call_filters(specs, _lib.FMT_DP, counts, numcalls, totaldp, want_mask=True, want_trigger=False)
        return res, counts, numcalls, totaldp

    tma = run()
    monkeypatch.setenv("TRT_CF_LEGACY", "1")
    leg = run()
    monkeypatch.delenv("TRT_CF_LEGACY")
    for a, b in zip(tma[1:], leg[1:]):
        assert np.array_equal(a, b, equal_nan=True)
    assert np.array_equal(tma[0]["call_mask"], leg[0]["call_mask"])
    assert np.array_equal(tma[0]["gt_masked"], leg[0]["gt_masked"])
    assert tma[0]["negative_dp_locus"] == leg[0]["negative_dp_locus"]
    # numpy restatement (filters.py:327-484: int32 fields compare as float64, float32 fields in float32)
    nocall = np.any(calls.gt[:, :, :2] == -1, axis=2)
    with np.errstate(divide="ignore", invalid="ignore"):
        f = [dp.astype(np.float64) > 45, (calls.dstutter / dp) > 0.1, calls.q > np.float32(0.9999),
             dp.astype(np.float64) < 12.5, ~np.isnan(host)]
    want_mask = np.zeros((L, S), np.uint32)
    for i, fi in enumerate(f):
        assert np.array_equal(tma[1][i], np.sum(fi & ~nocall, axis=0)), i
        want_mask |= fi.astype(np.uint32) << i
    want_mask |= nocall.astype(np.uint32) << 31
    assert np.array_equal(tma[0]["call_mask"], want_mask)
    passed = ~np.logical_or.reduce(f) & ~nocall
    assert np.array_equal(tma[2], passed.sum(axis=0))
    missing = passed & (dp == np.iinfo(np.int32).min)
    want_dp = np.where(passed & (dp > 0), dp, 0).sum(axis=0).astype(float)
    want_dp[missing.any(axis=0)] = np.nan
    assert np.array_equal(tma[3], want_dp, equal_nan=True)


def test_dumpstr_zip_writes_an_indexed_bgzf_vcf(data_dir, tmp_path):
    """dumpSTR --zip: the output is BGZF with a tabix index beside it (written natively: the image has no tabix binary),
    the records equal the plain run's, and a region query through the index returns what a linear scan returns."""
    from trtools_b200 import dumpSTR, cyvcf2_compat
    from trtools_b200.vcf_ingest import NativeVCF
    src = os.path.join(data_dir, "many_samples.vcf.gz")
    kw = dict(vcftype="hipstr", hipstr_min_call_DP=20, min_locus_hwep=1e-4)
    plain, zipped = str(tmp_path / "plain"), str(tmp_path / "zipped")
    assert dumpSTR.main(dump_args(plain, src, **kw)) == 0
    assert dumpSTR.main(dump_args(zipped, src, zip=True, **kw)) == 0
    assert os.path.isfile(zipped + ".vcf.gz") and os.path.isfile(zipped + ".vcf.gz.tbi")
    a = [str(r) for r in cyvcf2_compat.TextVCF(plain + ".vcf")]
    b = [str(r) for r in cyvcf2_compat.TextVCF(zipped + ".vcf.gz")]
    assert a == b and len(a) > 100
    v = NativeVCF(zipped + ".vcf.gz")
    for region in ("1:100000-200000", "1:3000000-3050000"):
        want = [(r.CHROM, r.POS) for r in cyvcf2_compat.TextVCF(plain + ".vcf")(region)]
        got = [(r.CHROM, r.POS) for r in v(region)]
        assert got == want and (v._region_stop or v._region_empty), region
