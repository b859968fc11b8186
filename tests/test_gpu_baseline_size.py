"""
Parity at the BASELINE.json sample counts (configs[2], [3]: 50 000 samples; configs[4]: 500 000 samples) for the two
tools whose kernels accumulate over the sample axis in a device-specific order: dumpSTR (call filters min-call-DP 20 +
max-call-flank-indel 0.15, locus filter HWE 1e-4 — C3's flags) and associaTR (trait ~ TR length + 10 PCs,
non-major cutoff 20 — C4).  The blocks are generated on the device (trt_synth_fill) and, bit-identically, on the host
(synth.fill_calls) for the oracle.  Integers (masks, counters, flags, n_tested, filter codes, allele counts) are
compared bit-exact; p / beta / se / R^2 / HET / HWEP within 1e-6 relative (tests/helpers.py REL_TOL).
"""
import numpy as np
import pytest

from helpers import assert_close

pytestmark = pytest.mark.gpu

S_BASE = 50000
S_BIOBANK = 500000


@pytest.fixture(scope="module")
def ctx():
    from trtools_b200 import _lib
    return _lib.default_context()


def _device_block(ctx, L, S, seed, with_format):
    from trtools_b200 import synth
    sl = synth.make_loci(L, seed=seed)
    ctx.block_begin(L, S, 2, "hipstr")
    ctx.synth_fill(seed, 0, sl.cum_freq, sl.miss_thresh, sl.half_thresh, with_format=with_format)
    ctx.block_set_alleles(*synth.allele_tables(sl))
    ctx._current_block = None
    ctx.check(ctx.lib.trt_harmonize(ctx.h))
    return sl


def test_dumpstr_c3_flags_at_50k_samples_vs_oracle(ctx):
    """trt_call_filters + trt_locus_filters at S = 50 000 on 40 loci against the oracle's ApplyCallFilters /
    ApplyLocusFilters / INFO recompute (dumpSTR.py:613-774, 917-973, 1307-1336) on every locus."""
    from oracle import dumpstr as od, trh as otrh
    from oracle.records import synth_to_loci
    from trtools_b200 import _lib, synth
    L, S = 40, S_BASE
    sl = _device_block(ctx, L, S, 4242, with_format=True)
    specs = [(_lib.CF_MIN, _lib.FMT_DP, 20), (_lib.CF_RATIO_GT, _lib.FMT_DFLANKINDEL, 0.15)]
    counts = np.zeros((2, S), np.int64)
    numcalls = np.zeros(S, np.int64)
    totaldp = np.zeros(S)
    res = ctx.call_filters(specs, _lib.FMT_DP, counts, numcalls, totaldp, want_mask=True, want_trigger=False)
    assert res["negative_dp_locus"] == -1
    lres = ctx.locus_filters([(_lib.LF_HWE, 1e-4)], False)

    calls = synth.fill_calls(sl, S)
    loci = synth_to_loci(sl, calls)
    cf = [od.min_value("HipSTRCallMinDepth", "DP", 20), od.hipstr_flank_indels(0.15)]
    lf = [od.LocusFilter("hwe", 1e-4, False)]
    sinfo, linfo = od.new_sample_info(S, cf), od.new_loc_info(lf)
    off = ctx.locus_off
    n_flagged = 0
    for j, l in enumerate(loci):
        h = otrh.harmonize(l)
        r = od.apply_call_filters(l, cf, sinfo)
        assert np.array_equal(r.gt, res["gt_masked"][j]), j
        nocall = ~otrh.called_samples(l.gt)
        want_mask = np.zeros(S, np.uint32)
        for f, filt in enumerate(cf):
            want_mask |= (~np.isnan(filt(l, l.gt))).astype(np.uint32) << f
        want_mask |= nocall.astype(np.uint32) << 31
        assert np.array_equal(res["call_mask"][j], want_mask), j
        filtered, text = od.apply_locus_filters(l, h, r.gt, lf, linfo)
        got = []
        if int(lres["flags"][j]) & 1:
            got.append(lf[0].filter_name())
        if int(lres["flags"][j]) & 0x80000000:
            got.append("NO_CALLS_REMAINING")
        assert (";".join(got) if got else "PASS") == text, (j, got, text)
        n_flagged += bool(got)
        info = od.recompute_info(h, r.gt, False)
        ac = lres["ac"][off[j]:off[j + 1]]
        assert ac[1:].tolist() == info["AC"] and int(ac[0]) == info["REFAC"], j
        assert int(lres["n_called"][j]) == int(np.sum(otrh.called_samples(r.gt))), j
        assert int(lres["hrun"][j]) == info["HRUN"], j
        assert_close(lres["het"][j], info["HET"], "HET %d" % j)
        assert_close(lres["hwep"][j], info["HWEP"], "HWEP %d" % j, abs_tol=1e-300)
    # per-sample accumulators over the block (dumpSTR.py:700-713)
    assert np.array_equal(numcalls, sinfo["numcalls"])
    assert np.array_equal(counts[0], sinfo[cf[0].name]) and np.array_equal(counts[1], sinfo[cf[1].name])
    assert np.array_equal(totaldp, sinfo["totaldp"], equal_nan=True)
    assert int(numcalls.sum()) > 0.3 * L * S


def _bench_design(S, seed, n_cov=10):
    from oracle import assoc as oassoc
    rng = np.random.default_rng(seed)
    traits = np.hstack([rng.standard_normal((S, 1)), rng.standard_normal((S, n_cov))])
    return oassoc.prepare_design([traits], S, None), traits


def _check_assoc_vs_oracle(ctx, sl, S, design, cutoff, loci_idx, tag):
    from oracle import assoc as oassoc, trh as otrh
    from oracle.records import synth_to_loci
    from trtools_b200 import _lib, synth
    ctx.assoc_set_design(design.covars, design.outcome, np.nonzero(design.sample_filter)[0].astype(np.int32))
    res = ctx.assoc_ols(cutoff)
    reasons = {_lib.AF_NO_CALLED: 'No called samples', _lib.AF_ONE_ALLELE: 'Only one called allele',
               _lib.AF_NCOVARS: 'n covars >= n samples', _lib.AF_NON_MAJOR: 'non-major allele count<{}'.format(cutoff)}
    off = ctx.locus_off
    n_ok = 0
    for j in loci_idx:
        calls = synth.fill_calls(sl, S, slice(j, j + 1))
        sub = synth.SynthLoci(seed=sl.seed, n_loci=1, chrom=sl.chrom[j:j + 1], pos=sl.pos[j:j + 1], start=sl.start[j:j + 1],
                              end=sl.end[j:j + 1], period=sl.period[j:j + 1], ref=sl.ref[j:j + 1], alts=sl.alts[j:j + 1],
                              n_alleles=sl.n_alleles[j:j + 1], cum_freq=sl.cum_freq[j:j + 1], locus_offset=j)
        l = synth_to_loci(sub, calls, with_fmt=False)[0]
        h = otrh.harmonize(l)
        loaded = oassoc.load_locus(l, h, design.sample_filter.copy(), cutoff)
        row = oassoc.regress_locus(loaded, design)
        what = "{} locus {}".format(tag, j)
        assert int(res["n_tested"][j]) == row.n_samples_tested, what
        code = int(res["filter_code"][j])
        assert (code == _lib.AF_OK) == (row.locus_filtered is False), (what, code, row.locus_filtered)
        # allele counts among the tested samples, keyed by allele index (the frequency detail column)
        tested = otrh.called_samples(l.gt) & design.sample_filter
        g = l.gt[tested, :2].astype(np.int64)
        assert np.array_equal(res["ac_len"][off[j]:off[j + 1]], np.bincount(g[g >= 0], minlength=off[j + 1] - off[j])), what
        if code != _lib.AF_OK:
            assert reasons[code] == row.locus_filtered, what
            continue
        assert_close(res["p"][j], row.p, what + " p", abs_tol=1e-300)
        assert_close(res["coef"][j] * design.pheno_std, row.coef, what + " coef")
        assert_close(res["se"][j] * design.pheno_std, row.se, what + " se")
        assert_close(res["r2"][j], row.r2, what + " r2", rel=1e-6, abs_tol=1e-12)
        n_ok += 1
    return n_ok


def test_associatr_c4_at_50k_samples_vs_oracle(ctx):
    """trt_assoc_ols with 10 PCs and non-major cutoff 20 at S = 50 000 (BASELINE configs[3]) against the oracle's
    load_trs + statsmodels-equivalent pinv OLS (lafg.py:157-259, associaTR.py:246-291) on 36 loci of a 600-locus block
    (two full 512-locus tiles of the fast path are exercised; the oracle runs on every 17th locus)."""
    L, S = 600, S_BASE
    sl = _device_block(ctx, L, S, 777, with_format=False)
    design, _ = _bench_design(S, 777)
    n_ok = _check_assoc_vs_oracle(ctx, sl, S, design, 20, list(range(0, L, 17)), "C4")
    assert n_ok >= 24


def test_associatr_c4_sample_subset_at_50k(ctx):
    """Same at S = 50 000 with a 70 % sample subset (the in-design flag of the z-table and the masked allele counts)."""
    from oracle import assoc as oassoc
    L, S = 300, S_BASE
    sl = _device_block(ctx, L, S, 778, with_format=False)
    rng = np.random.default_rng(778)
    traits = np.hstack([rng.standard_normal((S, 1)), rng.standard_normal((S, 10))])
    design = oassoc.prepare_design([traits], S, rng.random(S) < 0.7)
    n_ok = _check_assoc_vs_oracle(ctx, sl, S, design, 20, list(range(0, L, 29)), "C4 subset")
    assert n_ok >= 6


def test_associatr_c5_shape_at_500k_samples_vs_oracle(ctx):
    """One block of BASELINE configs[4]'s sample count (S = 500 000) against the ORACLE (not the library's own generic
    kernels) on 6 loci: FP64 accumulation over 5e5 terms must still land within 1e-6 of the pinv solution."""
    info = ctx.device_info()
    if info["free_mem_bytes"] < 8e9:
        pytest.skip("needs 8 GB of free HBM")
    L, S = 1024, S_BIOBANK
    sl = _device_block(ctx, L, S, 5005, with_format=False)
    design, _ = _bench_design(S, 5005)
    n_ok = _check_assoc_vs_oracle(ctx, sl, S, design, 20, [0, 255, 256, 511, 700, 1023], "C5")
    assert n_ok >= 4
