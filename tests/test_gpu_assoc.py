"""
GPU parity tests for the associaTR path (trt_assoc_set_design / trt_assoc_ols and the drop-in
``associaTR.main``) against the unmodified reference's outputs (plink2-pinned fixtures and synthetic
blocks) and against the oracle at full precision (1e-6 relative on p, beta, se, R^2; integers exact).
"""
import argparse
import json
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


def _assoc_args(out, vcf, traits, **kw):
    ns = argparse.Namespace(outfile=out, tr_vcf=vcf, phenotype_name="test_pheno", traits=traits, vcftype=None,
                            same_samples=True, sample_list=None, region=None, non_major_cutoff=0, beagle_dosages=False,
                            plotting_phenotype=None, paired_genotype_plot=False, plot_phenotype_residuals=False,
                            plotting_ci_alphas=[], imputed_ukb_strs_paper_period_check=False, block_size=100)
    for k, v in kw.items():
        assert hasattr(ns, k), k
        setattr(ns, k, v)
    return ns


def _compare_tsv(got: str, want: str):
    g, w = got.splitlines(), want.splitlines()
    assert len(g) == len(w)
    assert g[0] == w[0]
    for i, (a, b) in enumerate(zip(g[1:], w[1:])):
        ca, cb = a.split("\t"), b.split("\t")
        assert ca[:5] == cb[:5], (i, ca[:5], cb[:5])
        assert ca[9:] == cb[9:], (i, ca[9:], cb[9:])
        if cb[5] == "nan":
            assert ca[5:9] == cb[5:9], i
            continue
        # p is printed with 3 significant digits: allow the last digit to differ by one unit
        assert abs(float(ca[5]) - float(cb[5])) <= 1.01e-2 * float(cb[5]), (i, ca[5], cb[5])
        for c in (6, 7, 8):
            assert_close(float(ca[c]), float(cb[c]), "row %d col %d" % (i, c), rel=1e-6)


@pytest.mark.parametrize("key,traits,kw", [
    ("single", ["traits_0.npy"], {}),
    ("combined", ["traits_0.npy", "traits_1.npy"], {}),
    ("cutoff_5", ["traits_0.npy"], {"non_major_cutoff": 5}),
    ("cutoff_20", ["traits_0.npy", "traits_1.npy"], {"non_major_cutoff": 20}),
    ("single_40", ["traits_0.npy"], {"sample_list": "samples_6_to_45.txt"}),
    ("multi", ["traits_0.npy"], {"vcf": "many_samples_multiallelic.vcf.gz"}),
    ("multi_cutoff", ["traits_0.npy", "traits_1.npy"], {"vcf": "many_samples_multiallelic.vcf.gz", "non_major_cutoff": 6}),
])
def test_associatr_cli_matches_reference(golden_dir, data_dir, tmp_path, key, traits, kw):
    """associaTR.main on the reference's plink2-pinned fixtures == the unmodified reference's TSV."""
    import contextlib
    import io
    from trtools_b200 import associaTR
    want = json.load(open(os.path.join(golden_dir, "associatr.json")))[key]
    kw = dict(kw)
    vcf = os.path.join(data_dir, kw.pop("vcf", "many_samples_biallelic.vcf.gz"))
    if "sample_list" in kw:
        kw["sample_list"] = os.path.join(data_dir, kw["sample_list"])
    out = str(tmp_path / "assoc.tsv")
    with contextlib.redirect_stdout(io.StringIO()):
        associaTR.main(_assoc_args(out, vcf, [os.path.join(data_dir, t) for t in traits], **kw))
    _compare_tsv(open(out).read(), want)


@pytest.mark.parametrize("name", ["synth_small", "synth_wide"])
@pytest.mark.parametrize("key,cutoff,use_mask", [("assoc", 20, False), ("assoc_subset", 5, True)])
def test_assoc_kernels_vs_reference_and_oracle_on_synthetic(golden_dir, ctx, name, key, cutoff, use_mask):
    """C-ABI level on the synthetic golden blocks: filter codes / n_tested exact vs the reference's rows; p, beta,
    se, R^2 within 1e-6 of the oracle's full-precision OLS (pinv) on the same arrays."""
    from oracle import assoc as oassoc, trh as otrh
    from oracle.records import LocusAsVariant
    from trtools_b200 import _lib, block
    loci, extra, _ = fixture(golden_dir, name)
    traits = np.array(extra["traits"], dtype=float)
    S = loci[0].gt.shape[0]
    mask = np.array(extra["sample_mask"], dtype=bool) if use_mask else None
    design = oassoc.prepare_design([traits], S, mask)
    blk = block.build_block(ctx, "hipstr", [LocusAsVariant(l) for l in loci])
    ctx.assoc_set_design(design.covars, design.outcome, np.nonzero(design.sample_filter)[0].astype(np.int32))
    res = ctx.assoc_ols(cutoff)
    lines = extra[key].splitlines()[1:]
    reasons = {_lib.AF_NO_CALLED: 'No called samples', _lib.AF_ONE_ALLELE: 'Only one called allele',
               _lib.AF_NCOVARS: 'n covars >= n samples', _lib.AF_NON_MAJOR: 'non-major allele count<{}'.format(cutoff)}
    n_ok = 0
    for i, l in enumerate(loci):
        ref_cols = lines[i].split("\t")
        assert int(ref_cols[3]) == int(res["n_tested"][i]), i
        code = int(res["filter_code"][i])
        assert (ref_cols[4] == 'False') == (code == _lib.AF_OK), (i, ref_cols[4], code)
        if code != _lib.AF_OK:
            assert reasons[code] == ref_cols[4], (i, reasons[code], ref_cols[4])
            continue
        h = otrh.harmonize(l)
        loaded = oassoc.load_locus(l, h, design.sample_filter.copy(), cutoff)
        row = oassoc.regress_locus(loaded, design)
        what = "{} {} locus {}".format(name, key, i)
        assert_close(res["p"][i], row.p, what + " p", abs_tol=1e-300)
        assert_close(res["coef"][i] * design.pheno_std, row.coef, what + " coef")
        assert_close(res["se"][i] * design.pheno_std, row.se, what + " se")
        assert_close(res["r2"][i], row.r2, what + " r2", rel=1e-6, abs_tol=1e-12)
        n_ok += 1
    assert n_ok > 0


def test_assoc_extreme_pvalues_and_missingness(ctx):
    """Strong effects (p down to ~1e-300) and heavy per-locus missingness: the t-test tail and the exact
    down-dates stay within 1e-6 of scipy/numpy."""
    from oracle import assoc as oassoc, trh as otrh
    from oracle.records import Locus, LocusAsVariant
    from trtools_b200 import _lib, block
    rng = np.random.default_rng(3)
    S = 6000
    loci = []
    base = rng.integers(0, 3, size=(S, 2))
    for j, miss in enumerate([0.0, 0.02, 0.4, 0.9]):
        gt = np.concatenate([base, np.ones((S, 1), int)], axis=1).astype(np.int16)
        m = rng.random(S) < miss
        gt[m, 0] = -1
        gt[m, 1] = -2
        loci.append(Locus("hipstr", "1", 100 + 50 * j, "ACACACAC", ["ACACAC", "ACACACACACAC"],
                          {"START": 100 + 50 * j, "END": 107 + 50 * j, "PERIOD": 2}, gt))
    g = np.array([4.0, 3.0, 6.0])[base].sum(axis=1)
    pcs = rng.standard_normal((S, 10))
    for effect in (0.0, 0.05, 1.0, 5.0):
        trait = effect * g + pcs[:, 0] * 0.3 + rng.standard_normal(S)
        traits = np.hstack([trait[:, None], pcs])
        design = oassoc.prepare_design([traits], S, None)
        blk = block.build_block(ctx, "hipstr", [LocusAsVariant(l) for l in loci])
        ctx.assoc_set_design(design.covars, design.outcome, np.arange(S, dtype=np.int32))
        res = ctx.assoc_ols(0)
        for i, l in enumerate(loci):
            h = otrh.harmonize(l)
            row = oassoc.regress_locus(oassoc.load_locus(l, h, design.sample_filter.copy(), 0), design)
            what = "effect {} locus {}".format(effect, i)
            assert int(res["filter_code"][i]) == _lib.AF_OK
            assert_close(res["p"][i], row.p, what + " p", abs_tol=1e-305)
            assert_close(res["coef"][i] * design.pheno_std, row.coef, what + " coef", rel=1e-6, abs_tol=1e-12)
            assert_close(res["se"][i] * design.pheno_std, row.se, what + " se")
            assert_close(res["r2"][i], row.r2, what + " r2", rel=1e-6, abs_tol=1e-12)


@pytest.mark.parametrize("n_cov,S,use_subset", [(10, 3001, False), (3, 2048, True), (0, 1000, False)])
def test_assoc_tile_fast_path_matches_generic_kernels_and_oracle(ctx, monkeypatch, n_cov, S, use_subset):
    """The thread-per-locus TMA tile path (trt_assoc_tile.cu) against the generic kernels on the same block
    (several 256-locus tiles, one of them forced generic by a locus with > 14 alleles, S not a multiple of the
    40-sample chunk, optional sample subset) and against the oracle's pinv OLS on a sample of loci."""
    from oracle import assoc as oassoc, trh as otrh
    from oracle.records import synth_to_loci
    from trtools_b200 import _lib, synth
    L = 700
    sl = synth.make_loci(L, seed=77, max_alleles=16)
    # locus 300 gets 16 alleles (> 14): its whole 256-locus tile must take the generic kernels
    sl.alts[300] = [sl.ref[300] + "ACGT"[k % 4] * (k + 1) for k in range(15)]
    sl.n_alleles[300] = 16
    sl.cum_freq[300] = (np.arange(1, 17, dtype=np.float64) / 16 * 2 ** 32 - 1).astype(np.uint32)
    calls = synth.fill_calls(sl, S)
    rng = np.random.default_rng(5)
    traits = rng.standard_normal((S, 1 + n_cov))
    mask = (rng.random(S) < 0.7) if use_subset else None
    design = oassoc.prepare_design([traits], S, mask)
    idx = np.nonzero(design.sample_filter)[0].astype(np.int32)

    def run():
        ctx.block_begin(L, S, 2, "hipstr")
        ctx.block_set_gt(calls.gt)
        ctx.block_set_alleles(*synth.allele_tables(sl))
        ctx.check(ctx.lib.trt_harmonize(ctx.h))
        ctx.assoc_set_design(design.covars, design.outcome, idx)
        return ctx.assoc_ols(5)

    fast = run()
    monkeypatch.setenv("TRT_ASSOC_GENERIC", "1")
    gen = run()
    monkeypatch.delenv("TRT_ASSOC_GENERIC")
    assert np.array_equal(fast["filter_code"], gen["filter_code"])
    assert np.array_equal(fast["n_tested"], gen["n_tested"])
    assert np.array_equal(fast["ac_len"], gen["ac_len"])
    for k in ("p", "coef", "se", "r2", "std_g"):
        for i in range(L):
            assert_close(fast[k][i], gen[k][i], "{} locus {}".format(k, i), rel=1e-7, abs_tol=1e-300 if k == "p" else 1e-13)
    loci = synth_to_loci(sl, calls, with_fmt=False)
    n_ok = 0
    for i in list(range(0, L, 23)) + [300]:
        if int(fast["filter_code"][i]) != _lib.AF_OK:
            continue
        h = otrh.harmonize(loci[i])
        row = oassoc.regress_locus(oassoc.load_locus(loci[i], h, design.sample_filter.copy(), 5), design)
        assert_close(fast["p"][i], row.p, "p locus %d" % i, abs_tol=1e-300)
        assert_close(fast["coef"][i] * design.pheno_std, row.coef, "coef locus %d" % i, rel=1e-6, abs_tol=1e-12)
        assert_close(fast["se"][i] * design.pheno_std, row.se, "se locus %d" % i)
        assert_close(fast["r2"][i], row.r2, "r2 locus %d" % i, rel=1e-6, abs_tol=1e-12)
        n_ok += 1
    assert n_ok > 5


@pytest.mark.parametrize("n_cov,S,L,use_subset", [(10, 3001, 700, False), (3, 2048, 500, True), (14, 2500, 193, False),
                                                   (1, 70001, 260, True), (0, 300, 50, False)])
def test_assoc_tensor_path_matches_fp64_paths_and_oracle(ctx, monkeypatch, n_cov, S, L, use_subset):
    """The integer tensor-core path (trt_assoc_mma.cu: u8 x s8 mma.sync, exact int32 sums of 7 base-256 digits per design
    column) against the FP64 tile path and the generic kernels on the same block, and against the oracle's pinv OLS.
    The block mixes loci the integer form cannot hold — a 131-bp allele beside 1-bp steps (spread > 127 units), haploid
    calls ([a, -2] pads among the called samples), a locus with 16 alleles — with ragged sizes: S not a multiple of 32,
    L not a multiple of the 192-locus tile, more than one sample segment (S = 70001), K from 2 to 16."""
    from oracle import assoc as oassoc, trh as otrh
    from oracle.records import synth_to_loci
    from trtools_b200 import _lib, synth
    sl = synth.make_loci(L, seed=1000 + L, max_alleles=16)
    special = {}
    if L > 40:
        a = 7                                             # spread: a long allele next to single-bp steps
        sl.alts[a] = [sl.ref[a] + "T", sl.ref[a] + "T" * 131] + list(sl.alts[a][2:])
        special[a] = "spread"
        b = 21                                            # > 14 alleles: generic kernels
        sl.alts[b] = [sl.ref[b] + "ACGT"[k % 4] * (k + 1) for k in range(15)]
        sl.n_alleles[b] = 16
        sl.cum_freq[b] = (np.arange(1, 17, dtype=np.float64) / 16 * 2 ** 32 - 1).astype(np.uint32)
        special[b] = "wide"
    sl.n_alleles[:] = [1 + len(x) for x in sl.alts]
    calls = synth.fill_calls(sl, S)
    if L > 40:
        c = 33                                            # haploid calls among the called samples
        calls.gt[c, ::7, 1] = -2
        calls.gt[c, ::7, 0] = np.maximum(calls.gt[c, ::7, 0], 0)
        special[c] = "pads"
    rng = np.random.default_rng(S + L)
    traits = rng.standard_normal((S, 1 + n_cov))
    g0 = calls.gt[0, :, :2].astype(float).sum(axis=1)
    traits[:, 0] += 0.05 * g0
    if n_cov:
        traits[:, 1] *= 1e4                               # column scales far apart
        traits[:, -1] = traits[:, -1] * 1e-3 + 5.0
    mask = (rng.random(S) < 0.7) if use_subset else None
    design = oassoc.prepare_design([traits], S, mask)
    idx = np.nonzero(design.sample_filter)[0].astype(np.int32)

    def run():
        ctx.block_begin(L, S, 2, "hipstr")
        ctx.block_set_gt(calls.gt)
        ctx.block_set_alleles(*synth.allele_tables(sl))
        ctx.check(ctx.lib.trt_harmonize(ctx.h))
        ctx.assoc_set_design(design.covars, design.outcome, idx)
        return ctx.assoc_ols(5)

    mma = run()
    monkeypatch.setenv("TRT_ASSOC_NO_MMA", "1")
    fp64 = run()
    monkeypatch.setenv("TRT_ASSOC_GENERIC", "1")
    gen = run()
    monkeypatch.delenv("TRT_ASSOC_GENERIC")
    monkeypatch.delenv("TRT_ASSOC_NO_MMA")
    for other, name in ((fp64, "fp64 tiles"), (gen, "generic")):
        assert np.array_equal(mma["filter_code"], other["filter_code"]), name
        assert np.array_equal(mma["n_tested"], other["n_tested"]), name
        assert np.array_equal(mma["ac_len"], other["ac_len"]), name
        for k in ("p", "coef", "se", "r2", "std_g"):
            for i in range(L):
                assert_close(mma[k][i], other[k][i], "{} {} locus {} {}".format(name, k, i, special.get(i, "")), rel=2e-7,
                             abs_tol=1e-300 if k == "p" else 1e-13)
    loci = synth_to_loci(sl, calls, with_fmt=False)
    n_ok = 0
    for i in sorted(set(list(range(0, L, max(1, L // 8))) + list(special))):
        if int(mma["filter_code"][i]) != _lib.AF_OK:
            continue
        h = otrh.harmonize(loci[i])
        row = oassoc.regress_locus(oassoc.load_locus(loci[i], h, design.sample_filter.copy(), 5), design)
        assert_close(mma["p"][i], row.p, "p locus %d" % i, abs_tol=1e-300)
        assert_close(mma["coef"][i] * design.pheno_std, row.coef, "coef locus %d" % i, rel=1e-6, abs_tol=1e-12)
        assert_close(mma["se"][i] * design.pheno_std, row.se, "se locus %d" % i)
        assert_close(mma["r2"][i], row.r2, "r2 locus %d" % i, rel=1e-6, abs_tol=1e-12)
        n_ok += 1
    assert n_ok > 3
