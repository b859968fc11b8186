"""
GPU parity tests of the Beagle allele-probability dosage path: ``trt_dosages``
(TRRecord.GetDosages, tr_harmonizer.py:1098-1208) and ``trt_assoc_dosage_ols`` (the --beagle-dosages branch of
load_trs + the regression on summed dosages, lafg.py:175-238, associaTR.py:266-291) against outputs of the UNMODIFIED
reference (tests/golden/dosage_*.npz, associatr_dosage.json; generator tests/golden/make_golden.py ``dosage``) and
against the oracle.  Floats within 1e-6 relative (helpers.REL_TOL), text columns and integers exact.
"""
import argparse
import contextlib
import io
import json
import os

import numpy as np
import pytest

from helpers import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from trtools_b200 import _lib
    return _lib.default_context()


def _load(golden_dir, name):
    from oracle.records import load_loci
    loci, extra, _ = load_loci(os.path.join(golden_dir, name + ".npz"))
    for l in loci:                                   # the fixture pads AP columns to the widest locus
        for k in ("AP1", "AP2"):
            if k in l.fmt:
                l.fmt[k] = np.ascontiguousarray(l.fmt[k][:, :len(l.alts)], dtype=np.float32)
    return loci, extra


@pytest.mark.parametrize("name", ["dosage_real", "dosage_synth"])
def test_getdosages_matches_reference(golden_dir, ctx, name):
    """All four dosage types, strict and lenient, record by record through the block kernel."""
    from oracle.records import LocusAsVariant
    from trtools_b200 import block, tr_harmonizer as trh
    loci, extra = _load(golden_dir, name)
    recs = [LocusAsVariant(l) for l in loci]
    blk = block.build_block(ctx, "hipstr", recs)
    n_checked = 0
    for i, want in enumerate(extra["dosages"]):
        tr = trh.TRRecord._from_block(blk, i, recs[i])
        for kind in ("bestguess", "bestguess_norm", "beagleap", "beagleap_norm"):
            for strict in (True, False):
                w = want[kind + ("" if strict else "_lenient")]
                if "error" in w:
                    with pytest.raises(ValueError) as ei:
                        tr.GetDosages(trh.TRDosageTypes[kind], strict=strict)
                    assert str(ei.value) == w["error"], (i, kind)
                    continue
                got = tr.GetDosages(trh.TRDosageTypes[kind], strict=strict)
                assert len(got) == len(w["values"])
                if got.dtype != np.float32:          # the lenient NaN answer is a float64 array in the reference too
                    assert np.isnan(got).all() and all(np.isnan(x) for x in w["values"]), (i, kind)
                    continue
                for s, (a, b) in enumerate(zip(got, w["values"])):
                    assert_close(a, b, "{} locus {} {} sample {}".format(name, i, kind, s), rel=1e-6, abs_tol=1e-7)
                n_checked += 1
    assert n_checked > 50


def _parse_rows(text):
    lines = text.splitlines()
    return lines[0], [l.split("\t") for l in lines[1:]]


def _compare_rows(got_text, want_text):
    gh, g = _parse_rows(got_text)
    wh, w = _parse_rows(want_text)
    assert gh == wh
    assert len(g) == len(w)
    for i, (a, b) in enumerate(zip(g, w)):
        assert a[:5] == b[:5], (i, a[:5], b[:5])
        assert a[9:] == b[9:], (i, a[9:], b[9:])           # motif, period, ref_len, frequencies, dosage r2 columns
        if b[5] == "nan":
            assert a[5:9] == b[5:9], i
            continue
        assert abs(float(a[5]) - float(b[5])) <= 1.01e-2 * float(b[5]), (i, a[5], b[5])     # printed with 3 digits
        for c in (6, 7, 8):
            assert_close(float(a[c]), float(b[c]), "row %d col %d" % (i, c), rel=1e-6)


def _assoc_args(out, vcf, traits, **kw):
    ns = argparse.Namespace(outfile=out, tr_vcf=vcf, phenotype_name="test_pheno", traits=traits, vcftype=None,
                            same_samples=True, sample_list=None, region=None, non_major_cutoff=0, beagle_dosages=True,
                            plotting_phenotype=None, paired_genotype_plot=False, plot_phenotype_residuals=False,
                            plotting_ci_alphas=[], imputed_ukb_strs_paper_period_check=False, block_size=100)
    for k, v in kw.items():
        assert hasattr(ns, k), k
        setattr(ns, k, v)
    return ns


@pytest.mark.parametrize("key,vcf,traits,kw", [
    ("single_dosages", "many_samples_biallelic_dosages.vcf.gz", ["traits_0.npy"], {}),
    ("single_40_dosages", "many_samples_biallelic_dosages.vcf.gz", ["traits_0.npy"], {"sample_list": "samples_6_to_45.txt"}),
    ("combined_dosages_cutoff_20", "many_samples_biallelic_dosages.vcf.gz", ["traits_0.npy", "traits_1.npy"], {"non_major_cutoff": 20}),
    ("multi_dosages", "many_samples_multiallelic_dosages.vcf.gz", ["traits_0.npy"], {}),
    ("multi_dosages_cutoff_10", "many_samples_multiallelic_dosages.vcf.gz", ["traits_0.npy"], {"non_major_cutoff": 10}),
    ("multi_dosages_cutoff_20", "many_samples_multiallelic_dosages.vcf.gz", ["traits_0.npy"], {"non_major_cutoff": 20}),
    ("multi_dosages_cutoff_38", "many_samples_multiallelic_dosages.vcf.gz", ["traits_0.npy"], {"non_major_cutoff": 38}),
])
def test_associatr_beagle_dosages_cli_matches_reference(golden_dir, data_dir, tmp_path, key, vcf, traits, kw):
    """associaTR.main --beagle-dosages on the reference's plink2-pinned dosage fixtures == the unmodified reference's TSV."""
    from trtools_b200 import associaTR
    want = json.load(open(os.path.join(golden_dir, "associatr_dosage.json")))[key]
    kw = dict(kw)
    if "sample_list" in kw:
        kw["sample_list"] = os.path.join(data_dir, kw["sample_list"])
    out = str(tmp_path / "assoc.tsv")
    with contextlib.redirect_stdout(io.StringIO()):
        associaTR.main(_assoc_args(out, os.path.join(data_dir, vcf), [os.path.join(data_dir, t) for t in traits], **kw))
    _compare_rows(open(out).read(), want)


@pytest.mark.parametrize("key,cutoff,use_mask", [("assoc_dosage", 5, False), ("assoc_dosage_subset", 20, True)])
def test_assoc_dosage_kernel_vs_reference_and_oracle_on_synthetic(golden_dir, ctx, key, cutoff, use_mask):
    """C-ABI level on the synthetic AP block: the text columns equal the reference's rows; p / beta / se / R^2, the
    class sums and the r2 inputs within 1e-6 of the oracle's full-precision restatement."""
    from oracle import assoc as oassoc, dosage as odos, trh as otrh
    from oracle.records import LocusAsVariant
    from trtools_b200 import associaTR, block
    loci, extra = _load(golden_dir, "dosage_synth")
    good = [loci[j] for j in extra["good_index"]]
    traits = np.array(extra["traits"], dtype=float)
    S = good[0].gt.shape[0]
    mask = np.array(extra["sample_mask"], dtype=bool) if use_mask else None
    design = oassoc.prepare_design([traits], S, mask)
    blk = block.build_block(ctx, "hipstr", [LocusAsVariant(l) for l in good])
    ctx.assoc_set_design(design.covars, design.outcome, np.nonzero(design.sample_filter)[0].astype(np.int32))
    blk.ensure_ap()
    meta = associaTR.dosage_classes(blk)
    res = ctx.assoc_dosage_ols(*meta)
    buf = io.StringIO()
    associaTR._write_dosage_block(buf, blk, res, meta, design.pheno_std, cutoff, np.asarray(design.sample_filter, dtype=bool))
    want_lines = extra[key].splitlines()[1:]
    got_lines = buf.getvalue().splitlines()
    assert len(got_lines) == len(want_lines) == len(good)
    n_ok = 0
    for i, l in enumerate(good):
        g, w = got_lines[i].split("\t"), want_lines[i].split("\t")
        assert g[:5] == w[:5] and g[9:] == w[9:], (i, g, w)
        h = otrh.harmonize(l)
        loaded = odos.load_dosage_locus(l, h, design.sample_filter.copy(), cutoff)
        row = odos.regress_dosage_locus(loaded, design)
        assert int(res["n_tested"][i]) == row.n_samples_tested
        # class sums against the oracle's per-length dosage arrays
        sl = blk.allele_slice(i)
        cls, len_round, _ = meta
        for j in range(sl.stop - sl.start):
            if cls[sl.start + j] != j or loaded.gts is None:
                continue
            d = loaded.gts[float(len_round[sl.start + j])]
            assert_close(res["class_stats"][sl.start + j, 0], float(d.sum()), "sum d locus %d class %d" % (i, j), rel=1e-9, abs_tol=1e-9)
            assert_close(res["class_stats"][sl.start + j, 1], float((d * d).sum()), "sum d^2 locus %d class %d" % (i, j), rel=1e-9, abs_tol=1e-9)
        if row.locus_filtered:
            assert w[5] == "nan"
            continue
        what = "{} locus {}".format(key, i)
        assert_close(res["p"][i], row.p, what + " p", abs_tol=1e-300)
        assert_close(res["coef"][i] * design.pheno_std, row.coef, what + " coef", rel=1e-6, abs_tol=1e-12)
        assert_close(res["se"][i] * design.pheno_std, row.se, what + " se")
        assert_close(res["r2"][i], row.r2, what + " r2", rel=1e-6, abs_tol=1e-12)
        n_ok += 1
    assert n_ok > 5


def test_load_trs_dosage_generator_protocol(golden_dir, data_dir):
    """The per-locus generator of the drop-in load_trs (beagle_dosages=True) yields what the reference's consumer
    expects: detail names first, then 7-tuples whose dosage dicts and detail strings equal the reference's TSV."""
    from trtools_b200 import load_and_filter_genotypes as lafg
    want = json.load(open(os.path.join(golden_dir, "associatr_dosage.json")))["multi_dosages"]
    rows = [l.split("\t") for l in want.splitlines()[1:]]
    it = lafg.load_trs(os.path.join(data_dir, "many_samples_multiallelic_dosages.vcf.gz"), slice(None), None, 0, True, None)
    fields = next(it)
    assert fields[-2:] == ['dosage_estimated_r2_per_length_allele', 'r2_length_dosages_vs_best_guess_lengths']
    n = 0
    for (gts, uniq, chrom, pos, called, reason, details), w in zip(it, rows):
        assert [str(chrom), str(pos)] == w[:2]
        assert list(details) == w[9:]
        assert int(np.sum(called)) == int(w[3])
        if not reason:
            assert isinstance(gts, dict) and all(v.shape == (int(w[3]), 2) for v in gts.values())
        n += 1
    assert n == len(rows)
