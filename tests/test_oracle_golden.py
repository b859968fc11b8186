"""
CPU tests: pin the oracle (oracle/*.py, the numpy restatement of the reference's hot path)
against golden vectors produced by the UNMODIFIED reference (tests/golden/make_golden.py)
and against the reference's in-code known answers.
"""
import json
import os

import numpy as np
import pytest

from oracle import assoc as oassoc
from oracle import dumpstr as odump
from oracle import stats as ostats
from oracle import trh as otrh
from oracle.records import load_loci

from helpers import assert_close, assert_close_list

FIXTURES = ["hipstr_many", "hipstr_trio", "gangstr", "popstr", "eh", "advntr", "longtr", "synth_small",
            "synth_wide", "edge"]
_cache = {}


def fixture(golden_dir, name):
    if name not in _cache:
        _cache[name] = load_loci(os.path.join(golden_dir, name + ".npz"))
    return _cache[name]


def _groups(extra, n):
    groups = [None]
    for m in extra.get("group_masks", []):
        groups.append(np.array(m, dtype=bool))
    return groups


@pytest.mark.parametrize("name", FIXTURES)
def test_harmonize_matches_reference(golden_dir, name):
    loci, extra, _ = fixture(golden_dir, name)
    for i, (l, ref) in enumerate(zip(loci, extra["ref"])):
        if "error" in ref:
            with pytest.raises((TypeError, ValueError)) as e:
                otrh.harmonize(l)
            assert type(e.value).__name__ == ref["error"]
            continue
        h = otrh.harmonize(l)
        r = ref["harm"]
        assert h.ref_allele == r["ref_allele"], i
        assert h.alt_alleles == r["alt_alleles"], i
        assert h.motif == r["motif"], i
        assert h.record_id == r["record_id"], i
        assert h.pos == r["pos"] and h.end_pos == r["end_pos"], i
        assert h.ref_allele_length == r["ref_len"], i
        assert h.alt_allele_lengths == r["alt_lens"], i
        assert (h.full_alleles is not None) == r["has_full"], i
        assert h.quality_field == r["quality_field"], i


@pytest.mark.parametrize("name", FIXTURES)
def test_counts_and_stats_match_reference(golden_dir, name):
    loci, extra, _ = fixture(golden_dir, name)
    step = 7 if name in ("hipstr_trio",) else (3 if name == "hipstr_many" else 1)
    for i in range(0, len(loci), step):
        l, ref = loci[i], extra["ref"][i]
        if "error" in ref or l.gt is None or "counts" not in ref:
            continue
        h = otrh.harmonize(l)
        c = ref["counts"]
        ac = otrh.allele_counts(h, l.gt, index=True)
        assert {str(int(k)): int(v) for k, v in ac.items()} == c["ac_idx"], i
        gc = otrh.genotype_counts(h, l.gt, index=True)
        assert sorted([int(x) for x in g] + [int(n)] for g, n in gc.items()) == sorted(c["gc_idx"]), i
        assert int(np.sum(otrh.called_samples(l.gt))) == c["n_called"]
        assert int(np.sum(otrh.called_samples(l.gt, strict=False))) == c["n_called_nonstrict"]
        assert_close(otrh.call_rate(l.gt), c["callrate"], "callrate")
        if "stats_len" not in ref:
            continue
        groups = _groups(extra, l.gt.shape[0])
        for key, uselength in (("stats_len", True), ("stats_seq", False)):
            r = ref[key]
            got = ostats.locus_stats(h, l.gt, ostats.STAT_ORDER, groups, uselength=uselength)
            for stat in ("afreq", "acount", "nalleles", "numcalled"):
                assert [x for x in got[stat]] == r[stat], (i, key, stat)
            for stat in ("thresh", "hwep", "het", "entropy", "mean", "mode", "var"):
                assert_close_list(got[stat], r[stat], "{} {} {}".format(i, key, stat), rel=1e-12)


def test_statstr_rows_match_reference_tab(golden_dir):
    """Whole-file statSTR text (all 11 stats, precision 4) reproduced row by row."""
    loci, extra, _ = fixture(golden_dir, "hipstr_many")
    for key, uselength, groups in (("tab_all", False, [None]), ("tab_all_uselength", True, [None]),
                                   ("tab_strat", False, _groups(extra, 50)[1:])):
        lines = extra[key].splitlines()
        assert len(lines) == len(loci) + 1
        for i in range(0, len(loci), 5):
            l = loci[i]
            h = otrh.harmonize(l)
            vals = ostats.locus_stats(h, l.gt, ostats.STAT_ORDER, groups, uselength=uselength)
            assert ostats.format_row(l.chrom, l.pos, h, vals, precision=4) == lines[i + 1], (key, i)


def test_statstr_config1_rows(golden_dir):
    """BASELINE config 1: statSTR --afreq --mean --vcftype hipstr on the trio chr21 file."""
    loci, extra, _ = fixture(golden_dir, "hipstr_trio")
    lines = extra["tab_c1"].splitlines()
    assert lines[0] == "chrom\tstart\tend\tafreq\tmean"
    for i in range(0, len(loci), 3):
        l = loci[i]
        h = otrh.harmonize(l)
        vals = ostats.locus_stats(h, l.gt, ("afreq", "mean"), [None], uselength=False)
        assert ostats.format_row(l.chrom, l.pos, h, vals, precision=3) == lines[i + 1], i


def _build_filters(argd):
    cf, lf = [], []
    a = argd
    g = a.get
    if g("hipstr_max_call_flank_indel") is not None: cf.append(odump.hipstr_flank_indels(a["hipstr_max_call_flank_indel"]))
    if g("hipstr_max_call_stutter") is not None: cf.append(odump.hipstr_stutter(a["hipstr_max_call_stutter"]))
    if g("hipstr_min_call_DP") is not None: cf.append(odump.min_value("HipSTRCallMinDepth", "DP", a["hipstr_min_call_DP"]))
    if g("hipstr_max_call_DP") is not None: cf.append(odump.max_value("HipSTRCallMaxDepth", "DP", a["hipstr_max_call_DP"]))
    if g("hipstr_min_call_Q") is not None: cf.append(odump.min_value("HipSTRCallMinQ", "Q", a["hipstr_min_call_Q"]))
    if g("longtr_max_call_flank_indel") is not None: cf.append(odump.hipstr_flank_indels(a["longtr_max_call_flank_indel"], "LongTRCallFlankIndels"))
    if g("longtr_min_call_DP") is not None: cf.append(odump.min_value("LongTRCallMinDepth", "DP", a["longtr_min_call_DP"]))
    if g("longtr_max_call_DP") is not None: cf.append(odump.max_value("LongTRCallMaxDepth", "DP", a["longtr_max_call_DP"]))
    if g("longtr_min_call_Q") is not None: cf.append(odump.min_value("LongTRCallMinQ", "Q", a["longtr_min_call_Q"]))
    if g("gangstr_min_call_DP") is not None: cf.append(odump.min_value("GangSTRCallMinDepth", "DP", a["gangstr_min_call_DP"]))
    if g("gangstr_max_call_DP") is not None: cf.append(odump.max_value("GangSTRCallMaxDepth", "DP", a["gangstr_max_call_DP"]))
    if g("gangstr_min_call_Q") is not None: cf.append(odump.min_value("GangSTRCallMinQ", "Q", a["gangstr_min_call_Q"]))
    if g("gangstr_expansion_prob_het") is not None: cf.append(odump.gangstr_expansion("qexp_het", a["gangstr_expansion_prob_het"]))
    if g("gangstr_expansion_prob_hom") is not None: cf.append(odump.gangstr_expansion("qexp_hom", a["gangstr_expansion_prob_hom"]))
    if g("gangstr_expansion_prob_total") is not None: cf.append(odump.gangstr_expansion("qexp_total", a["gangstr_expansion_prob_total"]))
    if g("advntr_min_call_DP") is not None: cf.append(odump.min_value("AdVNTRCallMinDepth", "DP", a["advntr_min_call_DP"]))
    if g("advntr_max_call_DP") is not None: cf.append(odump.max_value("AdVNTRCallMaxDepth", "DP", a["advntr_max_call_DP"]))
    if g("advntr_min_spanning") is not None: cf.append(odump.min_value("AdVNTRCallMinSpanning", "SR", a["advntr_min_spanning"]))
    if g("advntr_min_flanking") is not None: cf.append(odump.min_value("AdVNTRCallMinFlanking", "FR", a["advntr_min_flanking"]))
    if g("advntr_min_ML") is not None: cf.append(odump.min_value("AdVNTRCallMinML", "ML", a["advntr_min_ML"]))
    if g("eh_min_call_LC") is not None: cf.append(odump.min_value("EHCallMinDepth", "LC", a["eh_min_call_LC"]))
    if g("popstr_min_call_DP") is not None: cf.append(odump.min_value("PopSTRMinCallDepth", "DP", a["popstr_min_call_DP"]))
    if g("popstr_max_call_DP") is not None: cf.append(odump.max_value("PopSTRMaxCallDepth", "DP", a["popstr_max_call_DP"]))
    ul = bool(g("use_length", False))
    if g("min_locus_callrate") is not None: lf.append(odump.LocusFilter("callrate", a["min_locus_callrate"]))
    if g("min_locus_hwep") is not None: lf.append(odump.LocusFilter("hwe", a["min_locus_hwep"], ul))
    if g("min_locus_het") is not None: lf.append(odump.LocusFilter("hetlow", a["min_locus_het"], ul))
    if g("max_locus_het") is not None: lf.append(odump.LocusFilter("hethigh", a["max_locus_het"], ul))
    if g("filter_hrun"): lf.append(odump.LocusFilter("hrun"))
    return cf, lf, ul


DUMP_CASES = [("hipstr_many", "dump_numeric"), ("hipstr_many", "dump_uselength"), ("hipstr_trio", "dump_numeric"),
              ("hipstr_trio", "dump_locus_only"), ("gangstr", "dump"), ("popstr", "dump"), ("eh", "dump"),
              ("advntr", "dump"), ("longtr", "dump"), ("synth_small", "dump"), ("synth_small", "dump_all"),
              ("synth_wide", "dump"), ("synth_wide", "dump_all")]


@pytest.mark.parametrize("name,key", DUMP_CASES)
def test_dumpstr_matches_reference(golden_dir, name, key):
    loci, extra, samples = fixture(golden_dir, name)
    ref = extra[key]
    cf, lf, ul = _build_filters(ref["args"])
    assert [c.name for c in cf] == ref["call_filter_names"]
    assert [x.filter_name() for x in lf] == ref["locus_filter_names"]
    n = loci[0].gt.shape[0]
    sinfo = odump.new_sample_info(n, cf)
    linfo = odump.new_loc_info(lf)
    for i, l in enumerate(loci):
        h = otrh.harmonize(l)
        res = odump.apply_call_filters(l, cf, sinfo)
        _, ftext = odump.apply_locus_filters(l, h, res.gt, lf, linfo)
        r = ref["per_locus"][i]
        assert ftext == r["filter"], (i, ftext, r["filter"])
        info = odump.recompute_info(h, res.gt, ul)
        assert info["HRUN"] == r["HRUN"] and info["AC"] == r["AC"] and info["REFAC"] == r["REFAC"], i
        assert_close(info["HET"], r["HET"], "HET %d" % i, rel=1e-12)
        assert_close(info["HWEP"], r["HWEP"], "HWEP %d" % i, rel=1e-12)
        if i < len(ref["calls"]):
            assert [str(x) for x in res.filter_text] == ref["calls"][i]["filter_text"], i
            assert res.gt.astype(int).tolist() == ref["calls"][i]["gt"], i
    names = samples if samples else ["S%06d" % i for i in range(n)]
    assert odump.samplog_text(sinfo, names) == ref["samplog"]
    assert odump.loclog_text(linfo) == ref["loclog"]


def test_dumpstr_reference_golden_logs_are_what_the_reference_produced(golden_dir):
    """The samplog/loclog the reference ships for hipstr_filters equal what it produced here."""
    _, extra, _ = fixture(golden_dir, "hipstr_trio")
    assert extra["dump_hipstr_filters"]["samplog"] == extra["golden_hipstr_filters_samplog"]
    assert extra["dump_hipstr_filters"]["loclog"] == extra["golden_hipstr_filters_loclog"]


@pytest.mark.parametrize("name", ["synth_small", "synth_wide"])
@pytest.mark.parametrize("key,cutoff,use_mask", [("assoc", 20, False), ("assoc_subset", 5, True)])
def test_associatr_matches_reference_on_synthetic(golden_dir, name, key, cutoff, use_mask):
    loci, extra, _ = fixture(golden_dir, name)
    traits = np.array(extra["traits"], dtype=float)
    n = loci[0].gt.shape[0]
    mask = np.array(extra["sample_mask"], dtype=bool) if use_mask else None
    design = oassoc.prepare_design([traits], n, mask)
    lines = extra[key].splitlines(keepends=True)
    assert lines[0] == oassoc.header_text("test_pheno")
    for i, l in enumerate(loci):
        h = otrh.harmonize(l)
        loaded = oassoc.load_locus(l, h, design.sample_filter.copy(), cutoff)
        row = oassoc.regress_locus(loaded, design)
        ref_cols = lines[i + 1].rstrip("\n").split("\t")
        got_cols = row.to_text().rstrip("\n").split("\t")
        assert got_cols[:5] == ref_cols[:5], i
        assert got_cols[9:] == ref_cols[9:], i
        assert got_cols[5] == ref_cols[5], (i, got_cols[5], ref_cols[5])     # p printed as %.2e
        for c in (6, 7, 8):
            assert_close(float(got_cols[c]), float(ref_cols[c]), "col %d locus %d" % (c, i), rel=1e-9)


def _read_vcf_loci(path, vcftype):
    from trtools_b200 import cyvcf2_compat
    from oracle.records import locus_from_variant
    vcf = cyvcf2_compat.VCF(path)
    return [locus_from_variant(r, vcftype, numeric_fmt=set()) for r in vcf], vcf.samples


@pytest.mark.parametrize("key,traits,kw", [
    ("single", ["traits_0.npy"], {}),
    ("combined", ["traits_0.npy", "traits_1.npy"], {}),
    ("cutoff_5", ["traits_0.npy"], {"cutoff": 5}),
    ("cutoff_20", ["traits_0.npy", "traits_1.npy"], {"cutoff": 20}),
    ("single_40", ["traits_0.npy"], {"sample_list": "samples_6_to_45.txt"}),
])
def test_associatr_matches_reference_on_plink_fixture(golden_dir, data_dir, key, traits, kw):
    ref_text = json.load(open(os.path.join(golden_dir, "associatr.json")))[key]
    loci, samples = _read_vcf_loci(os.path.join(data_dir, "many_samples_biallelic.vcf.gz"), "hipstr")
    arrays = [np.load(os.path.join(data_dir, t)) for t in traits]
    mask = None
    if "sample_list" in kw:
        keep = [x.strip() for x in open(os.path.join(data_dir, kw["sample_list"]))]
        mask = np.isin(np.array(samples), keep)
    design = oassoc.prepare_design(arrays, len(samples), mask)
    lines = ref_text.splitlines(keepends=True)
    assert len(lines) == len(loci) + 1
    for i, l in enumerate(loci):
        h = otrh.harmonize(l)
        loaded = oassoc.load_locus(l, h, design.sample_filter.copy(), kw.get("cutoff", 0))
        row = oassoc.regress_locus(loaded, design)
        ref_cols = lines[i + 1].rstrip("\n").split("\t")
        got_cols = row.to_text().rstrip("\n").split("\t")
        assert got_cols[:6] == ref_cols[:6], (i, got_cols[:6], ref_cols[:6])
        assert got_cols[9:] == ref_cols[9:], i
        for c in (6, 7, 8):
            assert_close(float(got_cols[c]), float(ref_cols[c]), "col %d locus %d" % (c, i), rel=1e-9)


# ---- the reference's own in-code known answers (trtools/utils/tests/test_utils.py:17-100) ----
def test_known_answers_utils():
    assert ostats.validate_allele_freqs({0: 1}) and ostats.validate_allele_freqs({0: 0.5, 1: 0.5})
    assert not ostats.validate_allele_freqs({}) and not ostats.validate_allele_freqs({0: 0.5})
    assert not ostats.validate_allele_freqs({-1: 1, 1: 1})
    afreqs = {0: 0.5, 1: 0.2, 2: 0.3}
    assert ostats.heterozygosity({0: 1}) == 0 and ostats.heterozygosity({0: 0.5, 1: 0.5}) == 0.5
    assert ostats.heterozygosity(afreqs) == 0.62
    assert np.isnan(ostats.heterozygosity({}))
    assert ostats.entropy({0: 1}) == 0 and ostats.entropy({0: 0.5, 1: 0.5}) == 1
    assert abs(ostats.entropy(afreqs) - 1.48) < .01
    assert ostats.mean({0: 1}) == 0 and ostats.mean({0: 0.5, 1: 0.5}) == 0.5 and ostats.mean(afreqs) == 0.8
    assert ostats.mode({0: 1}) == 0 and ostats.mode({0: 0.49, 1: 0.51}) == 1
    assert ostats.mode({0: 0.1, 1: 0.1, 2: 0.3, 3: 0.5}) == 3 and np.isnan(ostats.mode({}))
    assert ostats.variance({0: 1}) == 0 and ostats.variance({0: 0.5, 1: 0.5}) == 0.25
    assert np.isnan(ostats.variance({}))
    hw = ostats.hardy_weinberg_binomial_test
    assert round(hw(afreqs, {(0, 1): 10, (0, 0): 20, (1, 2): 5}), 2) == 0.02
    assert round(hw(afreqs, {(0, 0): 20}), 2) == 0.0
    assert round(hw(afreqs, {(0, 1): 20}), 2) == 0.0
    assert np.isnan(hw(afreqs, {(3, 3): 6})) and np.isnan(hw(afreqs, {(0, 3): 6}))
    assert np.isnan(hw({}, {(0, 3): 6}))
    assert otrh.infer_repeat_sequence('ATATATAT', 2) == 'AT'
    assert otrh.canonical_one_strand("CAG") == 'AGC'
    assert otrh.homopolymer_run("AATAAAATAAAAAT") == 5
    assert otrh.fabricate_allele("ACG", 2.4) == "ACGACGA"


# ---- Beagle allele-probability dosages -----------------------------------------------------------
def _dosage_fixture(golden_dir, name):
    loci, extra, _ = load_loci(os.path.join(golden_dir, name + ".npz"))
    for l in loci:
        for k in ("AP1", "AP2"):
            if k in l.fmt:
                l.fmt[k] = np.ascontiguousarray(l.fmt[k][:, :len(l.alts)], dtype=np.float32)
    return loci, extra


@pytest.mark.parametrize("name", ["dosage_real", "dosage_synth"])
def test_oracle_beagle_dosages_match_reference(golden_dir, name):
    """oracle.dosage.beagle_dosages == the unmodified reference's GetDosages(beagleap[_norm]) on real Beagle records and
    on synthetic AP blocks, including the three validation errors."""
    import numpy as np
    from oracle import dosage as odos, trh as otrh
    loci, extra = _dosage_fixture(golden_dir, name)
    n_err = 0
    for i, (l, w) in enumerate(zip(loci, extra["dosages"])):
        h = otrh.harmonize(l)
        for kind, norm in (("beagleap", False), ("beagleap_norm", True)):
            ww = w[kind]
            try:
                got = odos.beagle_dosages(h, l, norm=norm, strict=True)
            except odos.DosageError as e:
                assert ww.get("error") == str(e), (i, kind)
                n_err += 1
                assert np.isnan(odos.beagle_dosages(h, l, norm=norm, strict=False)).all()
                continue
            assert "values" in ww, (i, kind)
            assert np.array_equal(np.array(ww["values"], dtype=np.float32), got, equal_nan=True), (i, kind)
        want_bg = np.array(w["bestguess"]["values"], dtype=np.float32)
        assert np.array_equal(otrh.dosages_bestguess(h, l.gt), want_bg)
    assert n_err == (6 if name == "dosage_synth" else 0)


@pytest.mark.parametrize("key,cutoff,use_mask", [("assoc_dosage", 5, False), ("assoc_dosage_subset", 20, True)])
def test_oracle_assoc_dosage_rows_match_reference(golden_dir, key, cutoff, use_mask):
    """oracle load_dosage_locus + regress_dosage_locus reproduce the rows the unmodified perform_gwas_helper over the
    unmodified load_trs(beagle_dosages=True) wrote for the synthetic AP block (text columns exact, numbers to 1e-9)."""
    import numpy as np
    from oracle import assoc as oassoc, dosage as odos, trh as otrh
    loci, extra = _dosage_fixture(golden_dir, "dosage_synth")
    good = [loci[j] for j in extra["good_index"]]
    traits = np.array(extra["traits"], dtype=float)
    S = good[0].gt.shape[0]
    design = oassoc.prepare_design([traits], S, np.array(extra["sample_mask"], dtype=bool) if use_mask else None)
    want = extra[key].splitlines()[1:]
    for i, l in enumerate(good):
        h = otrh.harmonize(l)
        row = odos.regress_dosage_locus(odos.load_dosage_locus(l, h, design.sample_filter.copy(), cutoff), design)
        g, w = row.to_text().rstrip("\n").split("\t"), want[i].split("\t")
        assert g[:6] == w[:6] and g[9:] == w[9:], (i, g, w)
        if w[5] != "nan":
            for c in (6, 7, 8):
                assert abs(float(g[c]) - float(w[c])) <= 1e-9 * abs(float(w[c])), (i, c)


# ---- qcSTR / compareSTR reductions ---------------------------------------------------------------
def _records(path, vcftype, limit=None):
    from oracle.records import locus_from_variant
    from trtools_b200.cyvcf2_compat import TextVCF
    v = TextVCF(path)
    out = []
    for i, rec in enumerate(v):
        if limit is not None and i >= limit:
            break
        out.append(locus_from_variant(rec, vcftype, numeric_fmt={"Q", "DP"}))
    return out, v.samples


def test_oracle_qc_reductions_match_reference(golden_dir, data_dir):
    from oracle import reduce as ored
    want = json.load(open(os.path.join(golden_dir, "reductions.json")))
    loci, _ = _records(os.path.join(data_dir, "many_samples.vcf.gz"), "hipstr", 150)
    idx = np.array(want["sample_index"], dtype=bool)
    for key, ignore in (("zero", False), ("ignore", True)):
        got = ored.qc_reduce(loci, idx, "Q", ignore)
        w = want["runs"][key]
        assert got["sample_calls"].tolist() == w["sample_calls"]
        assert got["locus_calls"] == w["locus_calls"]
        assert_close_list(got["per_sample_total_qual"].tolist(), w["per_sample_total_qual"], "per-sample quality " + key, rel=1e-12)
        assert_close_list(got["per_locus"], w["per_locus"], "per-locus quality " + key, rel=1e-7)


@pytest.mark.parametrize("key,ignore_phasing", [("phased", False), ("ignore_phasing", True)])
def test_oracle_compare_reductions_match_reference(golden_dir, data_dir, key, ignore_phasing):
    from oracle import reduce as ored
    want = json.load(open(os.path.join(golden_dir, "reductions.json")))["compare"]
    l1, s1 = _records(os.path.join(data_dir, "test_gangstr1.vcf.gz"), "gangstr")
    l2, s2 = _records(os.path.join(data_dir, "test_gangstr2.vcf.gz"), "gangstr")
    by_pos = {(l.chrom, l.pos): l for l in l2}
    shared = want["shared"]
    idxs = [np.array([s1.index(s) for s in shared]), np.array([s2.index(s) for s in shared])]
    w = want["runs"][key]
    sample = {k: np.zeros(len(shared)) for k in ("numcalls", "conc-seq-count", "conc-len-count")}
    tot = np.zeros(5)
    n_seq = n_len = n_calls = 0
    rows = []
    for a in l1:
        b = by_pos.get((a.chrom, a.pos))
        if b is None:
            continue
        r = ored.compare_locus(a, otrh.harmonize(a), b, otrh.harmonize(b), idxs, ignore_phasing)
        if r is None:
            continue
        rows.append((r["numcalls"], float(np.sum(r["conc_seq"])) / r["numcalls"], float(np.sum(r["conc_len"])) / r["numcalls"]))
        sample["numcalls"] += r["both"]
        sample["conc-seq-count"][r["both"]] += r["conc_seq"]
        sample["conc-len-count"][r["both"]] += r["conc_len"]
        tot += r["sums"]
        n_calls += r["numcalls"]
        n_seq += int(np.sum(r["conc_seq"]))
        n_len += int(np.sum(r["conc_len"]))
    assert [x[0] for x in rows] == [int(x) for x in w["locus"]["numcalls"]]
    assert_close_list([x[1] for x in rows], w["locus"]["metric-conc-seq"], "conc-seq", rel=1e-12)
    assert_close_list([x[2] for x in rows], w["locus"]["metric-conc-len"], "conc-len", rel=1e-12)
    for k in sample:
        assert sample[k].tolist() == w["sample"][k], k
    o = w["overall"]
    assert (n_calls, n_seq, n_len) == (int(o["numcalls"]), int(o["conc_seq_count"]), int(o["conc_len_count"]))
    assert_close_list(tot.tolist(), [o["total_len_1"], o["total_len_2"], o["total_len_11"], o["total_len_12"], o["total_len_22"]],
                      "length sums", rel=1e-9)
