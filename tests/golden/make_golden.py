#!/usr/bin/env python3
"""
Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference
(/root/reference, TRTools v6.1.0 @ f8ef1e9) in THIS container behind the shims of
oracle/shims (cyvcf2/statsmodels/matplotlib/pysam are not installed here).

    python tests/golden/make_golden.py            # rewrites every fixture

The reference cannot travel to the GPU box, so what is committed is
  * small INPUT fixtures (copies of a few reference *data* files: VCFs and trait arrays),
  * the reference's OUTPUTS on them (statSTR .tab text, dumpSTR samplog/loclog text and
    per-locus FILTER columns, associaTR TSV text), and
  * function-level outputs at full float precision (``*.npz``) for the 1e-6 comparisons,
    including outputs on synthetic blocks from ``trtools_b200.synth`` pushed through the
    reference via ``oracle.records.LocusAsVariant``.
Nothing here is imported by the product package.
"""
import argparse
import contextlib
import io
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle import ref_import  # noqa: E402

assert ref_import.enable(), "the reference tree is required to (re)generate goldens"

import cyvcf2  # noqa: E402  (shim or real)
import trtools.utils.tr_harmonizer as trh  # noqa: E402
import trtools.utils.utils as rutils  # noqa: E402
import trtools.statSTR.statSTR as rstat  # noqa: E402
import trtools.dumpSTR.dumpSTR as rdump  # noqa: E402
import trtools.dumpSTR.filters as rfilters  # noqa: E402
import trtools.associaTR.associaTR as rassoc  # noqa: E402
import trtools.associaTR.load_and_filter_genotypes as rlafg  # noqa: E402

from oracle.records import Locus, LocusAsVariant, locus_from_variant, save_loci, synth_to_loci  # noqa: E402
from trtools_b200 import synth  # noqa: E402

REF = ref_import.REFERENCE_ROOT
SV = os.path.join(REF, "trtools", "testsupport", "sample_vcfs")
EX = os.path.join(REF, "example-files")
DATA = os.path.join(HERE, "data")

NUMERIC_FMT = {"DP", "DSTUTTER", "DFLANKINDEL", "Q", "QEXP", "ML", "SR", "FR", "LC", "AD", "REPCN",
               "ADFL", "ADIR", "ADSP"}


def f(x):
    """JSON-able float (NaN/inf survive python's json)."""
    if x is None:
        return None
    return float(x)


def harm_dict(tr):
    return dict(ref_allele=tr.ref_allele, alt_alleles=list(tr.alt_alleles), motif=tr.motif,
                record_id=tr.record_id, pos=int(tr.pos), end_pos=int(tr.end_pos),
                ref_len=f(tr.ref_allele_length), alt_lens=[f(x) for x in tr.alt_allele_lengths],
                has_full=tr.HasFullStringGenotypes(), quality_field=tr.quality_field,
                full_ref=None if tr.full_alleles is None else tr.full_alleles[0])


def stat_dict(tr, groups, uselength):
    """Full-precision statSTR values through the reference's own wrappers."""
    d = {}
    d["thresh"] = [f(x) for x in rstat.GetThresh(tr, groups)]
    d["afreq"] = rstat.GetAFreq(tr, groups, uselength=uselength)
    d["acount"] = rstat.GetAFreq(tr, groups, uselength=uselength, count=True)
    d["nalleles"] = [int(x) for x in rstat.GetNAlleles(tr, nalleles_thresh=0.01, sample_indexes=groups, uselength=uselength)]
    d["hwep"] = [f(x) for x in rstat.GetHWEP(tr, groups, uselength=uselength)]
    d["het"] = [f(x) for x in rstat.GetHet(tr, groups, uselength=uselength)]
    d["entropy"] = [f(x) for x in rstat.GetEntropy(tr, groups, uselength=uselength)]
    d["mean"] = [f(x) for x in rstat.GetMean(tr, groups)]
    d["mode"] = [f(x) for x in rstat.GetMode(tr, groups)]
    d["var"] = [f(x) for x in rstat.GetVariance(tr, groups)]
    d["numcalled"] = [int(x) for x in rstat.GetNumSamples(tr, groups)]
    return d


def count_dict(tr):
    ac = tr.GetAlleleCounts(index=True)
    gc = tr.GetGenotypeCounts(index=True)
    return dict(ac_idx={str(int(k)): int(v) for k, v in ac.items()},
                gc_idx=[[int(x) for x in g] + [int(c)] for g, c in gc.items()],
                callrate=f(tr.GetCallRate()) if tr.GetCalledSamples() is not None else None,
                n_called=int(np.sum(tr.GetCalledSamples())),
                n_called_nonstrict=int(np.sum(tr.GetCalledSamples(strict=False))))


def statstr_cli(vcf, vcftype="auto", samples=None, sample_prefixes=None, use_length=False,
                precision=3, stats=None, only_passing=False):
    ns = argparse.Namespace(vcf=vcf, out=None, vcftype=vcftype, samples=samples,
                            sample_prefixes=sample_prefixes, region=None, precision=precision,
                            nalleles_thresh=0.01, plot_afreq=False, use_length=use_length,
                            only_passing=only_passing)
    allstats = ["thresh", "afreq", "acount", "nalleles", "hwep", "het", "entropy", "mean", "mode", "var",
                "numcalled"]
    for s in allstats:
        setattr(ns, s, stats is None or s in stats)
    with tempfile.TemporaryDirectory() as td:
        ns.out = os.path.join(td, "o")
        with contextlib.redirect_stdout(io.StringIO()):
            assert rstat.main(ns) == 0
        return open(ns.out + ".tab").read()


def dump_args(**kw):
    # rebuild the reference tests' default namespace by hand (same defaults, test_dumpSTR.py:16-69)
    ns = argparse.Namespace(
        vcf=None, vcftype="auto", out=None, zip=False, min_locus_callrate=None, min_locus_hwep=None,
        min_locus_het=None, max_locus_het=None, use_length=False, filter_regions=None,
        filter_regions_names=None, filter_hrun=False, drop_filtered=False, hipstr_min_call_DP=None,
        hipstr_max_call_DP=None, hipstr_min_call_Q=None, hipstr_max_call_flank_indel=None,
        hipstr_max_call_stutter=None, hipstr_min_supp_reads=None, longtr_min_call_DP=None,
        longtr_max_call_DP=None, longtr_min_call_Q=None, longtr_max_call_flank_indel=None,
        longtr_min_supp_reads=None, gangstr_expansion_prob_het=None, gangstr_expansion_prob_hom=None,
        gangstr_expansion_prob_total=None, gangstr_filter_span_only=False,
        gangstr_filter_spanbound_only=False, gangstr_filter_badCI=None, gangstr_min_call_DP=None,
        gangstr_max_call_DP=None, gangstr_min_call_Q=None, advntr_min_call_DP=None,
        advntr_max_call_DP=None, advntr_min_spanning=None, advntr_min_flanking=None, advntr_min_ML=None,
        eh_min_ADFL=None, eh_min_ADIR=None, eh_min_ADSP=None, eh_min_call_LC=None, eh_max_call_LC=None,
        popstr_min_call_DP=None, popstr_max_call_DP=None, popstr_require_support=None, num_records=None,
        die_on_warning=False, verbose=False)
    for k, v in kw.items():
        assert hasattr(ns, k), k
        setattr(ns, k, v)
    return ns


def dumpstr_run(source, vcftype, capture_calls=0, **kw):
    """Drive the reference's ApplyCallFilters/ApplyLocusFilters/INFO recompute exactly as
    dumpSTR.main does (dumpSTR.py:1238-1338) over ``source`` = VCF path or list of Locus.
    Returns the two logs + the per-locus FILTER column / INFO values (+ per-call FILTER strings
    and masked genotypes for the first ``capture_calls`` loci)."""
    args = dump_args(**kw)
    if isinstance(source, str):
        invcf = cyvcf2.VCF(source)
        samples = invcf.samples
        records = iter(invcf)
    else:
        samples = ["S%06d" % i for i in range(source[0].gt.shape[0])]
        records = (LocusAsVariant(l) for l in source)
    locus_filters = rdump.BuildLocusFilters(args)
    call_filters = rdump.BuildCallFilters(args)
    import collections
    sample_info = collections.OrderedDict()
    sample_info['numcalls'] = np.zeros((len(samples)), dtype=int)
    sample_info['totaldp'] = np.zeros((len(samples)), dtype=float)
    for name in rdump.GetAllCallFilters(call_filters):
        sample_info[name] = np.zeros((len(samples)), dtype=int)
    loc_info = collections.OrderedDict()
    loc_info["totalcalls"] = 0
    loc_info["PASS"] = 0
    loc_info["NO_CALLS_REMAINING"] = 0
    for filt in locus_filters:
        loc_info[filt.filter_name()] = 0
    per_locus = []
    calls = []
    for i, rec in enumerate(records):
        tr = trh.HarmonizeRecord(vcftype, rec)
        tr = rdump.ApplyCallFilters(tr, call_filters, sample_info, np.array(samples))
        rdump.ApplyLocusFilters(tr, locus_filters, loc_info, False)
        out = dict(filter=tr.vcfrecord.FILTER if isinstance(tr.vcfrecord.FILTER, str) else None)
        if isinstance(rec, LocusAsVariant):
            out["filter"] = rec.FILTER
        else:
            out["filter"] = rec._filter_raw
        hrun = rutils.GetHomopolymerRun(tr.full_alleles[0] if tr.HasFullStringGenotypes() else tr.ref_allele)
        out["HRUN"] = int(hrun)
        if np.sum(tr.GetCalledSamples()) > 0:
            af = tr.GetAlleleFreqs(uselength=args.use_length)
            gc = tr.GetGenotypeCounts(uselength=args.use_length)
            out["HET"] = f(rutils.GetHeterozygosity(af))
            out["HWEP"] = f(rutils.GetHardyWeinbergBinomialTest(af, gc))
            ac = tr.GetAlleleCounts(index=True)
            n_alleles = len(tr.alt_alleles) + 1
            out["AC"] = [int(ac.get(k, 0)) for k in range(1, n_alleles)]
            out["REFAC"] = int(ac.get(0, 0))
        else:
            out["HET"] = -1
            out["HWEP"] = -1
            out["AC"] = [0] * len(tr.alt_alleles)
            out["REFAC"] = 0
        per_locus.append(out)
        if i < capture_calls:
            calls.append(dict(filter_text=[str(x) for x in tr.vcfrecord.format('FILTER')],
                              gt=tr.GetGenotypeIndicies().tolist()))
    with tempfile.TemporaryDirectory() as td:
        rdump.WriteSampLog(sample_info, samples, os.path.join(td, "s"))
        rdump.WriteLocLog(loc_info, os.path.join(td, "l"))
        samplog = open(os.path.join(td, "s")).read()
        loclog = open(os.path.join(td, "l")).read()
    return dict(args={k: v for k, v in kw.items()}, samplog=samplog, loclog=loclog,
                per_locus=per_locus, calls=calls,
                call_filter_names=[c.name for c in call_filters],
                locus_filter_names=[l.filter_name() for l in locus_filters])


def associatr_cli(vcf, traits, **kw):
    ns = argparse.Namespace(outfile=None, tr_vcf=vcf, phenotype_name="test_pheno", traits=traits,
                            vcftype=None, same_samples=True, sample_list=None, region=None,
                            non_major_cutoff=0, beagle_dosages=False, plotting_phenotype=None,
                            paired_genotype_plot=False, plot_phenotype_residuals=False,
                            plotting_ci_alphas=[], imputed_ukb_strs_paper_period_check=False)
    for k, v in kw.items():
        assert hasattr(ns, k), k
        setattr(ns, k, v)
    with tempfile.TemporaryDirectory() as td:
        ns.outfile = os.path.join(td, "o.tsv")
        with contextlib.redirect_stdout(io.StringIO()):
            rassoc.main(ns)
        return open(ns.outfile).read()


def associatr_loci(loci, trait_arrays, non_major_cutoff=20, sample_mask=None):
    """Reference perform_gwas_helper over in-memory loci via the injected genotype iterator
    (associaTR.py:205-209, 246; load_and_filter_genotypes.py:137-259 re-driven per Locus)."""
    samples = ["%d" % i for i in range(loci[0].gt.shape[0])]

    def get_genotype_iter(sample_filter):
        yield ['motif', 'period', 'ref_len', 'allele_frequency']
        for l in loci:
            tr = trh.HarmonizeRecord(l.vcftype, LocusAsVariant(l))
            called = tr.GetCalledSamples()
            called_filter = called[sample_filter]
            curr = sample_filter & called
            n_samples = int(np.sum(curr))
            len_alleles = [tr.ref_allele_length] + tr.alt_allele_lengths
            len_alleles = [round(x, 2) for x in len_alleles]
            gts = tr.GetLengthGenotypes()[curr, :-1]
            af = rlafg.clean_len_alleles(tr.GetAlleleFreqs(curr))
            details = [tr.motif, str(len(tr.motif)), str(round(tr.ref_allele_length, 2)),
                       rlafg.dict_str({k: '{:.2g}'.format(v) for k, v in af.items()})]
            if len(af) == 0:
                reason = 'No called samples'
            elif len(af) == 1:
                reason = 'Only one called allele'
            else:
                a = list(af.values())
                a.pop(np.argmax(a))
                reason = ('non-major allele count<{}'.format(non_major_cutoff)
                          if np.sum(a) * n_samples * 2 < non_major_cutoff else None)
            yield (None if reason else gts, np.unique(len_alleles), tr.chrom, tr.pos, called_filter,
                   reason, details)

    with tempfile.TemporaryDirectory() as td:
        fns = []
        for i, a in enumerate(trait_arrays):
            fn = os.path.join(td, "t%d.npy" % i)
            np.save(fn, a)
            fns.append(fn)
        sample_fname = None
        if sample_mask is not None:
            sample_fname = os.path.join(td, "samples.txt")
            with open(sample_fname, "w") as fh:
                fh.write("\n".join(s for s, m in zip(samples, sample_mask) if m) + "\n")
        out = io.StringIO()
        with contextlib.redirect_stdout(io.StringIO()):
            rassoc.perform_gwas_helper(out, samples, get_genotype_iter, "test_pheno", fns, True,
                                       sample_fname, False, None, False, False, [])
        return out.getvalue()


def loci_from_vcf(path, vcftype, limit=None):
    vcf = cyvcf2.VCF(path)
    loci = []
    for i, rec in enumerate(vcf):
        if limit is not None and i >= limit:
            break
        loci.append(locus_from_variant(rec, vcftype, numeric_fmt=NUMERIC_FMT))
    return loci, vcf.samples


def function_level(loci, groups=(None,), do_stats=True):
    out = []
    for l in loci:
        try:
            tr = trh.HarmonizeRecord(l.vcftype, LocusAsVariant(l))
        except (TypeError, ValueError) as e:
            out.append(dict(error=type(e).__name__))
            continue
        d = dict(harm=harm_dict(tr))
        if l.gt is not None and do_stats:
            d["counts"] = count_dict(tr)
            if tr.GetMaxPloidy() >= 2:
                d["stats_len"] = stat_dict(tr, list(groups), True)
                d["stats_seq"] = stat_dict(tr, list(groups), False)
        out.append(d)
    return out


def copy_data(src, name=None):
    os.makedirs(DATA, exist_ok=True)
    dst = os.path.join(DATA, name or os.path.basename(src))
    shutil.copyfile(src, dst)
    os.chmod(dst, 0o644)
    return dst


INFO_KEYS = {"START", "END", "PERIOD", "RU", "VID", "VARID", "RL", "Motif", "IMP", "REF"}


def main():
    import warnings
    warnings.simplefilter("ignore")
    os.makedirs(DATA, exist_ok=True)
    only = set(sys.argv[1:])          # e.g. `make_golden.py edge synth` regenerates just those

    def want(name):
        return not only or name in only
    if want("many"):
        section_many()
    if want("trio"):
        section_trio()
    if want("callers"):
        section_callers()
    if want("assoc"):
        section_assoc()
    if want("synth"):
        section_synth()
    if want("edge"):
        section_edge()
    if want("dosage"):
        section_dosage()
    if want("reduce"):
        section_reduce()


def section_many():

    # ---- 1. HipSTR, 1874 loci x 50 samples (the reference's statSTR golden input) -------------
    many = copy_data(os.path.join(SV, "many_samples.vcf.gz"))
    g1 = copy_data(os.path.join(SV, "many_samples_subsample1.txt"))
    g2 = copy_data(os.path.join(SV, "many_samples_subsample2.txt"))
    loci, samples = loci_from_vcf(many, "hipstr")
    all_samples = np.array(samples)
    masks = [np.isin(all_samples, np.array([x.strip() for x in open(p)])) for p in (g1, g2)]
    extra = dict(
        ref=function_level(loci, groups=[None] + masks),
        group_masks=[m.astype(int).tolist() for m in masks],
        tab_all=statstr_cli(many, precision=4),
        tab_all_uselength=statstr_cli(many, precision=4, use_length=True),
        tab_strat=statstr_cli(many, precision=4, samples=g1 + "," + g2, sample_prefixes="1,2"),
        dump_numeric=dumpstr_run(many, "hipstr", capture_calls=40, hipstr_min_call_DP=20,
                                 hipstr_max_call_DP=150, hipstr_min_call_Q=0.9,
                                 hipstr_max_call_flank_indel=0.15, hipstr_max_call_stutter=0.15,
                                 min_locus_callrate=0.8, min_locus_hwep=0.01, min_locus_het=0.05,
                                 max_locus_het=0.9, filter_hrun=True),
        dump_uselength=dumpstr_run(many, "hipstr", hipstr_min_call_DP=10, min_locus_hwep=0.0001,
                                   min_locus_het=0.1, use_length=True),
    )
    save_loci(os.path.join(HERE, "hipstr_many.npz"), loci, extra, info_keys=INFO_KEYS, sample_names=samples)
    print("hipstr_many", len(loci))



def section_trio():
    # ---- 2. HipSTR trio chr21 (BASELINE config 1 input; dumpSTR golden input) ---------------
    trio = copy_data(os.path.join(EX, "trio_chr21_hipstr.sorted.vcf.gz"))
    loci, samples = loci_from_vcf(trio, "hipstr")
    extra = dict(
        ref=function_level(loci),
        tab_c1=statstr_cli(trio, vcftype="hipstr", stats=["afreq", "mean"]),
        dump_hipstr_filters=dumpstr_run(
            trio, "hipstr", capture_calls=200, filter_hrun=True, use_length=True, max_locus_het=0.45,
            min_locus_het=0.05, min_locus_hwep=0.5, hipstr_max_call_flank_indel=0.05,
            hipstr_max_call_stutter=0.3, hipstr_min_supp_reads=10, hipstr_min_call_DP=30,
            hipstr_max_call_DP=200, hipstr_min_call_Q=0.9),
        dump_numeric=dumpstr_run(
            trio, "hipstr", capture_calls=200, filter_hrun=True, use_length=True, max_locus_het=0.45,
            min_locus_het=0.05, min_locus_hwep=0.5, hipstr_max_call_flank_indel=0.05,
            hipstr_max_call_stutter=0.3, hipstr_min_call_DP=30, hipstr_max_call_DP=200,
            hipstr_min_call_Q=0.9),
        dump_locus_only=dumpstr_run(trio, "hipstr", min_locus_callrate=0.5, min_locus_hwep=0.5,
                                    min_locus_het=0.05, max_locus_het=0.45),
        golden_hipstr_filters_samplog=open(os.path.join(SV, "dumpSTR_vcfs", "hipstr_filters.samplog.tab")).read(),
        golden_hipstr_filters_loclog=open(os.path.join(SV, "dumpSTR_vcfs", "hipstr_filters.loclog.tab")).read(),
    )
    save_loci(os.path.join(HERE, "hipstr_trio.npz"), loci, extra, info_keys=INFO_KEYS, sample_names=samples)
    print("hipstr_trio", len(loci))



def section_callers():
    # ---- 3. the other callers (harmonisation + stats) ------------------------------------------
    for name, path, vt, dump_kw in [
        ("gangstr", os.path.join(SV, "test_gangstr.vcf"), "gangstr",
         dict(gangstr_min_call_DP=10, gangstr_max_call_DP=100, gangstr_min_call_Q=0.9,
              gangstr_expansion_prob_het=0.001, gangstr_expansion_prob_hom=0.0005,
              gangstr_expansion_prob_total=0.001, min_locus_callrate=0.6)),
        ("popstr", os.path.join(SV, "test_popstr.vcf"), "popstr",
         dict(popstr_min_call_DP=30, popstr_max_call_DP=200, use_length=True)),
        ("eh", os.path.join(SV, "test_ExpansionHunter.vcf"), "eh", dict(eh_min_call_LC=20)),
        ("advntr", os.path.join(SV, "test_advntr.vcf"), "advntr",
         dict(advntr_min_call_DP=50, advntr_max_call_DP=2000, advntr_min_spanning=1,
              advntr_min_flanking=20, advntr_min_ML=0.95)),
        ("longtr", os.path.join(SV, "test_longtr.vcf"), "longtr",
         dict(longtr_min_call_DP=30, longtr_max_call_DP=200, longtr_min_call_Q=0.9,
              longtr_max_call_flank_indel=0.05, filter_hrun=True, use_length=True,
              min_locus_het=0.05)),
    ]:
        limit = 400 if name == "gangstr" else None
        src = path
        if limit is not None:
            # keep the fixture small: header + first `limit` records
            os.makedirs(DATA, exist_ok=True)
            src = os.path.join(DATA, "test_%s_head.vcf" % name)
            with open(path) as fi, open(src, "w") as fo:
                n = 0
                for line in fi:
                    if not line.startswith("#"):
                        n += 1
                        if n > limit:
                            break
                    fo.write(line)
        else:
            src = copy_data(path)
        loci, samples = loci_from_vcf(src, vt)
        extra = dict(ref=function_level(loci),
                     tab=statstr_cli(src, vcftype=vt, precision=4),
                     tab_uselength=statstr_cli(src, vcftype=vt, precision=4, use_length=True),
                     dump=dumpstr_run(src, vt, capture_calls=50, **dump_kw))
        save_loci(os.path.join(HERE, "%s.npz" % name), loci, extra, info_keys=INFO_KEYS, sample_names=samples)
        print(name, len(loci))



def section_assoc():
    # ---- 4. associaTR fixtures (plink2-pinned in the reference's own tests) -------------------
    AS = os.path.join(SV, "associaTR")
    bi = copy_data(os.path.join(AS, "many_samples_biallelic.vcf.gz"))
    multi = copy_data(os.path.join(AS, "many_samples_multiallelic.vcf.gz"))
    t0 = copy_data(os.path.join(AS, "traits_0.npy"))
    t1 = copy_data(os.path.join(AS, "traits_1.npy"))
    s640 = copy_data(os.path.join(AS, "samples_6_to_45.txt"))
    for fn in ("single.plink2.trait_0.glm.linear", "combined.plink2.trait_0.glm.linear",
               "single_40.plink2.trait_0.glm.linear", "single_cutoff_5.plink2.trait_0.glm.linear"):
        copy_data(os.path.join(AS, fn))
    assoc = dict(
        single=associatr_cli(bi, [t0]),
        combined=associatr_cli(bi, [t0, t1]),
        single_40=associatr_cli(bi, [t0], sample_list=s640),
        cutoff_5=associatr_cli(bi, [t0], non_major_cutoff=5),
        cutoff_20=associatr_cli(bi, [t0, t1], non_major_cutoff=20),
        multi=associatr_cli(multi, [t0]),
        multi_cutoff=associatr_cli(multi, [t0, t1], non_major_cutoff=6),
    )
    with open(os.path.join(HERE, "associatr.json"), "w") as fh:
        json.dump(assoc, fh)
    print("associatr", {k: v.count("\n") for k, v in assoc.items()})



def section_dosage():
    # ---- 7. Beagle allele-probability dosages: GetDosages + the --beagle-dosages branch ----
    AS = os.path.join(SV, "associaTR")
    bi_d = copy_data(os.path.join(AS, "many_samples_biallelic_dosages.vcf.gz"))
    multi_d = copy_data(os.path.join(AS, "many_samples_multiallelic_dosages.vcf.gz"))
    t0 = copy_data(os.path.join(AS, "traits_0.npy"))
    t1 = copy_data(os.path.join(AS, "traits_1.npy"))
    s640 = copy_data(os.path.join(AS, "samples_6_to_45.txt"))
    for fn in ("single_dosages.plink2.trait_0.glm.linear", "single_40_dosages.plink2.trait_0.glm.linear"):
        copy_data(os.path.join(AS, fn))
    out = dict(
        single_dosages=associatr_cli(bi_d, [t0], beagle_dosages=True),
        single_40_dosages=associatr_cli(bi_d, [t0], sample_list=s640, beagle_dosages=True),
        combined_dosages_cutoff_20=associatr_cli(bi_d, [t0, t1], beagle_dosages=True, non_major_cutoff=20),
        multi_dosages=associatr_cli(multi_d, [t0], beagle_dosages=True),
        multi_dosages_cutoff_10=associatr_cli(multi_d, [t0], beagle_dosages=True, non_major_cutoff=10),
        multi_dosages_cutoff_20=associatr_cli(multi_d, [t0], beagle_dosages=True, non_major_cutoff=20),
        multi_dosages_cutoff_38=associatr_cli(multi_d, [t0], beagle_dosages=True, non_major_cutoff=38),
    )
    with open(os.path.join(HERE, "associatr_dosage.json"), "w") as fh:
        json.dump(out, fh)
    print("associatr dosage", {k: v.count("\n") for k, v in out.items()})

    # function level: GetDosages on real records and on synthetic AP blocks (incl. invalid AP fields)
    def dos(tr, kind, strict):
        try:
            d = tr.GetDosages(getattr(trh.TRDosageTypes, kind), strict=strict)
            return dict(values=[f(x) for x in d])
        except ValueError as e:
            return dict(error=str(e))

    def pack(loci):
        rows = []
        for l in loci:
            tr = trh.HarmonizeRecord(l.vcftype, LocusAsVariant(l))
            rows.append({k + ("" if strict else "_lenient"): dos(tr, k, strict)
                         for k in ("bestguess", "bestguess_norm", "beagleap", "beagleap_norm") for strict in (True, False)})
        return rows

    vcf = cyvcf2.VCF(multi_d)
    loci = [locus_from_variant(rec, "hipstr", numeric_fmt={"AP1", "AP2", "DP", "Q"}) for rec in vcf]
    vcf = cyvcf2.VCF(bi_d)
    for i, rec in enumerate(vcf):
        if i >= 40:
            break
        loci.append(locus_from_variant(rec, "hipstr", numeric_fmt={"AP1", "AP2", "DP", "Q"}))
    save_loci(os.path.join(HERE, "dosage_real.npz"), loci, extra=dict(dosages=pack(loci)), info_keys=INFO_KEYS)
    # synthetic: HipSTR loci with Beagle-like AP fields consistent with (a noisy version of) the hard calls
    rng = np.random.default_rng(20261017)
    sl = synth.make_loci(24, seed=31)
    calls = synth.fill_calls(sl, 301)
    sloci = synth_to_loci(sl, calls, with_fmt=False)
    for j, l in enumerate(sloci):
        S, nalt = l.gt.shape[0], len(l.alts)
        for p in (1, 2):
            ap = np.zeros((S, nalt), dtype=np.float32)
            hap = l.gt[:, p - 1].astype(int)
            for s_ in range(S):
                w = rng.dirichlet(np.full(nalt + 1, 0.3))
                a = hap[s_] if hap[s_] >= 0 else int(rng.integers(0, nalt + 1))
                w = 0.15 * w
                w[a] += 0.85
                ap[s_] = np.round(w[1:], 2).astype(np.float32)
            l.fmt["AP%d" % p] = ap
        if j == 5:
            l.fmt["AP1"][7, 0] = np.float32(1.5)         # sums to more than 1.1
        if j == 9:
            l.fmt["AP2"][3, 0] = np.float32(-0.25)       # negative
        if j == 11:
            del l.fmt["AP2"]                             # field missing
    traits = synth.make_traits(sl, calls.gt[0], 301, n_covars=3)
    mask = rng.random(301) < 0.8
    extra = dict(dosages=pack(sloci), traits=traits.tolist(), sample_mask=mask.tolist())
    good = [l for j, l in enumerate(sloci) if j not in (11,)]
    extra["good_index"] = [j for j in range(len(sloci)) if j not in (11,)]
    extra["assoc_dosage"] = associatr_dosage_loci(good, [traits], non_major_cutoff=5)
    extra["assoc_dosage_subset"] = associatr_dosage_loci(good, [traits], non_major_cutoff=20, sample_mask=mask)
    save_loci(os.path.join(HERE, "dosage_synth.npz"), sloci, extra=extra, info_keys=INFO_KEYS)
    print("dosage fixtures", len(loci), len(sloci))


def section_reduce():
    # ---- 8. qcSTR / compareSTR reductions ---------------------------------------------------------
    import trtools.compareSTR.compareSTR as rcmp
    import trtools.utils.mergeutils as rmerge  # noqa: F401  (imported by compareSTR; keeps the shims honest)
    # qcSTR: the record loop of qcSTR.main (qcSTR.py:523-570) on the first 150 records of many_samples.vcf.gz
    many = os.path.join(DATA, "many_samples.vcf.gz")
    vcf = cyvcf2.VCF(many)
    rng = np.random.default_rng(5)
    sample_index = rng.random(len(vcf.samples)) < 0.7
    out = {"sample_index": sample_index.tolist(), "runs": {}}
    recs = [rec for _, rec in zip(range(150), vcf)]
    for ignore in (False, True):
        sample_calls = np.zeros(int(sample_index.sum()))
        per_sample_total_qual = np.zeros(int(sample_index.sum()))
        locus_calls, per_locus = [], []
        for rec in recs:
            tr = trh.HarmonizeRecord("hipstr", rec)
            idx_gts = tr.GetGenotypeIndicies()[sample_index, :-1]
            nocall = np.full((1, idx_gts.shape[1]), -1)
            calls = ~np.all(idx_gts == nocall, axis=1)
            sample_calls += calls
            locus_calls.append(int(np.sum(calls)))
            q = tr.GetQualityScores()[sample_index, :]
            q[~calls] = np.nan
            if not ignore:
                q[np.isnan(q)] = 0
                per_sample_total_qual += q.reshape(-1)
                per_locus.append(f(np.mean(q)))
            else:
                qi = ~np.isnan(q)
                per_sample_total_qual[qi.reshape(-1)] += q[qi].reshape(-1)
                per_locus.append(f(np.mean(q[qi])))
        out["runs"]["ignore" if ignore else "zero"] = dict(sample_calls=sample_calls.tolist(), locus_calls=locus_calls,
                                                            per_sample_total_qual=[f(x) for x in per_sample_total_qual],
                                                            per_locus=per_locus)
    # compareSTR: the unmodified UpdateComparisonResults on the reference's own pair of GangSTR call sets
    CS = os.path.join(SV, "compareSTR_vcfs")
    f1 = copy_data(os.path.join(CS, "test_gangstr1.vcf.gz"))
    f2 = copy_data(os.path.join(CS, "test_gangstr2.vcf.gz"))
    v1, v2 = cyvcf2.VCF(f1), cyvcf2.VCF(f2)
    shared = [s for s in v1.samples if s in v2.samples]
    idxs = [np.array([v1.samples.index(s) for s in shared]), np.array([v2.samples.index(s) for s in shared])]
    r2 = {(r.CHROM, r.POS): r for r in v2}
    pairs = []
    for r in v1:
        if (r.CHROM, r.POS) in r2:
            pairs.append((r, r2[(r.CHROM, r.POS)]))
    cmp_out = {"shared": shared, "n_pairs": len(pairs), "positions": [[p[0].CHROM, int(p[0].POS)] for p in pairs], "runs": {}}
    for ignore_phasing in (False, True):
        overall = {"ALL": rcmp.NewOverallPeriod([], [])}
        locus = {"chrom": [], "start": [], "numcalls": [], "metric-conc-seq": [], "metric-conc-len": []}
        sample = {"numcalls": np.zeros(len(shared)), "conc-seq-count": np.zeros(len(shared)), "conc-len-count": np.zeros(len(shared))}
        errors = []
        for a, b in pairs:
            t1, t2 = trh.HarmonizeRecord("gangstr", a), trh.HarmonizeRecord("gangstr", b)
            try:
                rcmp.UpdateComparisonResults(t1, t2, idxs, ignore_phasing, False, [], [], 0, overall, locus, sample, None)
            except ValueError as e:
                errors.append([a.CHROM, int(a.POS), str(e)])
        cmp_out["runs"]["ignore_phasing" if ignore_phasing else "phased"] = dict(
            overall={k: f(v) for k, v in overall["ALL"]["ALL"].items()},
            locus={k: [f(x) if not isinstance(x, str) else x for x in v] for k, v in locus.items()},
            sample={k: v.tolist() for k, v in sample.items()}, errors=errors)
    out["compare"] = cmp_out
    with open(os.path.join(HERE, "reductions.json"), "w") as fh:
        json.dump(out, fh)
    print("reductions", len(recs), cmp_out["n_pairs"], {k: len(v["errors"]) for k, v in cmp_out["runs"].items()})


def associatr_dosage_loci(loci, trait_arrays, non_major_cutoff=20, sample_mask=None):
    """Unmodified perform_gwas_helper over the unmodified load_trs (beagle_dosages=True), the records coming from
    the shim's in-memory VCF source."""
    samples = ["%d" % i for i in range(loci[0].gt.shape[0])]
    key = "golden_dosage_%d" % id(loci)
    cyvcf2.register_memory_vcf(key, [LocusAsVariant(l) for l in loci],
                               "##fileformat=VCFv4.1\n##command=HipSTR-synthetic\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\n", samples)

    def get_genotype_iter(sample_filter):
        return rlafg.load_trs("mem://" + key, sample_filter, None, non_major_cutoff, True, "hipstr")

    with tempfile.TemporaryDirectory() as td:
        fns = []
        for i, arr in enumerate(trait_arrays):
            fn = os.path.join(td, "t%d.npy" % i)
            np.save(fn, arr)
            fns.append(fn)
        sample_fname = None
        if sample_mask is not None:
            sample_fname = os.path.join(td, "samples.txt")
            with open(sample_fname, "w") as fh:
                fh.write("\n".join(s for s, m in zip(samples, sample_mask) if m) + "\n")
        out = io.StringIO()
        with contextlib.redirect_stdout(io.StringIO()):
            rassoc.perform_gwas_helper(out, samples, get_genotype_iter, "test_pheno", fns, True, sample_fname, True,
                                       None, False, False, [])
        return out.getvalue()


def section_synth():
    # ---- 5. synthetic blocks pushed through the reference ---------------------------------------
    for name, L, S, seed in [("synth_small", 96, 257, 20261017), ("synth_wide", 12, 4103, 7)]:
        sl = synth.make_loci(L, seed=seed)
        calls = synth.fill_calls(sl, S)
        loci = synth_to_loci(sl, calls)
        rng = np.random.default_rng(seed)
        masks = [rng.random(S) < 0.5, rng.random(S) < 0.1]
        traits = synth.make_traits(sl, calls.gt[0], S, n_covars=10, seed=seed)
        sample_mask = rng.random(S) < 0.9
        extra = dict(
            ref=function_level(loci, groups=[None] + masks),
            group_masks=[m.astype(int).tolist() for m in masks],
            dump=dumpstr_run(loci, "hipstr", capture_calls=L, hipstr_min_call_DP=20,
                             hipstr_max_call_flank_indel=0.15, min_locus_hwep=0.0001),
            dump_all=dumpstr_run(loci, "hipstr", capture_calls=8, hipstr_min_call_DP=20,
                                 hipstr_max_call_DP=60, hipstr_min_call_Q=0.95,
                                 hipstr_max_call_flank_indel=0.15, hipstr_max_call_stutter=0.2,
                                 min_locus_callrate=0.97, min_locus_hwep=0.01, min_locus_het=0.2,
                                 max_locus_het=0.8, filter_hrun=True, use_length=True),
            assoc=associatr_loci(loci, [traits], non_major_cutoff=20),
            assoc_subset=associatr_loci(loci, [traits], non_major_cutoff=5, sample_mask=sample_mask),
            traits=traits.tolist(), sample_mask=sample_mask.astype(int).tolist(),
            synth=dict(L=L, S=S, seed=seed),
        )
        save_loci(os.path.join(HERE, "%s.npz" % name), loci, extra, info_keys=INFO_KEYS)
        print(name, L, S)



def section_edge():
    # ---- 6. hand-made edge cases ------------------------------------------------------------------
    def hip(ref, alts, gt, start_off=0, end_trim=0, period=2, pos=100):
        return Locus("hipstr", "1", pos, ref, alts,
                     {"START": pos + start_off, "END": pos + len(ref) - 1 - end_trim, "PERIOD": period},
                     None if gt is None else np.array(gt, dtype=np.int16))
    edge = [
        hip("ACACACAC", [], [[0, 0, 1], [0, 0, 0]]),                                # single allele
        hip("ACACACAC", ["ACACAC", "ACACACACAC"], [[-1, -2, 0]] * 4),               # all missing
        hip("ACACACAC", ["ACACAC"], [[0, 1, 1], [1, -1, 1], [-1, 1, 0], [1, 1, 0]]),  # half calls
        hip("ACACACAC", ["ACACAC"], [[0, -2, 0], [1, 1, 0], [0, 1, 1]]),             # haploid pad -> HWE nan
        hip("ACACACAC", ["ACACAC", "ACAC"], [[0, 1, 2, 1], [1, 1, 0, 0], [2, 2, -2, 1], [0, -1, 1, 0]]),  # triploid
        hip("TACACACACG", ["TACACACG", "GACACACACG", "TACACACACC"],
            [[0, 1, 1], [2, 3, 1], [0, 2, 0], [3, 3, 1], [1, 1, 0]], start_off=1, end_trim=1),  # flank dups
        hip("AAAAAGAAAAAG", ["AAAAAG"], [[0, 1, 0], [0, 0, 0]], period=6),           # hrun
        hip("ACGACGACGAC", ["ACGACGAC", "ACGACGACGACGA"], [[0, 1, 0], [2, 2, 1], [0, 2, 1]], period=3),  # fractional
        hip("acacacac", ["acacac"], [[0, 1, 0], [1, 1, 1]]),                         # lower case
        hip("AC", ["ACAC"], [[0, 1, 0]], period=5),                                  # period > len -> NNNNN motif
        hip("ACACACAC", ["ACACAC"], None),                                           # no samples
        Locus("gangstr", "2", 500, "ACACAC", ["ACAC", "ACACACAC"], {"RU": "ac"},
              np.array([[0, 1, 0], [2, 2, 0], [1, 2, 0]], dtype=np.int16)),
        Locus("eh", "3", 900, "A", ["<STR5>", "<STR12>"], {"VARID": "v1", "RU": "CAG", "RL": 30, "REF": 10},
              np.array([[1, 2, 0], [0, 1, 0], [2, 2, 0]], dtype=np.int16)),
        Locus("popstr", "4", 1200, "ACACACAC", ["<3.5>", "<6>"], {"Motif": "AC"},
              np.array([[0, 1, 0], [1, 2, 0], [2, 2, 0]], dtype=np.int16), record_id="chr4:1200:M"),
        Locus("advntr", "5", 77, "GGTGGT", ["GGT", "GGTGGTGGT"], {"RU": "GGT", "VID": "vid7"},
              np.array([[0, 2, 0], [1, 1, 0], [-1, -1, 0]], dtype=np.int16)),
        Locus("hipstr", "6", 10, "ACAC", [], {"START": 10, "END": 13}, np.array([[0, 0, 0]], dtype=np.int16)),  # missing PERIOD
    ]
    save_loci(os.path.join(HERE, "edge.npz"), edge, dict(ref=function_level(edge)), info_keys=INFO_KEYS)
    print("edge", len(edge))


if __name__ == "__main__":
    main()
