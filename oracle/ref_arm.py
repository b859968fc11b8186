"""
TEST INFRASTRUCTURE ONLY — the CPU arm of bench.py.

Runs the UNMODIFIED reference (gymrek-lab/TRTools: ``/root/reference`` in the build container, its byte-for-byte
copy ``baseline/_ref`` on the GPU box; see baseline/install_ref.py and oracle/ref_import.py) over synthetic
cyvcf2-layout records of the bench workload, one process per host core on disjoint loci, for the three tools of the
metric:

* statSTR   — the loop body of ``statSTR.main`` (trtools/statSTR/statSTR.py:576-630) with all 11 statistics;
* dumpSTR   — ``BuildCallFilters`` / ``BuildLocusFilters`` / ``ApplyCallFilters`` / ``ApplyLocusFilters`` and the INFO
              recompute exactly as ``dumpSTR.main`` strings them together (trtools/dumpSTR/dumpSTR.py:1238-1336);
* associaTR — ``perform_gwas_helper`` over ``load_trs`` (trtools/associaTR/associaTR.py:114-304,
              load_and_filter_genotypes.py:47-259), the records coming from the cyvcf2 shim's ``mem://`` source.

Every worker returns its wall time AND the values it computed, so that bench.py can compare the GPU rows of the same
loci with them in the same run (the ``parity`` object of the bench line).  When the reference is not importable the
oracle port (``kind = "port"``) is timed instead.  VCF text parsing is excluded on both arms.
"""
import argparse
import collections
import contextlib
import io
import os
import tempfile
import time

import numpy as np

from . import ref_import
from .records import LocusAsVariant, synth_to_loci

ALL_STATS = ("thresh", "afreq", "acount", "nalleles", "hwep", "het", "entropy", "mean", "mode", "var", "numcalled")
C3_FLAGS = dict(hipstr_min_call_DP=20, hipstr_max_call_flank_indel=0.15, min_locus_hwep=1e-4)   # BASELINE configs[2]
C4_CUTOFF = 20                                                                                  # BASELINE configs[3]

_LOCI = None          # SynthLoci of the whole workload, set in the parent before the pool forks


def set_loci(sl):
    global _LOCI
    _LOCI = sl


def reference_kind() -> str:
    return "reference" if ref_import.reference_code_available() else "port"


def _sub_loci(lo, hi, S, with_fmt):
    from trtools_b200 import synth
    sl = _LOCI
    calls = synth.fill_calls(sl, S, slice(lo, hi))
    sub = synth.SynthLoci(seed=sl.seed, n_loci=hi - lo, chrom=sl.chrom[lo:hi], pos=sl.pos[lo:hi], start=sl.start[lo:hi],
                          end=sl.end[lo:hi], period=sl.period[lo:hi], ref=sl.ref[lo:hi], alts=sl.alts[lo:hi],
                          n_alleles=sl.n_alleles[lo:hi], cum_freq=sl.cum_freq[lo:hi], locus_offset=sl.locus_offset + lo)
    return synth_to_loci(sub, calls, with_fmt=with_fmt)


def bench_traits(S, seed, n_cov=10):
    """trait + 10 PC covariates of the bench workload, float64 [S, 11] (same array on both arms)."""
    rng = np.random.default_rng(seed)
    return np.hstack([rng.standard_normal((S, 1)), rng.standard_normal((S, n_cov))])


# ---------------------------------------------------------------------------------------------------------------
# statSTR
# ---------------------------------------------------------------------------------------------------------------
def statstr_worker(job):
    lo, hi, S, kind = job
    loci = _sub_loci(lo, hi, S, False)
    rows = []
    if kind == "reference":
        ref_import.enable()
        import trtools.utils.tr_harmonizer as rtrh
        import trtools.statSTR.statSTR as rstat
        recs = [LocusAsVariant(l) for l in loci]
        out = io.StringIO()
        fmt = "\t{:.3}"
        t0 = time.perf_counter()
        for rec in recs:                                            # statSTR.py:576-630, every statistic switched on
            tr = rtrh.HarmonizeRecord(rtrh.VcfTypes.hipstr, rec)
            out.write(str(rec.CHROM) + "\t" + str(rec.POS) + "\t" + str(rec.POS + len(tr.ref_allele)))
            v = {}
            v["thresh"] = rstat.GetThresh(tr, sample_indexes=[None])
            for val in v["thresh"]: out.write(rstat.format_nan_precision(fmt, val))
            v["afreq"] = rstat.GetAFreq(tr, sample_indexes=[None], uselength=False)
            for val in v["afreq"]: out.write("\t" + str(val))
            v["acount"] = rstat.GetAFreq(tr, sample_indexes=[None], uselength=False, count=True)
            for val in v["acount"]: out.write("\t" + str(val))
            v["nalleles"] = rstat.GetNAlleles(tr, nalleles_thresh=0.01, sample_indexes=[None], uselength=False)
            for val in v["nalleles"]: out.write("\t" + str(val))
            for key, fn in (("hwep", rstat.GetHWEP), ("het", rstat.GetHet), ("entropy", rstat.GetEntropy)):
                v[key] = fn(tr, sample_indexes=[None], uselength=False)
                for val in v[key]: out.write(rstat.format_nan_precision(fmt, val))
            for key, fn in (("mean", rstat.GetMean), ("mode", rstat.GetMode), ("var", rstat.GetVariance)):
                v[key] = fn(tr, sample_indexes=[None])
                for val in v[key]: out.write(rstat.format_nan_precision(fmt, val))
            v["numcalled"] = rstat.GetNumSamples(tr, sample_indexes=[None])
            for val in v["numcalled"]: out.write("\t" + str(val))
            out.write("\n")
            rows.append((tr, v))
        dt = time.perf_counter() - t0
        res = []
        for tr, v in rows:                                          # untimed: full-precision values for the parity check
            ac = tr.GetAlleleCounts(index=True)
            r = {k: float(v[k][0]) for k in ("thresh", "hwep", "het", "entropy", "mean", "mode", "var")}
            r["nalleles"] = int(v["nalleles"][0])
            r["numcalled"] = int(v["numcalled"][0])
            r["ac"] = [int(ac.get(i, 0)) for i in range(len(tr.alt_alleles) + 1)]
            res.append(r)
        return dt, len(res), res
    from . import stats as ostats, trh as otrh
    t0 = time.perf_counter()
    res = []
    for l in loci:
        h = otrh.harmonize(l)
        vals = ostats.locus_stats(h, l.gt, ALL_STATS, [None], uselength=False)
        ostats.format_row(l.chrom, l.pos, h, vals)
        ac = otrh.allele_counts(h, l.gt, index=True)
        r = {k: float(vals[k][0]) for k in ("thresh", "hwep", "het", "entropy", "mean", "mode", "var")}
        r["nalleles"] = int(vals["nalleles"][0])
        r["numcalled"] = int(vals["numcalled"][0])
        r["ac"] = [int(ac.get(i, 0)) for i in range(len(l.alts) + 1)]
        res.append(r)
    return time.perf_counter() - t0, len(res), res


# ---------------------------------------------------------------------------------------------------------------
# dumpSTR
# ---------------------------------------------------------------------------------------------------------------
def _dump_args():
    ns = argparse.Namespace(
        vcf=None, vcftype="hipstr", out=None, zip=False, min_locus_callrate=None, min_locus_hwep=None,
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
    for k, v in C3_FLAGS.items():
        setattr(ns, k, v)
    return ns


def dumpstr_worker(job):
    """-> seconds, n_loci, dict(per_locus=[...], numcalls, totaldp, counts{name: int[S]}, filter_names)."""
    lo, hi, S, kind = job
    loci = _sub_loci(lo, hi, S, True)
    for l in loci:                                   # the C3 filters read DP and DFLANKINDEL only
        l.fmt = {k: l.fmt[k] for k in ("DP", "DFLANKINDEL")}
    per_locus = []
    if kind == "reference":
        ref_import.enable()
        import trtools.utils.tr_harmonizer as rtrh
        import trtools.utils.utils as rutils
        import trtools.dumpSTR.dumpSTR as rdump
        args = _dump_args()
        samples = np.array(["S%06d" % i for i in range(S)])
        recs = [LocusAsVariant(l) for l in loci]
        t0 = time.perf_counter()
        locus_filters = rdump.BuildLocusFilters(args)
        call_filters = rdump.BuildCallFilters(args)
        sample_info = collections.OrderedDict()
        sample_info['numcalls'] = np.zeros((S,), dtype=int)
        sample_info['totaldp'] = np.zeros((S,), dtype=float)
        names = list(rdump.GetAllCallFilters(call_filters))
        for name in names:
            sample_info[name] = np.zeros((S,), dtype=int)
        loc_info = collections.OrderedDict([("totalcalls", 0), ("PASS", 0), ("NO_CALLS_REMAINING", 0)])
        for filt in locus_filters:
            loc_info[filt.filter_name()] = 0
        for rec in recs:                                            # dumpSTR.py:1271-1336
            tr = rtrh.HarmonizeRecord("hipstr", rec)
            tr = rdump.ApplyCallFilters(tr, call_filters, sample_info, samples)
            rdump.ApplyLocusFilters(tr, locus_filters, loc_info, False)
            out = dict(filter=rec.FILTER)
            out["HRUN"] = int(rutils.GetHomopolymerRun(tr.full_alleles[0] if tr.HasFullStringGenotypes() else tr.ref_allele))
            n_called = int(np.sum(tr.GetCalledSamples()))
            if n_called > 0:
                af = tr.GetAlleleFreqs(uselength=args.use_length)
                gc = tr.GetGenotypeCounts(uselength=args.use_length)
                out["HET"] = float(rutils.GetHeterozygosity(af))
                out["HWEP"] = float(rutils.GetHardyWeinbergBinomialTest(af, gc))
                ac = tr.GetAlleleCounts(index=True)
                out["AC"] = [int(ac.get(k, 0)) for k in range(len(tr.alt_alleles) + 1)]
            else:
                out["HET"] = out["HWEP"] = -1.0
                out["AC"] = [0] * (len(tr.alt_alleles) + 1)
            out["n_called"] = n_called
            per_locus.append(out)
        dt = time.perf_counter() - t0
        return dt, len(recs), dict(per_locus=per_locus, numcalls=sample_info["numcalls"], totaldp=sample_info["totaldp"],
                                   counts={n: sample_info[n] for n in names}, filter_names=names,
                                   locus_filter_names=[f.filter_name() for f in locus_filters])
    from . import dumpstr as od, trh as otrh
    cf = [od.min_value("HipSTRCallMinDepth", "DP", C3_FLAGS["hipstr_min_call_DP"]),
          od.hipstr_flank_indels(C3_FLAGS["hipstr_max_call_flank_indel"])]
    lf = [od.LocusFilter("hwe", C3_FLAGS["min_locus_hwep"], False)]
    sinfo, linfo = od.new_sample_info(S, cf), od.new_loc_info(lf)
    t0 = time.perf_counter()
    for l in loci:
        h = otrh.harmonize(l)
        r = od.apply_call_filters(l, cf, sinfo)
        _, text = od.apply_locus_filters(l, h, r.gt, lf, linfo)
        info = od.recompute_info(h, r.gt, False)
        per_locus.append(dict(filter=text, HRUN=info["HRUN"], HET=float(info["HET"]), HWEP=float(info["HWEP"]),
                              AC=[info["REFAC"]] + list(info["AC"]), n_called=int(np.sum(otrh.called_samples(r.gt)))))
    dt = time.perf_counter() - t0
    names = [f.name for f in cf]
    return dt, len(loci), dict(per_locus=per_locus, numcalls=sinfo["numcalls"], totaldp=sinfo["totaldp"],
                               counts={n: sinfo[n] for n in names}, filter_names=names,
                               locus_filter_names=[f.filter_name() for f in lf])


# ---------------------------------------------------------------------------------------------------------------
# associaTR
# ---------------------------------------------------------------------------------------------------------------
_HIPSTR_HEADER = "##fileformat=VCFv4.1\n##command=HipSTR-synthetic --trtools-b200\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\n"


def assoc_worker(job):
    """-> seconds, n_loci, list of dict(n_tested, filtered, p_text, coef, se, r2) parsed from the reference's TSV."""
    lo, hi, S, kind, traits_path = job
    loci = _sub_loci(lo, hi, S, False)
    rows = []
    if kind == "reference":
        ref_import.enable()
        import cyvcf2
        import trtools.associaTR.associaTR as rassoc
        import trtools.associaTR.load_and_filter_genotypes as rlafg
        samples = ["%d" % i for i in range(S)]
        key = "bench_%d_%d" % (lo, os.getpid())
        cyvcf2.register_memory_vcf(key, [LocusAsVariant(l) for l in loci], _HIPSTR_HEADER, samples)

        def get_genotype_iter(sample_filter):                       # the injection point of associaTR.py:443-446
            return rlafg.load_trs("mem://" + key, sample_filter, None, C4_CUTOFF, False, "hipstr")

        out = io.StringIO()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            rassoc.perform_gwas_helper(out, samples, get_genotype_iter, "bench", [traits_path], True, None, False,
                                       None, False, False, [])
        dt = time.perf_counter() - t0
        for line in out.getvalue().splitlines()[1:]:
            c = line.split("\t")
            rows.append(dict(n_tested=int(c[3]), filtered=c[4], p_text=c[5], coef=float(c[6]), se=float(c[7]), r2=float(c[8])))
        return dt, len(rows), rows
    from . import assoc as oassoc, trh as otrh
    design = oassoc.prepare_design([np.load(traits_path)], S, None)
    t0 = time.perf_counter()
    for l in loci:
        h = otrh.harmonize(l)
        row = oassoc.regress_locus(oassoc.load_locus(l, h, design.sample_filter.copy(), C4_CUTOFF), design)
        row.to_text()
        rows.append(dict(n_tested=row.n_samples_tested, filtered=str(row.locus_filtered), p_text="{:.2e}".format(row.p),
                         coef=float(row.coef), se=float(row.se), r2=float(row.r2)))
    return time.perf_counter() - t0, len(rows), rows


# ---------------------------------------------------------------------------------------------------------------
# pool
# ---------------------------------------------------------------------------------------------------------------
def run_pool(worker, n_loci, S, cores, kind, extra=()):
    """One process per core on disjoint loci [0, n_loci).  Returns (loci/s, loci done, slowest worker's seconds,
    results in locus order).  loci/s extrapolates linearly to the whole workload because loci are independent."""
    import multiprocessing as mp
    cores = max(1, min(cores, n_loci))
    base, rem = divmod(n_loci, cores)
    jobs, lo = [], 0
    for i in range(cores):
        hi = lo + base + (1 if i < rem else 0)
        jobs.append((lo, hi, S, kind) + tuple(extra))
        lo = hi
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(worker, jobs)
    wall = max(r[0] for r in res)
    done = sum(r[1] for r in res)
    return done / wall, done, wall, [r[2] for r in res]


def save_traits(traits):
    fd, path = tempfile.mkstemp(prefix="trt_bench_traits_", suffix=".npy")
    os.close(fd)
    np.save(path, traits)
    return path
