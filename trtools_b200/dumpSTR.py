"""
dumpSTR — call-level and locus-level filtering of TR VCFs (drop-in for reference
trtools/dumpSTR/dumpSTR.py): same flags, ``main(args) -> int``, output VCF, ``.samplog.tab`` and
``.loclog.tab``.

The reference's per-record loop (dumpSTR.py:1270-1338: ``ApplyCallFilters`` with numpy string ops,
``ApplyLocusFilters`` re-deriving allele counts per filter) is replaced by two GPU calls per block
of records: ``trt_call_filters`` (filter bitmask, masked genotypes, per-sample accumulators) and
``trt_locus_filters`` (scan of the masked genotypes + HET/HWEP/AC + filter flags).  Python keeps only
what is VCF text: FORMAT:FILTER strings for the filtered calls, nulling of their FORMAT fields, and
record emission.
"""
import argparse
import collections
import itertools
import os
import subprocess as sp
import sys
from typing import Dict, List

import numpy as np

from . import __version__, _lib, block as _block, common, dist as _dist
from . import filters as filters
from . import tr_harmonizer as trh
from . import utils

cyvcf2 = utils.cyvcf2
_NOCALL_INT_FORMAT_VAL = -2147483648


def MakeWriter(outfile, invcf, command):
    """reference dumpSTR.py:24-47."""
    invcf.add_to_header("##command-DumpSTR=" + command)
    return cyvcf2.Writer(outfile, invcf)


# ---- argument validation (reference dumpSTR.py:49-521) ---------------------------------------------------
def _range01(args, name, flag):
    v = getattr(args, name)
    if v is not None and (v < 0 or v > 1):
        common.WARNING("{} must be between 0 and 1".format(flag))
        return False
    return True


def _nonneg(args, name, flag):
    v = getattr(args, name)
    if v is not None and v < 0:
        common.WARNING("{} must be >= 0".format(flag))
        return False
    return True


def _ordered(args, lo, hi, flo, fhi):
    a, b = getattr(args, lo), getattr(args, hi)
    if a is not None and b is not None and b < a:
        common.WARNING("{} must be >= {}".format(fhi, flo))
        return False
    return True


def CheckLocusFilters(args, vcftype, is_beagle: bool):
    """reference dumpSTR.py:49-108."""
    if args.min_locus_callrate is not None and is_beagle:
        common.WARNING("--min-locus-callrate cannot be applied to Beagle imputed files at the moment "
                       "as there are currently no call level Beagle filters")
        return False
    for name, flag in (("min_locus_hwep", "--min-locus-hwep"), ("min_locus_het", "--min-locus-het"),
                       ("max_locus_het", "--max-locus-het")):
        v = getattr(args, name)
        if v is not None and (v < 0 or v > 1):
            common.WARNING("Invalid {}. Must be between 0 and 1".format(flag))
            return False
    if args.min_locus_het is not None and args.max_locus_het is not None and args.max_locus_het < args.min_locus_het:
        common.WARNING("Cannot have --max-locus-het less than --min-locus-het")
        return False
    seqcallers = [trh.VcfTypes["hipstr"], trh.VcfTypes["longtr"]]
    if args.use_length and vcftype not in seqcallers:
        common.WARNING("--use-length is only meaningful for HipSTR or LongTR, which report sequence level differences.")
    if args.filter_hrun and vcftype not in seqcallers:
        common.WARNING("--filter-hrun only relevant to HipSTR or LongTR files. This filter will have no effect.")
    if args.filter_regions is not None and args.filter_regions_names is not None:
        if len(args.filter_regions_names.split(",")) != len(args.filter_regions.split(",")):
            common.WARNING("Length of --filter-regions-names must match --filter-regions.")
            return False
    return True


def CheckHipSTRFilters(format_fields, args, prefix="hipstr"):
    """reference dumpSTR.py:110-161 (and :163-207 for the LongTR flavour)."""
    flag = "--" + prefix
    ok = _range01(args, prefix + "_max_call_flank_indel", flag + "-max-call-flank-indel")
    if prefix == "hipstr":
        ok = ok and _range01(args, "hipstr_max_call_stutter", "--hipstr-max-call-stutter")
    ok = ok and _nonneg(args, prefix + "_min_supp_reads", flag + "-min-supp-reads")
    ok = ok and _nonneg(args, prefix + "_min_call_DP", flag + "-min-call-DP")
    ok = ok and _nonneg(args, prefix + "_max_call_DP", flag + "-max-call-DP")
    ok = ok and _ordered(args, prefix + "_min_call_DP", prefix + "_max_call_DP", flag + "-min-call-DP",
                         flag + "-max-call-DP")
    ok = ok and _range01(args, prefix + "_min_call_Q", flag + "-min-call-Q")
    return ok


def CheckLongTRFilters(format_fields, args):
    return CheckHipSTRFilters(format_fields, args, prefix="longtr")


def CheckGangSTRFilters(format_fields, args):
    """reference dumpSTR.py:209-272."""
    ok = _nonneg(args, "gangstr_min_call_DP", "--gangstr-min-call-DP")
    ok = ok and _nonneg(args, "gangstr_max_call_DP", "--gangstr-max-call-DP")
    ok = ok and _ordered(args, "gangstr_min_call_DP", "gangstr_max_call_DP", "--gangstr-min-call-DP",
                         "--gangstr-max-call-DP")
    ok = ok and _range01(args, "gangstr_min_call_Q", "--gangstr-min-call-Q")
    for nm in ("het", "hom", "total"):
        ok = ok and _range01(args, "gangstr_expansion_prob_" + nm, "--gangstr-expansion-prob-" + nm)
    return ok


def CheckAdVNTRFilters(format_fields, args):
    """reference dumpSTR.py:274-322."""
    ok = _nonneg(args, "advntr_min_call_DP", "--advntr-min-call-DP")
    ok = ok and _nonneg(args, "advntr_max_call_DP", "--advntr-max-call-DP")
    ok = ok and _ordered(args, "advntr_min_call_DP", "advntr_max_call_DP", "--advntr-min-call-DP",
                         "--advntr-max-call-DP")
    ok = ok and _nonneg(args, "advntr_min_spanning", "--advntr-min-spanning")
    ok = ok and _nonneg(args, "advntr_min_flanking", "--advntr-min-flanking")
    ok = ok and _nonneg(args, "advntr_min_ML", "--advntr-min-ML")
    return ok


def CheckEHFilters(format_fields, args):  # pragma: no cover
    """reference dumpSTR.py:324-368."""
    ok = True
    for nm in ("eh_min_ADFL", "eh_min_ADIR", "eh_min_ADSP", "eh_min_call_LC", "eh_max_call_LC"):
        ok = ok and _nonneg(args, nm, "--" + nm.replace("_", "-", 2).replace("_", "-"))
    return ok and _ordered(args, "eh_min_call_LC", "eh_max_call_LC", "--eh-min-call-LC", "--eh-max-call-LC")


def CheckPopSTRFilters(format_fields, args):
    """reference dumpSTR.py:370-394."""
    ok = _nonneg(args, "popstr_min_call_DP", "--popstr-min-call-DP")
    ok = ok and _nonneg(args, "popstr_max_call_DP", "--popstr-max-call-DP")
    ok = ok and _ordered(args, "popstr_min_call_DP", "popstr_max_call_DP", "--popstr-min-call-DP",
                         "--popstr-max-call-DP")
    return ok and _nonneg(args, "popstr_require_support", "--popstr-require-support")


_TOOL_CHECKS = {"hipstr": CheckHipSTRFilters, "longtr": CheckLongTRFilters, "gangstr": CheckGangSTRFilters,
                "advntr": CheckAdVNTRFilters, "eh": CheckEHFilters, "popstr": CheckPopSTRFilters}


def CheckFilters(format_fields, args, vcftype, is_beagle):
    """reference dumpSTR.py:396-521: locus filters, then: call-level filter flags of another caller, or
    any call-level flag on a Beagle file, are errors; finally the caller's own range checks."""
    if not CheckLocusFilters(args, vcftype, is_beagle):
        return False
    set_prefixes = set()
    for name, value in vars(args).items():
        pre = name.split("_")[0]
        if pre in _TOOL_CHECKS and value is not None and value is not False:
            set_prefixes.add(pre)
    for pre in sorted(set_prefixes):
        if is_beagle:
            common.WARNING("{} call-level filters cannot be applied to Beagle imputed files".format(pre))
            return False
        if vcftype != trh.VcfTypes[pre]:
            common.WARNING("{} options can only be applied to {} VCFs".format(pre, pre))
            return False
    return _TOOL_CHECKS[vcftype.name](format_fields, args)


# ---- logs (reference dumpSTR.py:523-588) ---------------------------------------------------------------
def WriteLocLog(loc_info, fname):
    keys = list(loc_info.keys())
    assert "totalcalls" in keys and "PASS" in keys
    keys.remove("totalcalls")
    callrate = 0 if loc_info["PASS"] == 0 else float(loc_info["totalcalls"]) / loc_info["PASS"]
    with open(fname, "w") as f:
        f.write("MeanSamplesPerPassingSTR\t%s\n" % callrate)
        for k in keys:
            f.write("FILTER:%s\t%s\n" % (k, loc_info[k]))
    return True


def WriteSampLog(sample_info: Dict[str, np.ndarray], sample_names: List[str], fname: str):
    header = ["sample"]
    header.extend(sample_info.keys())
    header[header.index('totaldp')] = 'meanDP'
    with open(fname, "w") as f:
        f.write("\t".join(header) + "\n")
        for i, s in enumerate(sample_names):
            numcalls = sample_info["numcalls"][i]
            cols = [s, str(numcalls), str(sample_info["totaldp"][i] * 1.0 / numcalls) if numcalls > 0 else "0"]
            cols.extend(str(c[i]) for c in itertools.islice(sample_info.values(), 2, None))
            f.write("\t".join(cols) + "\n")


def GetAllCallFilters(call_filters):
    return [filt.name for filt in call_filters]


# ---- filter construction (reference dumpSTR.py:777-915) ------------------------------------------------
def BuildCallFilters(args):
    f = []
    a = args
    if a.hipstr_max_call_flank_indel is not None: f.append(filters.HipSTRCallFlankIndels(a.hipstr_max_call_flank_indel))
    if a.hipstr_max_call_stutter is not None: f.append(filters.HipSTRCallStutter(a.hipstr_max_call_stutter))
    if a.hipstr_min_supp_reads is not None: f.append(filters.HipSTRCallMinSuppReads(a.hipstr_min_supp_reads))
    if a.hipstr_min_call_DP is not None: f.append(filters.CallFilterMinValue("HipSTRCallMinDepth", "DP", a.hipstr_min_call_DP))
    if a.hipstr_max_call_DP is not None: f.append(filters.CallFilterMaxValue("HipSTRCallMaxDepth", "DP", a.hipstr_max_call_DP))
    if a.hipstr_min_call_Q is not None: f.append(filters.CallFilterMinValue("HipSTRCallMinQ", "Q", a.hipstr_min_call_Q))
    if a.longtr_max_call_flank_indel is not None:
        f.append(filters.HipSTRCallFlankIndels(a.longtr_max_call_flank_indel, rename="LongTRCallFlankIndels"))
    if a.longtr_min_supp_reads is not None:
        f.append(filters.HipSTRCallMinSuppReads(a.longtr_min_supp_reads, rename="LongTRMinSuppReads"))
    if a.longtr_min_call_DP is not None: f.append(filters.CallFilterMinValue("LongTRCallMinDepth", "DP", a.longtr_min_call_DP))
    if a.longtr_max_call_DP is not None: f.append(filters.CallFilterMaxValue("LongTRCallMaxDepth", "DP", a.longtr_max_call_DP))
    if a.longtr_min_call_Q is not None: f.append(filters.CallFilterMinValue("LongTRCallMinQ", "Q", a.longtr_min_call_Q))
    if a.gangstr_min_call_DP is not None: f.append(filters.CallFilterMinValue("GangSTRCallMinDepth", "DP", a.gangstr_min_call_DP))
    if a.gangstr_max_call_DP is not None: f.append(filters.CallFilterMaxValue("GangSTRCallMaxDepth", "DP", a.gangstr_max_call_DP))
    if a.gangstr_min_call_Q is not None: f.append(filters.CallFilterMinValue("GangSTRCallMinQ", "Q", a.gangstr_min_call_Q))
    if a.gangstr_expansion_prob_het is not None: f.append(filters.GangSTRCallExpansionProbHet(a.gangstr_expansion_prob_het))
    if a.gangstr_expansion_prob_hom is not None: f.append(filters.GangSTRCallExpansionProbHom(a.gangstr_expansion_prob_hom))
    if a.gangstr_expansion_prob_total is not None: f.append(filters.GangSTRCallExpansionProbTotal(a.gangstr_expansion_prob_total))
    if a.gangstr_filter_span_only: f.append(filters.GangSTRCallSpanOnly())
    if a.gangstr_filter_spanbound_only: f.append(filters.GangSTRCallSpanBoundOnly())
    if a.gangstr_filter_badCI: f.append(filters.GangSTRCallBadCI())
    if a.advntr_min_call_DP is not None: f.append(filters.CallFilterMinValue("AdVNTRCallMinDepth", "DP", a.advntr_min_call_DP))
    if a.advntr_max_call_DP is not None: f.append(filters.CallFilterMaxValue("AdVNTRCallMaxDepth", "DP", a.advntr_max_call_DP))
    if a.advntr_min_spanning is not None: f.append(filters.CallFilterMinValue("AdVNTRCallMinSpanning", "SR", a.advntr_min_spanning))
    if a.advntr_min_flanking is not None: f.append(filters.CallFilterMinValue("AdVNTRCallMinFlanking", "FR", a.advntr_min_flanking))
    if a.advntr_min_ML is not None: f.append(filters.CallFilterMinValue("AdVNTRCallMinML", "ML", a.advntr_min_ML))
    if a.eh_min_call_LC is not None: f.append(filters.CallFilterMinValue("EHCallMinDepth", "LC", a.eh_min_call_LC))
    if a.eh_max_call_LC is not None: f.append(filters.CallFilterMaxValue("EHCallMaxDepth", "LC", a.eh_max_call_LC))
    if a.eh_min_ADFL is not None: f.append(filters.CallFilterMinValue("EHCallMinADFL", "ADFL", a.eh_min_ADFL))
    if a.eh_min_ADIR is not None: f.append(filters.CallFilterMinValue("EHCallMinADFL", "ADIR", a.eh_min_ADIR))
    if a.eh_min_ADSP is not None: f.append(filters.CallFilterMinValue("EHCallMinADSP", "ADSP", a.eh_min_ADSP))
    if a.popstr_min_call_DP is not None: f.append(filters.CallFilterMinValue("PopSTRMinCallDepth", "DP", a.popstr_min_call_DP))
    if a.popstr_max_call_DP is not None: f.append(filters.CallFilterMaxValue("PopSTRMaxCallDepth", "DP", a.popstr_max_call_DP))
    if a.popstr_require_support is not None: f.append(filters.PopSTRCallRequireSupport(a.popstr_require_support))
    return f


def BuildLocusFilters(args):
    f = []
    if args.min_locus_callrate is not None: f.append(filters.Filter_MinLocusCallrate(args.min_locus_callrate))
    if args.min_locus_hwep is not None: f.append(filters.Filter_MinLocusHWEP(args.min_locus_hwep, args.use_length))
    if args.min_locus_het is not None: f.append(filters.Filter_MinLocusHet(args.min_locus_het, args.use_length))
    if args.max_locus_het is not None: f.append(filters.Filter_MaxLocusHet(args.max_locus_het, args.use_length))
    if args.filter_hrun: f.append(filters.Filter_LocusHrun())
    if args.filter_regions is not None:
        files = args.filter_regions.split(",")
        if args.filter_regions_names is not None:
            names = args.filter_regions_names.split(",")
        else:
            names = ['FILTER' + str(i) for i in range(len(files))]
        for i in range(len(names)):
            rf = filters.create_region_filter(names[i], files[i])
            if rf is None:
                raise ValueError('Could not load regions file: {}'.format(files[i]))
            f.append(rf)
    return f


# ---- block-level engine ------------------------------------------------------------------------------------
class _BlockFilterResult:
    __slots__ = ("call_mask", "gt_masked", "flags", "het", "hwep", "ac", "n_called", "hrun", "host_values", "blk")


def _needed_fmt(call_filters):
    keys = []
    for f in call_filters:
        for k in f.needs:
            if k not in keys:
                keys.append(k)
    return keys


def _filter_block(ctx, vcftype, recs, call_filters, locus_filters, sample_info, use_length, tr_records=None):
    """ApplyCallFilters + ApplyLocusFilters for a block of records (two GPU calls)."""
    fmt_keys = _needed_fmt(call_filters)
    dp_key = None
    first_fmt = recs[0].FORMAT if recs else []
    if 'DP' in first_fmt:
        dp_key = 'DP'
    elif 'LC' in first_fmt:
        dp_key = 'LC'
    if dp_key and dp_key not in fmt_keys:
        fmt_keys.append(dp_key)
    blk = _block.build_block(ctx, vcftype.name, recs, fmt_keys)
    res = _BlockFilterResult()
    res.blk = blk
    res.host_values = {}
    host_filters = [f for f in call_filters if f.gpu_kind == _lib.CF_HOST_VALUE]
    if host_filters:
        trs = tr_records or [trh.TRRecord._from_block(blk, i, r) for i, r in enumerate(recs)]
        for f in host_filters:
            vals = np.stack([f.host_values(t) for t in trs]) if blk.S else np.zeros((blk.L, 0))
            blk.add_host_filter_values("__host__" + f.name, vals)
            res.host_values[f.name] = vals
    blk._activate()
    specs = [f.gpu_spec(blk) for f in call_filters]
    n = len(call_filters)
    counts = np.zeros((n, blk.S), np.int64)
    numcalls = sample_info['numcalls']
    totaldp = sample_info['totaldp']
    dp_slot = blk.fmt_slot.get(dp_key, -1) if dp_key else -1
    out = ctx.call_filters(specs, dp_slot, counts, numcalls, totaldp, want_mask=True, want_trigger=False, want_gt=True)
    if out["negative_dp_locus"] >= 0:
        m = blk.metas[out["negative_dp_locus"]]
        raise ValueError("The following samples have calls but negative DP values "
                         "at chromosome {} pos {}".format(m.chrom, m.harmonized_pos or m.vcf_pos))
    for i, f in enumerate(call_filters):
        sample_info[f.name] += counts[i]
    res.call_mask = out["call_mask"]
    res.gt_masked = out["gt_masked"]
    lspecs = [(f.gpu_kind, f.threshold if f.gpu_kind != _lib.LF_HRUN else 0.0) if f.gpu_kind is not None else None
              for f in locus_filters]
    gpu_specs = [s for s in lspecs if s is not None]
    lres = ctx.locus_filters(gpu_specs, use_length)
    res.flags, res.het, res.hwep, res.ac, res.n_called, res.hrun = (lres["flags"], lres["het"], lres["hwep"], lres["ac"],
                                                                 lres["n_called"], lres["hrun"])
    return res


def _trigger_value(filt, blk, res, l, s):
    """The value the operator returned for one filtered call (only needed for the FORMAT:FILTER text)."""
    if filt.gpu_kind == _lib.CF_HOST_VALUE:
        return float(res.host_values[filt.name][l, s])
    if filt.gpu_kind in (_lib.CF_MIN, _lib.CF_MAX):
        return float(blk.fmt[filt.field][l, s].reshape(-1)[0])
    if filt.gpu_kind == _lib.CF_RATIO_GT:
        with np.errstate(divide='ignore', invalid='ignore'):
            return float(np.float64(blk.fmt[filt.field][l, s].reshape(-1)[0]) / np.float64(blk.fmt['DP'][l, s].reshape(-1)[0]))
    q = blk.fmt['QEXP'][l, s]
    if filt.gpu_kind == _lib.CF_QEXP_HET:
        return float(q[1])
    if filt.gpu_kind == _lib.CF_QEXP_HOM:
        return float(q[2])
    return float(np.float32(q[1]) + np.float32(q[2]))


def _filter_text(call_filters, blk, res, l):
    """FORMAT:FILTER strings of one record (dumpSTR.py:664-682)."""
    mask = res.call_mask[l]
    text = np.full(mask.shape, 'PASS', dtype=object)
    nocall = (mask & np.uint32(0x80000000)) != 0
    text[nocall] = 'NOCALL'
    fired = (mask & np.uint32(0x7fffffff)) != 0
    for s in np.nonzero(fired & ~nocall)[0]:
        parts = []
        for i, f in enumerate(call_filters):
            if (int(mask[s]) >> i) & 1:
                parts.append(f.name + '_' + ('%g' % _trigger_value(f, blk, res, l, s)))
        text[s] = ','.join(parts)
    return text.astype(str), fired & ~nocall


def _apply_to_record(record, blk, res, l, call_filters):
    """Write the block's call-filter results back into one cyvcf2 record (dumpSTR.py:684, 716-746)."""
    text, filtered = _filter_text(call_filters, blk, res, l)
    v = record
    v.set_format('FILTER', np.char.encode(text))
    if not np.any(filtered):
        return
    ploidy = v.ploidy
    gts = v.genotypes
    for idx in filtered.nonzero()[0]:
        gts[idx] = [-1] * ploidy + [False]
    v.genotypes = gts
    for field in list(v.FORMAT):
        if field in ('GT', 'FILTER'):
            continue
        vals = v.format(field)
        if vals is None:
            continue
        if vals.dtype.kind == 'U':
            vals[filtered] = '.'
            vals = np.char.encode(vals)
        elif vals.dtype.kind == 'f':
            vals[filtered] = np.nan
        elif vals.dtype.kind == 'i':
            vals[filtered] = _NOCALL_INT_FORMAT_VAL
        else:
            raise ValueError("Found an unexpected format dtype for format field " + field)
        v.set_format(field, vals)


def ApplyCallFilters(record, call_filters, sample_info, sample_names):
    """Per-record entry point kept for API compatibility (reference dumpSTR.py:613-774): a block of one."""
    blk0 = record._blk
    res = _filter_block(blk0.ctx, trh.VcfTypes[blk0.vcftype] if blk0.vcftype in trh.VcfTypes.__members__ else
                        trh.VcfTypes.gangstr, [record.vcfrecord], call_filters, [], sample_info, False, [record])
    _apply_to_record(record.vcfrecord, res.blk, res, 0, call_filters)
    out = trh.TRRecord._from_block(res.blk, 0, record.vcfrecord)
    out._masked = res
    return out


def ApplyLocusFilters(record, locus_filters, loc_info, drop_filtered) -> bool:
    """Per-record entry point (reference dumpSTR.py:917-973)."""
    filtered = False
    for filt in locus_filters:
        if filt(record) is None:
            continue
        loc_info[filt.filter_name()] += 1
        if not drop_filtered:
            if not filtered:
                record.vcfrecord.FILTER = filt.filter_name()
            else:
                record.vcfrecord.FILTER += ';' + filt.filter_name()
        filtered = True
    n_samples_called = np.sum(record.GetCalledSamples())
    if n_samples_called == 0:
        loc_info['NO_CALLS_REMAINING'] += 1
        if not drop_filtered:
            if not filtered:
                record.vcfrecord.FILTER = 'NO_CALLS_REMAINING'
            else:
                record.vcfrecord.FILTER += ';' + 'NO_CALLS_REMAINING'
        filtered = True
    if not filtered:
        if not drop_filtered:
            record.vcfrecord.FILTER = "PASS"
        loc_info["PASS"] += 1
        loc_info["totalcalls"] += n_samples_called
    return filtered


def getargs():  # pragma: no cover
    """reference dumpSTR.py:976-1058 (same flags and defaults)."""
    parser = argparse.ArgumentParser(__doc__, formatter_class=utils.ArgumentDefaultsHelpFormatter)
    g = parser.add_argument_group("Input/output")
    g.add_argument("--vcf", help="Input STR VCF file", type=str, required=True)
    g.add_argument("--out", help="Prefix for output files", type=str, required=True)
    g.add_argument("--zip", help="Produce a bgzipped and tabix indexed output VCF", action="store_true")
    g.add_argument("--vcftype", help="Options=%s" % [str(item) for item in trh.VcfTypes.__members__], type=str, default="auto")
    g = parser.add_argument_group("Locus-level filters (tool agnostic)")
    g.add_argument("--min-locus-callrate", help="Minimum locus call rate", type=float)
    g.add_argument("--min-locus-hwep", help="Filter loci failing HWE at this p-value threshold", type=float)
    g.add_argument("--min-locus-het", help="Minimum locus heterozygosity", type=float)
    g.add_argument("--max-locus-het", help="Maximum locus heterozygosity", type=float)
    g.add_argument("--use-length", help="Calculate per-locus stats (het, HWE) collapsing alleles by length", action="store_true")
    g.add_argument("--filter-regions", help="Comma-separated list of BED files of regions to filter. Must be bgzipped and tabix indexed", type=str)
    g.add_argument("--filter-regions-names", help="Comma-separated list of filter names for each BED filter file", type=str)
    g.add_argument("--filter-hrun", help="Filter STRs with long homopolymer runs.", action="store_true")
    g.add_argument("--drop-filtered", help="Drop filtered records from output", action="store_true")
    for tool, flagset in (
        ("HipSTR", [("--hipstr-max-call-flank-indel", float), ("--hipstr-max-call-stutter", float), ("--hipstr-min-supp-reads", int),
                    ("--hipstr-min-call-DP", int), ("--hipstr-max-call-DP", int), ("--hipstr-min-call-Q", float)]),
        ("LongTR", [("--longtr-max-call-flank-indel", float), ("--longtr-min-supp-reads", int), ("--longtr-min-call-DP", int),
                    ("--longtr-max-call-DP", int), ("--longtr-min-call-Q", float)]),
        ("GangSTR", [("--gangstr-min-call-DP", int), ("--gangstr-max-call-DP", int), ("--gangstr-min-call-Q", float),
                     ("--gangstr-expansion-prob-het", float), ("--gangstr-expansion-prob-hom", float),
                     ("--gangstr-expansion-prob-total", float)]),
        ("adVNTR", [("--advntr-min-call-DP", int), ("--advntr-max-call-DP", int), ("--advntr-min-spanning", int),
                    ("--advntr-min-flanking", int), ("--advntr-min-ML", float)]),
        ("ExpansionHunter", [("--eh-min-ADFL", int), ("--eh-min-ADIR", int), ("--eh-min-ADSP", int), ("--eh-min-call-LC", int),
                             ("--eh-max-call-LC", int)]),
        ("PopSTR", [("--popstr-min-call-DP", int), ("--popstr-max-call-DP", int), ("--popstr-require-support", int)]),
    ):
        g = parser.add_argument_group("Call-level filters specific to {} output".format(tool))
        for flag, typ in flagset:
            g.add_argument(flag, type=typ, help="see the TRTools dumpSTR documentation")
        if tool == "GangSTR":
            g.add_argument("--gangstr-filter-span-only", help="Filter out all calls that only have spanning read support", action="store_true")
            g.add_argument("--gangstr-filter-spanbound-only", help="Filter out all reads except spanning and bounding", action="store_true")
            g.add_argument("--gangstr-filter-badCI", help="Filter regions where the ML estimate is not in the CI", action="store_true")
    g = parser.add_argument_group("Debugging parameters")
    g.add_argument("--num-records", help="Only process this many records", type=int)
    g.add_argument("--die-on-warning", help="Quit if a record can't be parsed", action="store_true")
    g.add_argument("--verbose", help="Print out extra info", action="store_true")
    g = parser.add_argument_group("GPU")
    g.add_argument("--block-size", help="Records staged per GPU block", type=int, default=512)
    g = parser.add_argument_group("Version")
    g.add_argument("--version", action="version", version='{version}'.format(version=__version__))
    return parser.parse_args()


_INFO_DEFS = [("AC", 'Alternate allele counts', 'Integer', 'A'), ("REFAC", 'Reference allele count', 'Integer', 1),
              ("HET", 'Heterozygosity', 'Float', 1), ("HWEP", 'HWE p-value for obs. vs. exp het rate', 'Float', 1),
              ("HRUN", 'Length of longest homopolymer run', 'Integer', 1)]

_FIELD_ISSUE = ("Error: The {} field '{}' is present in the input VCF and doesn't have the expected Type and Number "
                "so it can't be worked with. Please use 'bcftools annotate --rename-annots' or another equivalent tool "
                "to rename or remove the field and then rerun dumpSTR.")


def main(args):
    """reference dumpSTR.py:1060-1354."""
    invcf = utils.LoadSingleReader(args.vcf, checkgz=False)
    if invcf is None:
        return 1
    if not os.path.exists(os.path.dirname(os.path.abspath(args.out))):
        common.WARNING("Error: The directory which contains the output location {} does not exist".format(args.out))
        return 1
    if os.path.isdir(args.out + ".vcf"):
        common.WARNING("Error: The output location {} is a directory".format(args.out))
        return 1
    if args.out[-1] in {'.', '/'}:
        common.WARNING("Output prefix must not end in '/' or '.'")
        return 1
    block_size = int(getattr(args, "block_size", 512) or 512)
    vcftype = trh.InferVCFType(invcf, args.vcftype)
    is_beagle = trh.IsBeagleVCF(invcf)

    format_fields, info_fields, preexisting_filter_fields = {}, {}, {}
    for h in invcf.header_iter():
        if h['HeaderType'] == 'INFO':
            info_fields[h['ID']] = h
        elif h['HeaderType'] == 'FORMAT':
            format_fields[h['ID']] = h
        elif h['HeaderType'] == 'FILTER':
            preexisting_filter_fields[h['ID']] = h
    if not CheckFilters(format_fields, args, vcftype, is_beagle):
        return 1

    field_issues = False
    if 'FILTER' not in format_fields:
        invcf.add_format_to_header({'ID': 'FILTER', 'Description': 'call-level filters that have been applied',
                                    'Type': 'String', 'Number': 1})
    elif format_fields['FILTER']['Type'] != 'String' or format_fields['FILTER']['Number'] != '1':
        field_issues = True
        common.WARNING(_FIELD_ISSUE.format('format', 'FILTER'))
    for fid, desc, typ, num in _INFO_DEFS:
        if fid not in info_fields:
            invcf.add_info_to_header({'ID': fid, 'Description': desc, 'Type': typ, 'Number': num})
        elif info_fields[fid]['Type'] != typ or info_fields[fid]['Number'] != str(num):
            field_issues = True
            common.WARNING(_FIELD_ISSUE.format('info', fid))
        elif info_fields[fid]['Description'].strip('"') != desc:
            common.WARNING("Overwriting the preexisting info {} field".format(fid))
    if field_issues:
        return 1

    invcf.add_filter_to_header({"ID": "NO_CALLS_REMAINING",
                                "Description": ("All calls at this locus were already nocalls or were individually "
                                                "filtered before the locus level filters were applied.")})
    try:
        locus_filters = BuildLocusFilters(args)
    except ValueError:
        return 1
    for f in locus_filters:
        if f.filter_name() not in preexisting_filter_fields:
            invcf.add_filter_to_header({"ID": f.filter_name(), "Description": f.description()})
        elif preexisting_filter_fields[f.filter_name()]['Description'].strip('"') != f.description():
            common.WARNING("Using locus level filter " + f.filter_name() + "which has the same name as a FILTER field "
                           "that already exists in the input VCF.")
    call_filters = BuildCallFilters(args)
    if len(call_filters) > 16:
        common.WARNING("At most 16 call-level filters can be combined")
        return 1

    ctx = _lib.default_context()
    # several GPUs (torchrun): blocks are dealt round-robin; rank 0 gathers the records' text (NCCL) and writes the
    # VCF, the per-sample accumulators and locus counters are summed over ranks (NaN poison preserved, dumpSTR.py:710-713)
    comm = _dist.cli_comm(ctx)
    sharder = _dist.BlockSharder(comm)
    suffix = '.vcf.gz' if args.zip else '.vcf'
    outvcf = None
    if sharder.rank == 0:
        outvcf = MakeWriter(args.out + suffix, invcf, " ".join(sys.argv))
    failed = np.array([1 if (sharder.rank == 0 and outvcf is None) else 0], dtype=np.int64)
    if int(_dist.allreduce_sum(comm, failed)[0]):
        if comm is not None:
            comm.close()
        return 1

    n_samples = len(invcf.samples)
    sample_info = collections.OrderedDict()
    sample_info['numcalls'] = np.zeros((n_samples), dtype=np.int64)
    sample_info['totaldp'] = np.zeros((n_samples), dtype=float)
    for name in GetAllCallFilters(call_filters):
        sample_info[name] = np.zeros((n_samples), dtype=np.int64)
    loc_info = collections.OrderedDict()
    loc_info["totalcalls"] = 0
    loc_info["PASS"] = 0
    loc_info["NO_CALLS_REMAINING"] = 0
    for filt in locus_filters:
        loc_info[filt.filter_name()] = 0

    if hasattr(invcf, "_prefetch"):
        # C++ block reader: parse the call filters' numeric FORMAT keys in the same pass as GT,
        # in runs of the GPU block length
        invcf._prefetch = tuple(dict.fromkeys(_needed_fmt(call_filters) + ['DP', 'LC']))
        invcf._native_block_loci = block_size
    harmonizer_idx = 1
    record_counter = 0
    done = False
    it = iter(invcf)
    while not done:
        recs = []
        parse_error = None
        while len(recs) < block_size:
            harmonizer_idx += 1
            try:
                rec = next(it)
            except StopIteration:
                done = True
                break
            except Exception:
                parse_error = ("Unable to parse the " + str(harmonizer_idx) + "th tandem repeat in the provided VCF. "
                               "Check that it is properly formatted.")
                done = True
                break
            if recs and (rec.ploidy != recs[0].ploidy):
                it = itertools.chain([rec], it)
                harmonizer_idx -= 1
                break
            record_counter += 1
            if args.num_records is not None and record_counter > args.num_records:
                done = True
                break
            recs.append(rec)
        # a record missing mandatory INFO fields ends the run when it is reached (dumpSTR.py:1273-1289)
        bad_message = None
        good = []
        for r in recs:
            try:
                _block.record_meta(vcftype.name, r)
                good.append(r)
            except TypeError as te:
                message = te.args[0]
                if 'missing' in message and 'mandatory' in message:
                    bad_message = message
                    break
                raise
        recs = good
        if recs and not sharder.mine():
            recs = []
        if recs:
            chunk = []
            if args.verbose:
                for r in recs:
                    common.MSG("Processing %s:%s" % (r.CHROM, r.POS), debug=True)
            res = _filter_block(ctx, vcftype, recs, call_filters, locus_filters, sample_info, args.use_length)
            blk = res.blk
            host_locus = [f for f in locus_filters if f.gpu_kind is None]
            for l, rec in enumerate(recs):
                _apply_to_record(rec, blk, res, l, call_filters)
                names = []
                bit = 0
                tr_view = None
                for f in locus_filters:
                    if f.gpu_kind is not None:
                        hit = (int(res.flags[l]) >> bit) & 1
                        bit += 1
                    else:
                        if tr_view is None:
                            tr_view = trh.TRRecord._from_block(blk, l, rec)
                        hit = f(tr_view) is not None
                    if hit:
                        loc_info[f.filter_name()] += 1
                        names.append(f.filter_name())
                n_called = int(res.n_called[l])
                if n_called == 0:
                    loc_info['NO_CALLS_REMAINING'] += 1
                    names.append('NO_CALLS_REMAINING')
                locus_filtered = bool(names)
                if not locus_filtered:
                    loc_info["PASS"] += 1
                    loc_info["totalcalls"] += n_called
                if args.drop_filtered and locus_filtered:
                    continue
                if not args.drop_filtered:
                    rec.FILTER = ";".join(names) if names else "PASS"
                sl = blk.allele_slice(l)
                ac = res.ac[sl]
                rec.INFO['HRUN'] = int(res.hrun[l])
                if n_called > 0:
                    rec.INFO['HET'] = float(res.het[l])
                    rec.INFO['HWEP'] = float(res.hwep[l])
                else:
                    rec.INFO['HET'] = -1
                    rec.INFO['HWEP'] = -1
                rec.INFO['AC'] = ",".join(str(int(x)) for x in ac[1:]) if len(ac) > 1 else 0
                rec.INFO['REFAC'] = int(ac[0])
                if comm is None:
                    outvcf.write_record(rec)
                else:
                    chunk.append(str(rec))
            if comm is not None:
                sharder.add("".join(chunk))
        if bad_message is not None or parse_error is not None:      # every rank reads every record: all ranks stop here
            if sharder.rank == 0:
                common.WARNING("Could not parse VCF.\n" + (bad_message if bad_message is not None else parse_error))
            if comm is not None:
                comm.close()
            return 1
    invcf.close()
    if comm is not None:
        merged = sharder.finish()
        if merged is not None:
            outvcf.write_text(b"".join(merged).decode("utf-8"))
        for k in list(sample_info.keys()):
            sample_info[k] = _dist.allreduce_sum(comm, np.ascontiguousarray(sample_info[k]))
        keys = list(loc_info.keys())
        tot = _dist.allreduce_sum(comm, np.array([loc_info[k] for k in keys], dtype=np.int64))
        for k, v in zip(keys, tot):
            loc_info[k] = int(v)
        comm.barrier()
        comm.close()
        if sharder.rank != 0:
            return 0
    outvcf.close()
    WriteSampLog(sample_info, invcf.samples, args.out + ".samplog.tab")
    WriteLocLog(loc_info, args.out + ".loclog.tab")
    if args.zip:
        # reference dumpSTR.py:1347-1352 shells out to `tabix`; the index is written natively here (same linear index,
        # same header; see tabix_index.py), so --zip does not depend on the binary being installed
        try:
            from .tabix_index import write_tbi
            write_tbi(args.out + suffix)
        except Exception as e:
            common.WARNING("Tabix failed with returncode 1 ({})".format(e))
            return 1
    return 0


def run():  # pragma: no cover
    sys.exit(main(getargs()))


if __name__ == "__main__":  # pragma: no cover
    run()
