"""
Locus-level and call-level filter operators (drop-in for reference trtools/dumpSTR/filters.py).

Same class names, constructor arguments, ``.name`` / ``filter_name()`` strings and ``__call__``
contracts as the reference, but the numeric operators are descriptors for the CUDA kernels:
``gpu_spec(block)`` returns the ``(kind, field slot, threshold)`` triple that ``trt_call_filters``
evaluates for a whole block of loci, and ``__call__(record)`` (the reference's per-record entry point)
runs the same kernel on that record's block.  The operators that parse per-sample strings
(ALLREADS/GB, RC, REPCI — SURVEY.md §8a row D4) are evaluated on the host exactly as the reference does
and merged into the kernel's filter order as ``TRT_CF_HOST_VALUE`` fields.
"""
import ast
import gzip
import os
from typing import Optional

import numpy as np

from . import _lib, common


class FilterBase:
    """Interface of locus-level filters (reference filters.py:15-29)."""
    name = 'NotYetImplemented'
    gpu_kind: Optional[int] = None

    def __call__(self, record):
        raise NotImplementedError

    def filter_name(self):
        raise NotImplementedError

    def description(self):
        return ''

    def gpu_spec(self):
        """(kind, threshold) for trt_locus_filters, or None for host-only filters."""
        if self.gpu_kind is None:
            return None
        return (self.gpu_kind, getattr(self, "threshold", 0.0))


def _locus_result(record, flt):
    """Evaluate one locus filter for one record through trt_locus_filters; returns the reference's
    ``__call__`` value (None = not filtered)."""
    blk = record._blk
    blk._activate()
    res = blk.ctx.locus_filters([flt.gpu_spec()], getattr(flt, "uselength", False))
    l = record._l
    if not (int(res["flags"][l]) & 1):
        return None
    if flt.gpu_kind == _lib.LF_CALLRATE:
        return res["n_called"][l] / blk.S
    if flt.gpu_kind == _lib.LF_HWE:
        return float(res["hwep"][l])
    if flt.gpu_kind in (_lib.LF_HETLOW, _lib.LF_HETHIGH):
        return float(res["het"][l])
    return int(res["hrun"][l])


class Filter_MinLocusCallrate(FilterBase):
    """reference filters.py:35-64."""
    name = 'CALLRATE'
    gpu_kind = _lib.LF_CALLRATE

    def __init__(self, min_locus_callrate):
        self.threshold = min_locus_callrate

    def __call__(self, record):
        return _locus_result(record, self)

    def filter_name(self):
        return self.name + str(self.threshold)


class Filter_MinLocusHWEP(FilterBase):
    """reference filters.py:66-106."""
    name = 'HWE'
    gpu_kind = _lib.LF_HWE

    def __init__(self, min_locus_hwep, uselength=False):
        self.threshold = min_locus_hwep
        self.uselength = uselength

    def __call__(self, record):
        return _locus_result(record, self)

    def filter_name(self):
        return self.name + str(self.threshold)


class Filter_MinLocusHet(FilterBase):
    """reference filters.py:108-147."""
    name = 'HETLOW'
    gpu_kind = _lib.LF_HETLOW

    def __init__(self, min_locus_het, uselength=False):
        self.threshold = min_locus_het
        self.uselength = uselength

    def __call__(self, record):
        return _locus_result(record, self)

    def filter_name(self):
        return self.name + str(self.threshold)


class Filter_MaxLocusHet(FilterBase):
    """reference filters.py:149-188."""
    name = 'HETHIGH'
    gpu_kind = _lib.LF_HETHIGH

    def __init__(self, max_locus_het, uselength=False):
        self.threshold = max_locus_het
        self.uselength = uselength

    def __call__(self, record):
        return _locus_result(record, self)

    def filter_name(self):
        return self.name + str(self.threshold)


class Filter_LocusHrun(FilterBase):
    """reference filters.py:190-217 (5-/6-mers with a homopolymer run >= the period)."""
    name = 'HRUN'
    gpu_kind = _lib.LF_HRUN
    threshold = 0.0

    def __init__(self):
        pass

    def __call__(self, record):
        return _locus_result(record, self)

    def filter_name(self):
        return self.name


class _BedIndex:
    """Overlap queries against a (b)gzipped BED file (stands in for pysam.TabixFile.fetch)."""

    def __init__(self, filename):
        per = {}
        with gzip.open(filename, "rt") as f:
            for line in f:
                if not line.strip() or line.startswith(("#", "track", "browser")):
                    continue
                c = line.split("\t")
                per.setdefault(c[0], []).append((int(c[1]), int(c[2])))
        self.starts, self.maxend = {}, {}
        for chrom, iv in per.items():
            iv.sort()
            s = np.array([x[0] for x in iv], dtype=np.int64)
            e = np.array([x[1] for x in iv], dtype=np.int64)
            self.starts[chrom] = s
            self.maxend[chrom] = np.maximum.accumulate(e)

    def overlaps(self, chrom, beg1, end1) -> bool:
        """any BED interval [bs, be) overlapping the 1-based closed region [beg1, end1]"""
        if chrom not in self.starts:
            return False
        s = self.starts[chrom]
        k = int(np.searchsorted(s, end1, side="left"))     # intervals with bs < end1
        return k > 0 and bool(self.maxend[chrom][k - 1] > beg1 - 1)


def create_region_filter(name, filename):
    """reference filters.py:219-300: locus filter for records overlapping a BED file.  Host-side
    (interval lookup per locus); returns None if the file fails the reference's checks."""

    class Filter_Regions(FilterBase):
        def __init__(self, name, filename):
            self.threshold = ""
            self.name = name
            self.pass_checks = True
            self.LoadRegions(filename)

        def LoadRegions(self, filename):
            self.regions = None
            if not filename.endswith(".bed.gz") and not filename.endswith(".bed.bgz"):
                common.WARNING("Make sure %s is bgzipped and indexed" % filename)
                self.pass_checks = False
                return
            if not os.path.isfile(filename):
                common.WARNING("Could not find regions BED file %s" % filename)
                self.pass_checks = False
                return
            if not os.path.isfile(filename + ".tbi"):
                common.WARNING("Could not find tabix index %s.tbi" % filename)
                self.pass_checks = False
                return
            self.regions = _BedIndex(filename)

        def __call__(self, record):
            if self.regions is None:
                return None
            chrom = str(record.chrom)
            beg = int(record.pos)
            end = int(record.pos + record.ref_allele_length)      # htslib parses 'pos+len' and drops the fraction
            other = chrom.replace("chr", "") if "chr" in chrom else "chr" + chrom
            if self.regions.overlaps(chrom, beg, end) or self.regions.overlaps(other, beg, end):
                return self.name
            return None

        def filter_name(self):
            return self.name

        def description(self):
            return 'Filter TRs overlapping this region'

    f = Filter_Regions(name, filename)
    if not f.pass_checks:
        return None
    return f


# ---------------------------------------------------------------------------------------------------
# call-level filters
# ---------------------------------------------------------------------------------------------------
class Reason:
    """Base call-level filter (reference filters.py:306-325): ``__call__(record)`` returns one float
    per sample, NaN = not filtered, anything else = the value that triggered the filter."""
    name = ""
    gpu_kind: Optional[int] = None
    field: Optional[str] = None       # FORMAT field the kernel reads
    needs = ()                        # numeric FORMAT fields this operator needs in the block

    def GetReason(self):
        return self.name

    def gpu_spec(self, blk):
        """(kind, field slot, threshold) for trt_call_filters."""
        slot = blk.fmt_slot[self.field] if self.field is not None else 0
        return (self.gpu_kind, slot, float(self.threshold))

    def host_values(self, record):
        """Host-evaluated operators override this (float array [S], NaN = keep)."""
        return None

    def __call__(self, record):
        from . import block as _block
        hv = self.host_values(record)
        if hv is not None:
            return hv
        ctx = record._blk.ctx
        blk = _block.build_block(ctx, record._blk.vcftype, [record.vcfrecord], self.needs)
        counts = np.zeros((1, blk.S), np.int64)
        res = ctx.call_filters([self.gpu_spec(blk)], -1, counts, np.zeros(blk.S, np.int64), np.zeros(blk.S),
                               want_mask=False, want_trigger=True, want_gt=False)
        record._blk._activate()
        return res["trigger_values"][0, 0].copy()


class CallFilterMinValue(Reason):
    """reference filters.py:327-367."""
    gpu_kind = _lib.CF_MIN

    def __init__(self, name, field, threshold):
        self.name = name + str(threshold)
        self.field = field
        self.threshold = threshold
        self.needs = (field,)


class CallFilterMaxValue(Reason):
    """reference filters.py:369-409."""
    gpu_kind = _lib.CF_MAX

    def __init__(self, name, field, threshold):
        self.name = name + str(threshold)
        self.field = field
        self.threshold = threshold
        self.needs = (field,)


class HipSTRCallFlankIndels(Reason):
    """reference filters.py:415-449 (DFLANKINDEL/DP > threshold)."""
    name = "HipSTRCallFlankIndels"
    gpu_kind = _lib.CF_RATIO_GT
    field = "DFLANKINDEL"
    needs = ("DFLANKINDEL", "DP")

    def __init__(self, threshold, rename=None):
        self.threshold = threshold
        if rename is not None:
            self.name = rename
        self.name += str(threshold)


class HipSTRCallStutter(Reason):
    """reference filters.py:451-484 (DSTUTTER/DP > threshold)."""
    name = "HipSTRCallStutter"
    gpu_kind = _lib.CF_RATIO_GT
    field = "DSTUTTER"
    needs = ("DSTUTTER", "DP")

    def __init__(self, threshold, rename=None):
        self.threshold = threshold
        if rename is not None:
            self.name = rename
        self.name += str(threshold)


class _HostFilter(Reason):
    gpu_kind = _lib.CF_HOST_VALUE
    threshold = 0.0

    def gpu_spec(self, blk):
        return (self.gpu_kind, blk.fmt_slot["__host__" + self.name], 0.0)


class HipSTRCallMinSuppReads(_HostFilter):
    """reference filters.py:486-567: per-sample ALLREADS ('bp|n;...') / GB ('a|b') string parsing —
    host-side."""
    name = "HipSTRMinSuppReads"

    def __init__(self, threshold, rename=None):
        self.threshold = threshold
        if rename is not None:
            self.name = rename
        self.name += str(threshold)

    def host_values(self, record):
        n = record.GetNumSamples()
        called = record.GetCalledSamples()
        if not np.any(called):
            return np.full((n,), np.nan)
        if "ALLREADS" not in record.format:
            return np.zeros((n,), dtype=float)
        allreads = record.format["ALLREADS"]
        check = called & (allreads != '') & (allreads != '.')
        if not np.any(check):
            out = np.full((n,), np.nan)
            out[called] = 0
            return out
        gbs = record.format["GB"]
        first = gbs[check][0]
        if "/" in first:
            delim = "/"
        elif "|" in first:
            delim = "|"
        else:
            raise ValueError("Cant't identify phasing char ('|' or '/') in GB field")
        out = np.full((n,), np.nan)
        for i in np.nonzero(check)[0]:
            reads = ast.literal_eval("{" + str(allreads[i]).replace(";", ",").replace("|", ":") + "}")
            low = np.inf
            for g in str(gbs[i]).split(delim):
                low = min(low, reads[int(g)]) if int(g) in reads else 0
            out[i] = low
        out[out >= self.threshold] = np.nan
        out[called & ~check] = 0
        return out


class GangSTRCallExpansionProbHom(Reason):
    """reference filters.py:573-605."""
    name = "GangSTRCallExpansionProbHom"
    gpu_kind = _lib.CF_QEXP_HOM
    field = "QEXP"
    needs = ("QEXP",)

    def __init__(self, threshold):
        self.threshold = threshold
        self.name += str(threshold)


class GangSTRCallExpansionProbHet(Reason):
    """reference filters.py:607-639."""
    name = "GangSTRCallExpansionProbHet"
    gpu_kind = _lib.CF_QEXP_HET
    field = "QEXP"
    needs = ("QEXP",)

    def __init__(self, threshold):
        self.threshold = threshold
        self.name += str(threshold)


class GangSTRCallExpansionProbTotal(Reason):
    """reference filters.py:641-674."""
    name = "GangSTRCallExpansionProbTotal"
    gpu_kind = _lib.CF_QEXP_TOT
    field = "QEXP"
    needs = ("QEXP",)

    def __init__(self, threshold):
        self.threshold = threshold
        self.name += str(threshold)


class GangSTRCallSpanOnly(_HostFilter):
    """reference filters.py:676-697 (RC 'a,b,c,d' string parse; host-side)."""
    name = "GangSTRCallSpanOnly"

    def __init__(self):
        pass

    def host_values(self, record):
        out = np.full((record.GetNumSamples()), np.nan)
        called = record.GetCalledSamples()
        if not np.any(called):
            return out
        rc = np.stack(np.char.split(record.format['RC'][called], ','), axis=0).astype(int)
        hit = rc[:, 1] == record.format['DP'][called, 0]
        out[np.nonzero(called)[0][hit]] = rc[:, 1][hit]
        return out


class GangSTRCallSpanBoundOnly(_HostFilter):
    """reference filters.py:699-722."""
    name = "GangSTRCallSpanBoundOnly"

    def __init__(self):
        pass

    def host_values(self, record):
        out = np.full((record.GetNumSamples()), np.nan)
        called = record.GetCalledSamples()
        if not np.any(called):
            return out
        rc = np.stack(np.char.split(record.format['RC'][called], ','), axis=0).astype(int)
        sb = rc[:, 1] + rc[:, 3]
        hit = sb == record.format['DP'][called, 0]
        out[np.nonzero(called)[0][hit]] = sb[hit]
        return out


class GangSTRCallBadCI(_HostFilter):
    """reference filters.py:724-757 (REPCI 'lo-hi,lo-hi' vs REPCN; host-side)."""
    name = "GangSTRCallBadCI"

    def __init__(self):
        pass

    def host_values(self, record):
        out = np.full((record.GetNumSamples()), np.nan)
        called = record.GetCalledSamples()
        if not np.any(called):
            return out
        ml = record.format["REPCN"][called]
        ci = np.stack(np.char.split(record.format["REPCI"][called], ","))
        ci = np.array(np.char.split(ci, '-').tolist(), dtype=int)
        bad = np.logical_or(ml < ci[:, :, 0], ci[:, :, 1] < ml)
        rows = np.any(bad, axis=1)
        if not np.any(rows):
            return out
        which = np.argmax(bad[rows, :], axis=1)
        out[np.nonzero(called)[0][rows]] = ml[rows, which]
        return out


class PopSTRCallRequireSupport(_HostFilter):
    """reference filters.py:835-867 (AD[sample, genotype] < threshold; host-side gather)."""
    name = "PopSTRCallRequireSupport"

    def __init__(self, threshold):
        self.threshold = threshold
        self.name += str(threshold)

    def host_values(self, record):
        n = record.GetNumSamples()
        out = np.full((n,), np.nan)
        ad = record.format["AD"]
        gt = record.GetGenotypeIndicies()[:, :-1]
        rows = np.arange(n)
        for ploid in range(gt.shape[1]):
            support = ad[rows, gt[:, ploid]]
            low = support < self.threshold
            out[low] = support[low]
        return out
