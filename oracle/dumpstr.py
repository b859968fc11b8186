"""
TEST INFRASTRUCTURE ONLY — restatement of the dumpSTR hot path of the reference:
call-level filter operators (``trtools/dumpSTR/filters.py:327-484, 573-674, 835-867``),
``ApplyCallFilters`` (``trtools/dumpSTR/dumpSTR.py:613-774``), locus filters
(``filters.py:35-217``), ``ApplyLocusFilters`` (``dumpSTR.py:917-973``), the INFO
recompute (``dumpSTR.py:1307-1336``) and the two logs (``dumpSTR.py:523-588``).

The text side (FORMAT:FILTER strings) is restated too because the golden
``*.samplog.tab``/``*.loclog.tab`` and FILTER strings pin it.  String-parsing call filters
(HipSTRCallMinSuppReads, GangSTR RC/REPCI — SURVEY.md §8a row D4) are out of the
BASELINE configs and are not restated.
"""
import collections
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np

from . import stats, trh
from .records import Locus

INT32_MISSING = -2147483648


# ---- call filter operators ----------------------------------------------------
@dataclass
class CallFilter:
    kind: str                  # 'min', 'max', 'ratio_gt', 'qexp_het', 'qexp_hom', 'qexp_total', 'popstr_support'
    name: str
    threshold: float
    field: Optional[str] = None
    num_field: Optional[str] = None   # numerator for ratio filters

    def __call__(self, locus: Locus, gt: np.ndarray) -> np.ndarray:
        n = gt.shape[0]
        out = np.full((n,), np.nan)
        if self.kind == 'min':                                        # filters.py:363-367
            v = locus.fmt[self.field][:, 0]
            out[v < self.threshold] = v[v < self.threshold]
        elif self.kind == 'max':                                      # filters.py:405-409
            v = locus.fmt[self.field][:, 0]
            out[v > self.threshold] = v[v > self.threshold]
        elif self.kind == 'ratio_gt':                                 # filters.py:444-449, 479-484
            with np.errstate(divide='ignore', invalid='ignore'):
                ratio = locus.fmt[self.num_field][:, 0] / locus.fmt['DP'][:, 0]
            out[ratio > self.threshold] = ratio[ratio > self.threshold]
        elif self.kind in ('qexp_het', 'qexp_hom', 'qexp_total'):     # filters.py:573-674
            called = trh.called_samples(gt)
            if not np.any(called):
                return out
            q = locus.fmt['QEXP']
            if self.kind == 'qexp_hom':
                p = q[called, 2]
            elif self.kind == 'qexp_het':
                p = q[called, 1]
            else:
                p = q[called, 1] + q[called, 2]
            out[np.nonzero(called)[0][p < self.threshold]] = p[p < self.threshold]
        elif self.kind == 'popstr_support':                           # filters.py:835-867
            ad = locus.fmt['AD']
            idx = trh.genotype_indices(gt)[:, :-1]
            rows = np.arange(n)
            for ploid in range(idx.shape[1]):
                bad = ad[rows, idx[:, ploid]] < self.threshold
                # (the reference indexes AD with a bool mask AND an int vector, which only
                #  broadcasts for 1 offending sample or 1-sample VCFs; the intent is restated)
                out[bad] = ad[rows, idx[:, ploid]][bad]
        else:
            raise ValueError(self.kind)
        return out


def min_value(name, fld, thr):
    return CallFilter('min', name + str(thr), thr, field=fld)


def max_value(name, fld, thr):
    return CallFilter('max', name + str(thr), thr, field=fld)


def hipstr_flank_indels(thr, rename=None):
    return CallFilter('ratio_gt', (rename or "HipSTRCallFlankIndels") + str(thr), thr, num_field='DFLANKINDEL')


def hipstr_stutter(thr, rename=None):
    return CallFilter('ratio_gt', (rename or "HipSTRCallStutter") + str(thr), thr, num_field='DSTUTTER')


def gangstr_expansion(kind, thr):
    nm = {'qexp_hom': 'GangSTRCallExpansionProbHom', 'qexp_het': 'GangSTRCallExpansionProbHet',
          'qexp_total': 'GangSTRCallExpansionProbTotal'}[kind]
    return CallFilter(kind, nm + str(thr), thr)


def popstr_require_support(thr):
    return CallFilter('popstr_support', "PopSTRCallRequireSupport" + str(thr), thr)


# ---- ApplyCallFilters ----------------------------------------------------------
@dataclass
class CallFilterResult:
    gt: np.ndarray                      # masked genotype array int16 [S, P+1]
    filter_text: np.ndarray             # FORMAT:FILTER strings [S]
    fmt: Dict[str, np.ndarray]          # FORMAT arrays after nulling filtered calls
    filtered: np.ndarray                # bool [S]: call removed by a call-level filter


def new_sample_info(n_samples: int, call_filters: List[CallFilter]):
    """dumpSTR.py:1251-1260."""
    info = collections.OrderedDict()
    info['numcalls'] = np.zeros((n_samples,), dtype=int)
    info['totaldp'] = np.zeros((n_samples,), dtype=float)
    for f in call_filters:
        info[f.name] = np.zeros((n_samples,), dtype=int)
    return info


def apply_call_filters(locus: Locus, call_filters: List[CallFilter], sample_info) -> CallFilterResult:
    """dumpSTR.py:613-774."""
    gt = np.array(locus.gt, dtype=np.int16)
    n = gt.shape[0]
    text = np.empty((n,), 'U4')
    nocalls = ~trh.called_samples(gt)
    for filt in call_filters:
        vals = filt(locus, gt)
        nans = np.isnan(vals)
        if np.all(nans):
            continue
        sample_info[filt.name] += np.logical_and(~nans, ~nocalls)
        t = np.char.add(filt.name, np.char.add('_', np.char.mod('%g', vals)))
        t[nans] = ''
        not_first = np.logical_and(~nans, text != '')
        text[not_first] = np.char.add(text[not_first], ',')
        text = np.char.add(text, t)
    if np.any(nocalls):
        nc = np.empty((n,), dtype='U6')
        nc[nocalls] = 'NOCALL'
        text[nocalls] = ''
        text = np.char.add(text, nc)
    text[text == ''] = 'PASS'
    extant = text == 'PASS'
    sample_info['numcalls'] += extant
    dp = locus.fmt.get('DP', locus.fmt.get('LC'))
    if dp is not None:
        dp = dp.reshape(-1)
        neg = np.logical_and(np.logical_and(dp < 0, dp != INT32_MISSING), extant)
        if np.any(neg):
            raise ValueError("The following samples have calls but negative DP values "
                             "at chromosome {} pos {}".format(locus.chrom, locus.pos))
        acc = np.logical_and(extant, dp > 0)
        sample_info['totaldp'][acc] += dp[acc]
        sample_info['totaldp'][np.logical_and(extant, dp == INT32_MISSING)] = np.nan
    else:
        sample_info['totaldp'][:] = np.nan
    filtered = np.logical_and(text != 'PASS', text != 'NOCALL')
    fmt = {k: np.array(v, copy=True) for k, v in locus.fmt.items()}
    if np.any(filtered):
        ploidy = gt.shape[1] - 1
        gt[filtered, :ploidy] = -1                                    # :722-727
        gt[filtered, ploidy] = 0
        for k, v in fmt.items():                                      # :730-746
            if v.dtype.kind == 'U':
                v[filtered] = '.'
            elif v.dtype.kind == 'f':
                v[filtered] = np.nan
            elif v.dtype.kind == 'i':
                v[filtered] = INT32_MISSING
    return CallFilterResult(gt=gt, filter_text=text, fmt=fmt, filtered=filtered)


# ---- locus filters --------------------------------------------------------------
@dataclass
class LocusFilter:
    kind: str            # 'callrate', 'hwe', 'hetlow', 'hethigh', 'hrun'
    threshold: Optional[float] = None
    uselength: bool = False

    def filter_name(self):
        base = {'callrate': 'CALLRATE', 'hwe': 'HWE', 'hetlow': 'HETLOW', 'hethigh': 'HETHIGH',
                'hrun': 'HRUN'}[self.kind]
        return base if self.kind == 'hrun' else base + str(self.threshold)

    def __call__(self, locus: Locus, h: trh.Harmonized, gt):
        if self.kind == 'callrate':                                   # filters.py:35-64
            cr = trh.call_rate(gt)
            return cr if cr < self.threshold else None
        if self.kind == 'hwe':                                        # filters.py:66-106
            f = trh.allele_freqs(h, gt, uselength=self.uselength)
            g = trh.genotype_counts(h, gt, uselength=self.uselength)
            p = stats.hardy_weinberg_binomial_test(f, g)
            return p if p < self.threshold else None
        if self.kind in ('hetlow', 'hethigh'):                        # filters.py:108-188
            het = stats.heterozygosity(trh.allele_freqs(h, gt, uselength=self.uselength))
            if self.kind == 'hetlow':
                return het if het < self.threshold else None
            return het if het > self.threshold else None
        if self.kind == 'hrun':                                       # filters.py:190-217
            seq = h.full_alleles[0] if h.full_alleles is not None else h.ref_allele
            hrun = trh.homopolymer_run(seq)
            if "PERIOD" not in locus.info:
                return None
            if locus.info["PERIOD"] in [5, 6] and hrun >= locus.info["PERIOD"]:
                return hrun
            return None
        raise ValueError(self.kind)


def new_loc_info(locus_filters: List[LocusFilter]):
    """dumpSTR.py:1262-1268."""
    info = collections.OrderedDict()
    info["totalcalls"] = 0
    info["PASS"] = 0
    info["NO_CALLS_REMAINING"] = 0
    for f in locus_filters:
        info[f.filter_name()] = 0
    return info


def apply_locus_filters(locus: Locus, h, gt, locus_filters, loc_info):
    """dumpSTR.py:917-973 (drop_filtered=False view): returns (filtered, FILTER column text)."""
    names = []
    for filt in locus_filters:
        if filt(locus, h, gt) is None:
            continue
        loc_info[filt.filter_name()] += 1
        names.append(filt.filter_name())
    n_called = np.sum(trh.called_samples(gt))
    if n_called == 0:
        loc_info['NO_CALLS_REMAINING'] += 1
        names.append('NO_CALLS_REMAINING')
    if not names:
        loc_info["PASS"] += 1
        loc_info["totalcalls"] += n_called
        return False, "PASS"
    return True, ";".join(names)


def recompute_info(h, gt, uselength: bool):
    """dumpSTR.py:1307-1336 -> dict(HRUN, HET, HWEP, AC(list), REFAC)."""
    seq = h.full_alleles[0] if h.full_alleles is not None else h.ref_allele
    out = {'HRUN': trh.homopolymer_run(seq)}
    n_alleles = len(h.alt_alleles) + 1
    if np.sum(trh.called_samples(gt)) > 0:
        f = trh.allele_freqs(h, gt, uselength=uselength)
        g = trh.genotype_counts(h, gt, uselength=uselength)
        out['HET'] = stats.heterozygosity(f)
        out['HWEP'] = stats.hardy_weinberg_binomial_test(f, g)
        ac = trh.allele_counts(h, gt, index=True)
        out['AC'] = [int(ac.get(i, 0)) for i in range(1, n_alleles)]
        out['REFAC'] = int(ac.get(0, 0))
    else:
        out['HET'] = -1
        out['HWEP'] = -1
        out['AC'] = [0] * (n_alleles - 1)
        out['REFAC'] = 0
    return out


# ---- logs ------------------------------------------------------------------------
def samplog_text(sample_info, sample_names) -> str:
    """dumpSTR.py:553-588."""
    header = ["sample"] + list(sample_info.keys())
    header[header.index('totaldp')] = 'meanDP'
    lines = ["\t".join(header) + "\n"]
    extra = list(sample_info.values())[2:]
    for i, s in enumerate(sample_names):
        numcalls = sample_info["numcalls"][i]
        row = s + "\t" + str(numcalls) + "\t"
        row += str(sample_info["totaldp"][i] * 1.0 / numcalls) if numcalls > 0 else "0"
        for c in extra:
            row += "\t" + str(c[i])
        lines.append(row + "\n")
    return "".join(lines)


def loclog_text(loc_info) -> str:
    """dumpSTR.py:523-551."""
    keys = [k for k in loc_info.keys() if k != "totalcalls"]
    rate = 0 if loc_info["PASS"] == 0 else float(loc_info["totalcalls"]) / loc_info["PASS"]
    out = "MeanSamplesPerPassingSTR\t%s\n" % rate
    for k in keys:
        out += "FILTER:%s\t%s\n" % (k, loc_info[k])
    return out
