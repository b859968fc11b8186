"""
TEST INFRASTRUCTURE ONLY — restatement of the per-locus statistics of the
reference: ``trtools/utils/utils.py:118-338`` (dict-of-frequency statistics and the
Hardy-Weinberg binomial test) and the ``statSTR`` stat wrappers / row writer
(``trtools/statSTR/statSTR.py:104-426, 544-628``).

Like the reference, every statistic re-derives its allele-frequency dictionary
from the genotype array (one ``np.unique`` per statistic per sample group) — this
is the reference's CPU cost model and is what ``bench.py``'s cpu_baseline times.
"""
from typing import Any, Dict, List, Optional

import numpy as np
import scipy.stats

from . import trh


# ---- utils.py ----------------------------------------------------------------
def validate_allele_freqs(freqs: Dict[Any, float]) -> bool:
    """utils.py:118-140."""
    if len(freqs) == 0:
        return False
    return abs(1 - sum(freqs.values())) <= 0.001


def heterozygosity(freqs) -> float:
    """utils.py:142-175."""
    if not validate_allele_freqs(freqs):
        return np.nan
    return 1 - sum([f ** 2 for f in freqs.values()])


def entropy(freqs) -> float:
    """utils.py:178-212 (scipy.stats.entropy, base 2)."""
    if not validate_allele_freqs(freqs):
        return np.nan
    return float(scipy.stats.entropy(list(freqs.values()), base=2))


def mean(freqs) -> float:
    """utils.py:215-236."""
    if not validate_allele_freqs(freqs):
        return np.nan
    return sum([k * freqs[k] for k in freqs])


def mode(freqs) -> float:
    """utils.py:238-271 (ties -> smallest allele)."""
    if not validate_allele_freqs(freqs):
        return np.nan
    top = max(freqs.values())
    return min(k for k, f in freqs.items() if f == top)


def variance(freqs) -> float:
    """utils.py:273-296."""
    if not validate_allele_freqs(freqs):
        return np.nan
    m = mean(freqs)
    return sum([freqs[k] * (k - m) ** 2 for k in freqs])


def hardy_weinberg_binomial_test(freqs, gcounts) -> float:
    """utils.py:298-338.  ``num_hom`` compares only the first two (sorted) haplotypes;
    a genotype whose allele is not a frequency key (e.g. the -2 ploidy pad) -> NaN."""
    if not validate_allele_freqs(freqs):
        return np.nan
    exp_hom = sum([f ** 2 for f in freqs.values()])
    total = sum(gcounts.values())
    num_hom = 0
    for g, c in gcounts.items():
        if g[0] not in freqs or g[1] not in freqs:
            return np.nan
        if g[0] == g[1]:
            num_hom += c
    return scipy.stats.binomtest(int(num_hom), n=int(total), p=exp_hom).pvalue


# ---- statSTR wrappers (statSTR.py:104-426) -------------------------------------
def stat_thresh(h, gt, groups=(None,)):
    return [trh.max_allele(h, gt, si) for si in groups]


def stat_afreq(h, gt, groups=(None,), count=False, uselength=True) -> List[str]:
    """statSTR.py:128-172."""
    out = []
    for si in groups:
        if count:
            d = trh.allele_counts(h, gt, si, uselength=uselength)
            out.append("." if not d else ",".join("%s:%i" % (a, d.get(a, 0)) for a in sorted(d)))
        else:
            d = trh.allele_freqs(h, gt, si, uselength=uselength)
            out.append("." if not d else ",".join("%s:%.3f" % (a, d.get(a, 0)) for a in sorted(d)))
    return out


def stat_nalleles(h, gt, groups=(None,), thresh=0.01, uselength=True):
    """statSTR.py:174-208."""
    return [len([1 for f in trh.allele_freqs(h, gt, si, uselength=uselength).values() if f >= thresh])
            for si in groups]


def stat_hwep(h, gt, groups=(None,), uselength=True):
    """statSTR.py:210-248."""
    return [hardy_weinberg_binomial_test(trh.allele_freqs(h, gt, si, uselength=uselength),
                                         trh.genotype_counts(h, gt, si, uselength=uselength))
            for si in groups]


def stat_het(h, gt, groups=(None,), uselength=True):
    return [heterozygosity(trh.allele_freqs(h, gt, si, uselength=uselength)) for si in groups]


def stat_entropy(h, gt, groups=(None,), uselength=True):
    return [entropy(trh.allele_freqs(h, gt, si, uselength=uselength)) for si in groups]


def stat_mean(h, gt, groups=(None,)):
    return [mean(trh.allele_freqs(h, gt, si, uselength=True)) for si in groups]


def stat_mode(h, gt, groups=(None,)):
    return [mode(trh.allele_freqs(h, gt, si, uselength=True)) for si in groups]


def stat_var(h, gt, groups=(None,)):
    return [variance(trh.allele_freqs(h, gt, si, uselength=True)) for si in groups]


def stat_numcalled(h, gt, groups=(None,)):
    """statSTR.py:404-426 (length genotypes regardless of --use-length)."""
    return [sum(trh.genotype_counts(h, gt, si).values()) for si in groups]


STAT_ORDER = ("thresh", "afreq", "acount", "nalleles", "hwep", "het", "entropy",
              "mean", "mode", "var", "numcalled")


def locus_stats(h, gt, stats, groups=(None,), uselength=True, nalleles_thresh=0.01) -> Dict[str, list]:
    """All requested statistics of one locus as python values (one entry per group)."""
    out = {}
    if "thresh" in stats: out["thresh"] = stat_thresh(h, gt, groups)
    if "afreq" in stats: out["afreq"] = stat_afreq(h, gt, groups, uselength=uselength)
    if "acount" in stats: out["acount"] = stat_afreq(h, gt, groups, count=True, uselength=uselength)
    if "nalleles" in stats: out["nalleles"] = stat_nalleles(h, gt, groups, nalleles_thresh, uselength)
    if "hwep" in stats: out["hwep"] = stat_hwep(h, gt, groups, uselength)
    if "het" in stats: out["het"] = stat_het(h, gt, groups, uselength)
    if "entropy" in stats: out["entropy"] = stat_entropy(h, gt, groups, uselength)
    if "mean" in stats: out["mean"] = stat_mean(h, gt, groups)
    if "mode" in stats: out["mode"] = stat_mode(h, gt, groups)
    if "var" in stats: out["var"] = stat_var(h, gt, groups)
    if "numcalled" in stats: out["numcalled"] = stat_numcalled(h, gt, groups)
    return out


def format_row(chrom, vcf_pos, h, values: Dict[str, list], precision: int = 3) -> str:
    """statSTR.py:557, 586-628 — one ``.tab`` line (without the newline)."""
    fmt = "\t{:." + str(precision) + "}"

    def num(v):
        return "\tnan" if np.isnan(v) else fmt.format(v)
    row = str(chrom) + "\t" + str(vcf_pos) + "\t" + str(vcf_pos + len(h.ref_allele))
    for key in STAT_ORDER:
        if key not in values:
            continue
        for v in values[key]:
            if key in ("afreq", "acount", "nalleles", "numcalled"):
                row += "\t" + str(v)
            else:
                row += num(v)
    return row
