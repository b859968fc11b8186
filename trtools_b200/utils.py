"""
Mirror of the parts of ``trtools.utils.utils`` (reference trtools/utils/utils.py) that the hot path's
callers touch: the VCF reader loader and the dict-of-frequency helpers that user scripts call on
the (tiny) dictionaries returned by ``TRRecord.GetAlleleFreqs``.  The per-locus statistics that
statSTR/dumpSTR/associaTR print are NOT computed here — they come from the CUDA epilogue kernel
(csrc/trt_stats.cu); these helpers exist only so that code written against the reference API
(`utils.GetHeterozygosity(record.GetAlleleFreqs())`) keeps working.
"""
import argparse
import itertools
import math
import os
from typing import Any, Dict, List, Optional, Set

import numpy as np

from . import common

try:  # the real reader if it is installed, else the text reader of this package
    import cyvcf2  # type: ignore
except ImportError:  # pragma: no cover
    from . import cyvcf2_compat as cyvcf2

nucToNumber = {"A": 0, "C": 1, "G": 2, "T": 3}


def LoadSingleReader(vcf_loc: str, checkgz: bool = True, lazy: bool = False, samples: Set[str] = None):
    """reference utils.py:19-68."""
    if not os.path.exists(vcf_loc) or os.path.isdir(vcf_loc):
        common.WARNING("Could not find VCF file %s" % vcf_loc)
        return None
    if checkgz:
        if not vcf_loc.endswith(".vcf.gz") and not vcf_loc.endswith(".vcf.bgz"):
            common.WARNING("Make sure %s is bgzipped and indexed" % vcf_loc)
            return None
        if not os.path.isfile(vcf_loc + ".tbi"):
            common.WARNING("Could not find VCF index %s.tbi" % vcf_loc)
            return None
    if samples is not None:
        if not isinstance(samples, set):
            common.WARNING("Samples cannot be loaded in a particular order. Order will be ignored")
        samples = list(samples)
    try:
        return cyvcf2.VCF(vcf_loc, lazy=lazy, samples=samples)
    except OSError:
        common.WARNING("Could not open VCF file %s. Is it really VCF?" % vcf_loc)
        return None


def LoadReaders(vcf_locs: List[str], checkgz: bool = True):
    """reference utils.py:70-96."""
    readers = []
    for f in vcf_locs:
        rdr = LoadSingleReader(f, checkgz)
        if rdr is None:
            return None
        readers.append(rdr)
    return readers


def GetContigs(vcf) -> List[str]:
    """reference utils.py:98-116."""
    return [h['ID'] for h in vcf.header_iter() if h['HeaderType'].lower() == 'contig']


def ValidateAlleleFreqs(allele_freqs) -> bool:
    """reference utils.py:118-140."""
    if len(allele_freqs.keys()) == 0:
        return False
    return abs(1 - sum(allele_freqs.values())) <= 0.001


def GetHeterozygosity(allele_freqs) -> float:
    """reference utils.py:142-175."""
    if not ValidateAlleleFreqs(allele_freqs):
        return np.nan
    return 1 - sum([freq ** 2 for freq in allele_freqs.values()])


def GetEntropy(allele_freqs: Dict[Any, float]) -> float:
    """reference utils.py:178-212 (bit entropy of the renormalised frequencies)."""
    if not ValidateAlleleFreqs(allele_freqs):
        return np.nan
    vals = np.array(list(allele_freqs.values()), dtype=float)
    vals = vals / vals.sum()
    nz = vals[vals > 0]
    return float(np.sum(-nz * np.log(nz)) / math.log(2))


def GetMean(allele_freqs) -> float:
    """reference utils.py:215-236."""
    if not ValidateAlleleFreqs(allele_freqs):
        return np.nan
    return sum([key * allele_freqs[key] for key in allele_freqs])


def GetMode(allele_freqs) -> float:
    """reference utils.py:238-271."""
    if not ValidateAlleleFreqs(allele_freqs):
        return np.nan
    top = max(allele_freqs.values())
    return min(k for k, f in allele_freqs.items() if f == top)


def GetVariance(allele_freqs) -> float:
    """reference utils.py:273-296."""
    if not ValidateAlleleFreqs(allele_freqs):
        return np.nan
    mean = GetMean(allele_freqs)
    return sum([allele_freqs[key] * (key - mean) ** 2 for key in allele_freqs.keys()])


def GetHardyWeinbergBinomialTest(allele_freqs, genotype_counts) -> float:
    """reference utils.py:298-338 (user-script helper on dicts; statSTR/dumpSTR use the device test)."""
    import scipy.stats
    if not ValidateAlleleFreqs(allele_freqs):
        return np.nan
    exp_hom_frac = sum([val ** 2 for val in allele_freqs.values()])
    total_samples = sum(genotype_counts.values())
    num_hom = 0
    for gt in genotype_counts:
        if gt[0] not in allele_freqs.keys():
            return np.nan
        if gt[1] not in allele_freqs.keys():
            return np.nan
        if gt[0] == gt[1]:
            num_hom += genotype_counts[gt]
    return scipy.stats.binomtest(int(num_hom), n=int(total_samples), p=exp_hom_frac).pvalue


def GetHomopolymerRun(seq: str) -> int:
    """reference utils.py:340-360."""
    if len(seq) == 0:
        return 0
    seq = seq.upper()
    return max(len(list(y)) for (c, y) in itertools.groupby(seq))


def GetCanonicalOneStrand(repseq: str) -> str:
    """reference utils.py:396-427."""
    repseq = repseq.upper()
    size = len(repseq)
    canonical = repseq
    for i in range(size):
        newseq = repseq[size - i:] + repseq[0:size - i]
        for j in range(size):
            if nucToNumber[newseq[j]] < nucToNumber[canonical[j]]:
                canonical = newseq
            elif nucToNumber[newseq[j]] > nucToNumber[canonical[j]]:
                break
    return canonical


def FabricateAllele(motif: str, length: float) -> str:
    """reference utils.py:566-602."""
    fab = math.floor(length) * motif
    idx = 0
    while (len(fab) + 1) / len(motif) < length:
        fab += motif[idx]
        idx += 1
    return fab


class ArgumentDefaultsHelpFormatter(argparse.HelpFormatter):  # pragma: no cover
    """reference utils.py:605-626."""

    def _get_help_string(self, action):
        help = action.help
        if '%(default)' not in action.help:
            if action.default is not argparse.SUPPRESS and action.default is not None:
                defaulting_nargs = [argparse.OPTIONAL, argparse.ZERO_OR_MORE]
                if action.option_strings or action.nargs in defaulting_nargs:
                    help += ' (default: %(default)s)'
        return help
