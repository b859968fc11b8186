"""
Drop-in for reference trtools/associaTR/load_and_filter_genotypes.py: ``load_trs`` yields, per locus,
``(gts, unique_alleles, chrom, pos, called_samples_filter, locus_filtered, locus_details)`` after a first
yield of the detail-field names.  Records are harmonized in GPU blocks; allele frequencies come from the
scan kernel; the float64 ``gts`` array is materialised only at this API edge (associaTR's own GPU path,
``associaTR.perform_gwas``, never materialises it).  The ``--beagle-dosages`` branch is outside the
accelerated path.
"""
import sys
from typing import Optional, Union

import numpy as np

from . import tr_harmonizer as trh
from . import utils

cyvcf2 = utils.cyvcf2

allele_len_precision = 2
allele_frequency_precision = 2
dosage_precision = 2
r2_precision = 2


def dict_str(d):
    """reference lafg.py:23-35."""
    out = '{'
    first = True
    for key in sorted(d.keys()):
        if not first:
            out += ', '
        first = False
        out += '{}: {}'.format(repr(str(key)), repr(d[key]))
    out += '}'
    return out.replace("'", '"').replace('(', '[').replace(')', ']').replace('nan', '"NaN"')


def clean_len_alleles(d):
    """reference lafg.py:37-45."""
    new_d = {}
    for key, val in d.items():
        new_key = round(key, allele_len_precision)
        if new_key not in new_d:
            new_d[new_key] = val
        else:
            new_d[new_key] += val
    return new_d


def round_vals(d, precision):
    return {key: round(val, precision) for key, val in d.items()}


def locus_filter_reason(allele_frequency, n_samples, non_major_cutoff, beagle_dosages=False):
    """reference lafg.py:228-238."""
    if len(allele_frequency) == 0:
        return 'No called samples'
    if len(allele_frequency) == 1:
        return 'Only one called allele'
    af = list(allele_frequency.values())
    af.pop(int(np.argmax(af)))
    if np.sum(af) * n_samples * 2 < non_major_cutoff:
        return 'non-major allele {}<{}'.format("dosage" if beagle_dosages else "count", non_major_cutoff)
    return None


def load_trs(vcf_fname: str, samples: Union[np.ndarray, slice], region: Optional[str] = None,
             non_major_cutoff: float = 20, beagle_dosages: bool = False, vcftype: Optional[str] = None,
             _imputed_ukb_strs_paper_period_check: bool = False, block_size: int = 512):
    """reference lafg.py:60-259 (non-dosage branch)."""
    if beagle_dosages:
        raise NotImplementedError("--beagle-dosages is outside the accelerated path of trtools_b200")
    vcf = cyvcf2.VCF(vcf_fname)
    inferred = trh.InferVCFType(vcf, vcftype if vcftype else 'auto')
    region_start = None
    if region is not None:
        region_start = int(region.split(':')[1].split('-')[0])
        vcf = vcf(region)
    yield ['motif', 'period', 'ref_len', 'allele_frequency']

    def wanted():
        for record in vcf:
            if region_start is not None and record.POS < region_start:
                continue
            if _imputed_ukb_strs_paper_period_check and record.INFO.get('PERIOD') is None:
                continue
            yield record

    class _Iter:
        raw_header = vcf.raw_header

        def __init__(self):
            self.it = wanted()

        def __iter__(self):
            return self

        def __next__(self):
            return next(self.it)

    harmonizer = trh.TRRecordHarmonizer(_Iter(), inferred.name, block_size=block_size)
    for trrecord in harmonizer:
        called = trrecord.GetCalledSamples()
        if isinstance(samples, slice):
            assert samples == slice(None)
            called_samples_filter = called
            curr_samples = called
        else:
            called_samples_filter = called[samples]
            curr_samples = samples & called
        n_samples = int(np.sum(curr_samples))
        len_alleles = [round(x, allele_len_precision) for x in [trrecord.ref_allele_length] + trrecord.alt_allele_lengths]
        gts = trrecord.GetLengthGenotypes()[curr_samples, :-1]
        allele_frequency = clean_len_alleles(trrecord.GetAlleleFreqs(curr_samples))
        locus_details = [trrecord.motif, str(len(trrecord.motif)),
                         str(round(trrecord.ref_allele_length, allele_len_precision)),
                         dict_str({key: '{:.2g}'.format(val) for key, val in allele_frequency.items()})]
        reason = locus_filter_reason(allele_frequency, n_samples, non_major_cutoff)
        yield (None if reason else gts, np.unique(len_alleles), trrecord.chrom, trrecord.pos,
               called_samples_filter, reason, locus_details)
