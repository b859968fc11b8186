"""
Drop-in for reference trtools/associaTR/load_and_filter_genotypes.py: ``load_trs`` yields, per locus,
``(gts, unique_alleles, chrom, pos, called_samples_filter, locus_filtered, locus_details)`` after a first
yield of the detail-field names.  Records are harmonized in GPU blocks; allele frequencies come from the
scan kernel; the float64 ``gts`` array is materialised only at this API edge (associaTR's own GPU path,
``associaTR.perform_gwas``, never materialises it).  With ``beagle_dosages`` the per-length dosage arrays are
materialised here from the records' AP1 / AP2 fields for the caller, while the frequencies and the dosage r2 columns
come from the dosage kernel (``trt_assoc_dosage_ols``).
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
    """reference lafg.py:60-259."""
    vcf = cyvcf2.VCF(vcf_fname)
    inferred = trh.InferVCFType(vcf, vcftype if vcftype else 'auto')
    region_start = None
    if region is not None:
        region_start = int(region.split(':')[1].split('-')[0])
        vcf = vcf(region)
    deets = ['motif', 'period', 'ref_len', 'allele_frequency']
    if beagle_dosages:
        deets.extend(['dosage_estimated_r2_per_length_allele', 'r2_length_dosages_vs_best_guess_lengths'])
    yield deets

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
    first = True
    for trrecord in harmonizer:
        if first and beagle_dosages and "AP1" not in (trrecord.vcfrecord.FORMAT or []):
            print("--beagle-dosages specified, missing required field AP1 for the TR")
            if "GP" in (trrecord.vcfrecord.FORMAT or []):
                print("We could support the GP field, but currently only support the AP fields")
            print("Erroring out")
            sys.exit(1)
        first = False
        if beagle_dosages:
            yield _dosage_locus(trrecord, samples, non_major_cutoff)
            continue
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


def dosage_arrays(trrecord, curr, len_alleles):
    """Per-length haplotype dosages [n][2] of the samples ``curr`` from AP1 / AP2 (reference lafg.py:175-186)."""
    n_samples = int(np.sum(curr))
    gts = {_len: np.zeros((n_samples, 2)) for _len in np.unique(len_alleles)}
    for p in (1, 2):
        ap = trrecord.format['AP{}'.format(p)]
        gts[len_alleles[0]][:, (p - 1)] += np.maximum(0, 1 - np.sum(ap[curr, :], axis=1))
        for i in range(ap.shape[1]):
            gts[len_alleles[i + 1]][:, (p - 1)] += ap[curr, i]
    return gts


def flat_locus_length_r2(trrecord, curr, gts):
    """``r2_length_dosages_vs_best_guess_lengths`` of a locus whose best-guess lengths are all equal.  There the
    reference's np.corrcoef (lafg.py:208-213) divides rounding residue by rounding residue — the mean of n copies of
    a length need not be that length — so the printed value (nan, 0.0 or 1.0) is a property of numpy's summation and
    not of the moments the device accumulates.  Such loci are filtered ('Only one called allele'); their one detail
    column is computed here with the reference's own expression."""
    best_guesses = trrecord.GetLengthGenotypes()[curr, :-1]
    with np.errstate(divide='ignore', invalid='ignore'):
        return np.corrcoef(best_guesses.flatten(),
                           np.add.reduce([len_ * dosages for len_, dosages in gts.items()]).flatten())[0, 1] ** 2


def _dosage_locus(trrecord, samples, non_major_cutoff):
    """One locus of the ``beagle_dosages`` branch (reference lafg.py:160-259) for the generator protocol.  The dict of
    per-length haplotype dosages is what the consumer regresses on, so it is built here from the record's AP fields;
    the statistics printed beside it are the device's (one ``trt_assoc_dosage_ols`` pass per block, cached)."""
    from . import associaTR as _assoc
    blk, l = trrecord._blk, trrecord._l
    called = trrecord.GetCalledSamples()
    if isinstance(samples, slice):
        assert samples == slice(None)
        called_samples_filter, curr = called, called
        design_idx = np.arange(blk.S, dtype=np.int32)
    else:
        called_samples_filter, curr = called[samples], samples & called
        design_idx = np.nonzero(samples)[0].astype(np.int32)
    n_samples = int(np.sum(curr))
    len_alleles = [round(x, allele_len_precision) for x in [trrecord.ref_allele_length] + trrecord.alt_allele_lengths]
    gts = dosage_arrays(trrecord, curr, len_alleles)
    cache = blk.__dict__.setdefault("_dosage_stats", {})
    key = design_idx.tobytes()
    if key not in cache:
        blk.ensure_ap()
        meta = _assoc.dosage_classes(blk)
        # a design of the requested samples with a bare intercept: only the per-class sums are read here
        covars = np.ones((len(design_idx), 2))
        blk.ctx.assoc_set_design(covars, np.zeros(len(design_idx)), design_idx)
        cache.clear()
        cache[key] = (meta, blk.ctx.assoc_dosage_ols(*meta))
    (cls, len_round, _), res = cache[key]
    sl = blk.allele_slice(l)
    cs = res["class_stats"][sl]
    lr = len_round[sl]
    with np.errstate(divide='ignore', invalid='ignore'):
        allele_frequency = {float(lr[j]): np.float64(cs[j, 0]) / (2 * n_samples)
                            for j in sorted([j for j in range(len(lr)) if cls[sl.start + j] == j], key=lambda j: lr[j])}
    r2 = {}
    for j in range(len(lr)):
        if float(lr[j]) not in r2:
            c = int(cls[sl.start + j])
            r2[float(lr[j])] = _assoc._r2(2 * n_samples, cs[c, 2], cs[c, 2], cs[c, 0], cs[c, 1], cs[c, 3])
    ls = res["length_stats"][l]
    length_r2 = _assoc._r2(2 * n_samples, ls[0], ls[1], ls[2], ls[3], ls[4])
    if n_samples > 0 and any(cs[j, 2] == 2 * n_samples for j in range(len(lr))):
        length_r2 = flat_locus_length_r2(trrecord, curr, gts)
    locus_details = [trrecord.motif, str(len(trrecord.motif)), str(round(trrecord.ref_allele_length, allele_len_precision)),
                     dict_str({k: '{:.2g}'.format(v) for k, v in allele_frequency.items()}),
                     dict_str(round_vals(r2, r2_precision)), str(round(length_r2, r2_precision))]
    reason = locus_filter_reason(allele_frequency, n_samples, non_major_cutoff, True)
    return (None if reason else gts, np.unique(len_alleles), trrecord.chrom, trrecord.pos, called_samples_filter, reason,
            locus_details)
