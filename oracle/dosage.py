"""
TEST INFRASTRUCTURE ONLY — restatement of the Beagle allele-probability dosage path of the reference:

* ``TRRecord.GetDosages`` (``trtools/utils/tr_harmonizer.py:1098-1208``): ``beagleap`` / ``beagleap_norm`` (and, for
  completeness of the dtype handling, ``bestguess`` / ``bestguess_norm``);
* the ``--beagle-dosages`` branch of ``load_trs`` (``trtools/associaTR/load_and_filter_genotypes.py:175-214, 216-259``)
  and the regression on summed dosages (``trtools/associaTR/associaTR.py:266-291``).

dtype notes that matter for parity (numpy >= 2 promotion rules): ``AP1`` / ``AP2`` are float32 ``[S, A-1]``;
``np.sum(ap, axis=1)``, ``1 - sum`` and the product with the (python float) reference length stay float32, the dot product
with the alt lengths is float64, the four terms are added in float64 and cast to float32.

Pinned by tests/test_oracle_golden.py against the unmodified reference's outputs on its own Beagle fixtures
(tests/golden/dosage.json, generator tests/golden/make_golden.py section ``dosage``).
"""
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np

from . import trh
from .assoc import ALLELE_LEN_PRECISION, AssocRow, Design, clean_len_alleles, dict_str, ols_fit
from .records import Locus

R2_PRECISION = 2      # load_and_filter_genotypes.py:21


class DosageError(ValueError):
    pass


def beagle_dosages(h: trh.Harmonized, locus: Locus, norm: bool = False, strict: bool = True) -> Optional[np.ndarray]:
    """tr_harmonizer.py:1154-1208 -> float32 [S] (NaN array when not strict and the AP fields are invalid)."""
    gt = locus.gt
    if gt is None or gt.shape[0] == 0:
        return None
    n = gt.shape[0]

    def fail(msg):
        if strict:
            raise DosageError(msg)
        return np.array([np.nan] * n)

    if "AP1" not in locus.fmt or "AP2" not in locus.fmt:
        return fail("Requested Beagle dosages for record at {}:{} but AP1/AP2 fields not found.".format(locus.chrom, h.pos))
    ap1 = locus.fmt["AP1"]
    ap2 = locus.fmt["AP2"]
    ref1 = np.clip(1 - np.sum(ap1, axis=1), 0, 1)
    ref2 = np.clip(1 - np.sum(ap2, axis=1), 0, 1)
    if np.any(np.sum(ap1, axis=1) > 1.1) or np.any(np.sum(ap2, axis=1) > 1.1):
        return fail("{}:{} AP1 or AP2 field summing to more than 1 detected".format(locus.chrom, h.pos))
    if np.any(ap1 < 0) or np.any(ap2 < 0):
        return fail("Negative AP1 or AP2 fields detected")
    if len(h.alt_allele_lengths) > 0:
        max_alt = max(h.alt_allele_lengths)
        h1 = np.clip(np.dot(ap1, h.alt_allele_lengths), 0, max_alt)
        h2 = np.clip(np.dot(ap2, h.alt_allele_lengths), 0, max_alt)
    else:
        h1 = h2 = 0
    unnorm = (h1 + h2 + ref1 * h.ref_allele_length + ref2 * h.ref_allele_length).astype(np.float32)
    if not norm:
        return unnorm
    if h.min_allele_length == h.max_allele_length:
        return np.zeros(n, dtype=np.float32)
    d = (unnorm - 2 * h.min_allele_length) / (h.max_allele_length - h.min_allele_length)
    if np.any(d >= 2.1) or np.any(d <= -0.1):
        return fail("{}:{} Error normalizing dosages: value >=2.1 or <=-0.1 detected".format(locus.chrom, h.pos))
    return np.clip(d, 0, 2)


@dataclass
class LoadedDosageLocus:
    gts: Optional[Dict[float, np.ndarray]]   # rounded length -> float64 [n_called, 2] per-haplotype dosages
    unique_alleles: np.ndarray
    chrom: str
    pos: int
    called_samples_filter: np.ndarray
    filter_reason: Optional[str]
    details: List[str]
    n_samples: int
    allele_frequency: Dict[float, float]
    allele_dosage_r2: Dict[float, float]
    length_r2: float


def load_dosage_locus(locus: Locus, h: trh.Harmonized, samples, non_major_cutoff: float = 20) -> LoadedDosageLocus:
    """One iteration of ``load_trs`` with ``beagle_dosages=True`` (lafg.py:157-259)."""
    called = trh.called_samples(locus.gt)
    if isinstance(samples, slice):
        called_filter, curr = called, called
    else:
        called_filter, curr = called[samples], samples & called
    n_samples = int(np.sum(curr))
    len_alleles = [round(x, ALLELE_LEN_PRECISION) for x in h.allele_lengths]
    gts = {ln: np.zeros((n_samples, 2)) for ln in np.unique(len_alleles)}
    for p in (1, 2):
        ap = locus.fmt['AP{}'.format(p)]
        gts[len_alleles[0]][:, p - 1] += np.maximum(0, 1 - np.sum(ap[curr, :], axis=1))
        for i in range(ap.shape[1]):
            gts[len_alleles[i + 1]][:, p - 1] += ap[curr, i]
    with np.errstate(divide='ignore', invalid='ignore'):
        freq = {ln: np.sum(gts[ln]) / (2 * n_samples) for ln in gts}
        best = trh.length_genotypes(h, locus.gt)[curr, :-1]
        rounded_best = np.around(best, ALLELE_LEN_PRECISION)
        r2 = {}
        for length in len_alleles:
            if length in r2:
                continue
            calls = rounded_best == length
            r2[length] = np.corrcoef(calls.reshape(-1), gts[length].reshape(-1))[0, 1] ** 2
        length_r2 = np.corrcoef(best.flatten(), np.add.reduce([ln * d for ln, d in gts.items()]).flatten())[0, 1] ** 2
    details = [h.motif, str(len(h.motif)), str(round(h.ref_allele_length, ALLELE_LEN_PRECISION)),
               dict_str({k: '{:.2g}'.format(v) for k, v in freq.items()}),
               dict_str({k: round(v, R2_PRECISION) for k, v in r2.items()}), str(round(length_r2, R2_PRECISION))]
    if len(freq) == 0:
        reason = 'No called samples'
    elif len(freq) == 1:
        reason = 'Only one called allele'
    else:
        af = list(freq.values())
        af.pop(int(np.argmax(af)))
        reason = ('non-major allele dosage<{}'.format(non_major_cutoff)
                  if np.sum(af) * n_samples * 2 < non_major_cutoff else None)
    return LoadedDosageLocus(gts=None if reason else gts, unique_alleles=np.unique(len_alleles), chrom=locus.chrom,
                             pos=h.pos, called_samples_filter=called_filter, filter_reason=reason, details=details,
                             n_samples=n_samples, allele_frequency=freq, allele_dosage_r2=r2, length_r2=float(length_r2))


def regress_dosage_locus(loaded: LoadedDosageLocus, design: Design) -> AssocRow:
    """associaTR.py:246-291 with ``beagle_dosages`` (summed dosage = sum over lengths of length x both haplotypes)."""
    covars = design.covars
    covars[:, 0] = np.nan
    csf = loaded.called_samples_filter
    alleles = ','.join(list(loaded.unique_alleles.astype(str)))
    n_tested = int(np.sum(csf))
    reason = loaded.filter_reason
    if not reason and covars.shape[1] >= n_tested:
        reason = 'n covars >= n samples'
    if reason:
        return AssocRow(loaded.chrom, loaded.pos, alleles, n_tested, reason, np.nan, np.nan, np.nan, np.nan, loaded.details)
    summed = np.sum([ln * np.sum(d, axis=1) for ln, d in loaded.gts.items()], axis=0)
    std = np.std(summed)
    with np.errstate(divide='ignore', invalid='ignore'):
        summed = (summed - np.mean(summed)) / np.std(summed)
    covars[csf, 0] = summed
    res = ols_fit(design.outcome[csf], covars[csf, :])
    with np.errstate(divide='ignore', invalid='ignore'):
        return AssocRow(loaded.chrom, loaded.pos, alleles, n_tested, False, res.pvalue, res.coef / std * design.pheno_std,
                        res.se / std * design.pheno_std, res.rsquared, loaded.details)


DOSAGE_HEADER_FIELDS = ['motif', 'period', 'ref_len', 'allele_frequency', 'dosage_estimated_r2_per_length_allele',
                        'r2_length_dosages_vs_best_guess_lengths']
