"""
TEST INFRASTRUCTURE ONLY — restatement of the reductions the qcSTR and compareSTR consumers take over the harmonized
records:

* the record loop of ``qcSTR.main`` (``trtools/qcSTR/qcSTR.py:523-570``): calls per sample / per locus and the quality sums;
* ``compareSTR.UpdateComparisonResults`` (``trtools/compareSTR/compareSTR.py:508-643``, no FORMAT stratification): per
  locus and per sample concordance counts and the sums behind the overall R^2.

Pinned by tests/test_oracle_golden.py against tests/golden/reductions.json (the unmodified reference on
many_samples.vcf.gz and on its own GangSTR pair; generator tests/golden/make_golden.py ``reduce``).
"""
from typing import Dict, List, Optional

import numpy as np

from . import trh
from .records import Locus


def qc_reduce(loci: List[Locus], sample_index: np.ndarray, quality_key: Optional[str], ignore_no_call: bool) -> Dict:
    """qcSTR.py:523-556 over a list of loci."""
    n = int(np.sum(sample_index))
    sample_calls = np.zeros(n)
    per_sample_total_qual = np.zeros(n)
    locus_calls, per_locus = [], []
    for l in loci:
        idx_gts = trh.genotype_indices(l.gt)[sample_index, :-1]
        calls = ~np.all(idx_gts == -1, axis=1)                       # :532-533
        sample_calls += calls
        locus_calls.append(int(np.sum(calls)))
        if quality_key is None:
            continue
        q = np.array(l.fmt[quality_key], dtype=np.float32).reshape(l.gt.shape[0], -1)[sample_index, :1].copy()
        q[~calls] = np.nan
        with np.errstate(invalid='ignore'):
            if not ignore_no_call:
                q[np.isnan(q)] = 0
                per_sample_total_qual += q.reshape(-1)
                per_locus.append(float(np.mean(q)))
            else:
                qi = ~np.isnan(q)
                per_sample_total_qual[qi.reshape(-1)] += q[qi].reshape(-1)
                per_locus.append(float(np.mean(q[qi])) if qi.any() else float('nan'))
    return dict(sample_calls=sample_calls, locus_calls=locus_calls, per_sample_total_qual=per_sample_total_qual,
                per_locus=per_locus)


class CompareError(ValueError):
    pass


def compare_locus(l1: Locus, h1: trh.Harmonized, l2: Locus, h2: trh.Harmonized, sample_idxs, ignore_phasing: bool):
    """compareSTR.py:545-612 for one pair of records -> None when nothing is called in both, else a dict."""
    both = trh.called_samples(l1.gt)[sample_idxs[0]] & trh.called_samples(l2.gt)[sample_idxs[1]]
    numcalls = int(np.sum(both))
    if numcalls == 0:
        return None
    i1, i2 = sample_idxs[0][both], sample_idxs[1][both]
    if not np.all(trh.sample_ploidies(l1.gt)[i1] == trh.sample_ploidies(l2.gt)[i2]):
        raise CompareError("Found sample(s) of different ploidy at %s:%s" % (l1.chrom, h1.pos))
    s1, s2 = trh.string_genotypes(h1, l1.gt)[i1, :], trh.string_genotypes(h2, l2.gt)[i2, :]
    if ignore_phasing:
        all_unphased = True
    else:
        unphased = (s1[:, -1] == '0') & (s2[:, -1] == '0')
        all_unphased = bool(np.all(unphased))
        if not (all_unphased or np.all(~unphased)):
            raise CompareError("Found sample(s) with different phasedness at %s:%s" % (l1.chrom, h1.pos))
    s1, s2 = s1[:, :-1], s2[:, :-1]
    g1, g2 = trh.length_genotypes(h1, l1.gt)[i1, :-1], trh.length_genotypes(h2, l2.gt)[i2, :-1]
    if all_unphased:
        s1, s2 = np.sort(s1, axis=1), np.sort(s2, axis=1)
        g1, g2 = np.sort(g1, axis=1), np.sort(g2, axis=1)
    conc_seq = np.all(s1 == s2, axis=1)
    conc_len = np.all(g1 == g2, axis=1)
    reflen = len(h1.ref_allele) / len(h1.motif)
    d1, d2 = np.sum(g1 - reflen, axis=1), np.sum(g2 - reflen, axis=1)
    return dict(both=both, numcalls=numcalls, conc_seq=conc_seq, conc_len=conc_len,
                sums=[float(np.sum(d1)), float(np.sum(d2)), float(np.sum(d1 ** 2)), float(np.sum(d1 * d2)), float(np.sum(d2 ** 2))])
