"""
TEST INFRASTRUCTURE ONLY — restatement of the associaTR hot path of the reference:
the genotype loader/filter (``trtools/associaTR/load_and_filter_genotypes.py:137-259``,
non-dosage branch), covariate preparation (``trtools/associaTR/associaTR.py:140-194``) and the
per-locus regression + output row (``associaTR.py:246-304``).

``statsmodels.OLS(...).fit()`` (third-party, pinned 0.14.x in the reference's pyproject.toml,
absent from /root/reference) is restated from its published algorithm in :func:`ols_fit` and
pinned against the plink2 ``--glm`` goldens the reference's own tests use.
"""
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import scipy.stats

from . import trh
from .records import Locus

ALLELE_LEN_PRECISION = 2      # load_and_filter_genotypes.py:15


def dict_str(d) -> str:
    """load_and_filter_genotypes.py:23-35."""
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
    """load_and_filter_genotypes.py:37-45."""
    out = {}
    for k, v in d.items():
        nk = round(k, ALLELE_LEN_PRECISION)
        out[nk] = out.get(nk, 0) + v if nk in out else v
    return out


@dataclass
class LoadedLocus:
    gts: Optional[np.ndarray]            # float64 [n_called, P] or None when filtered
    unique_alleles: np.ndarray
    chrom: str
    pos: int
    called_samples_filter: np.ndarray
    filter_reason: Optional[str]
    details: List[str]
    n_samples: int


def load_locus(locus: Locus, h: trh.Harmonized, samples, non_major_cutoff: float = 20) -> LoadedLocus:
    """One iteration of ``load_trs`` (load_and_filter_genotypes.py:157-259), non-dosage branch.
    ``samples``: bool mask over all samples, or ``slice(None)``."""
    called = trh.called_samples(locus.gt)
    if isinstance(samples, slice):
        called_filter = called
        curr = called
    else:
        called_filter = called[samples]
        curr = samples & called
    n_samples = int(np.sum(curr))
    len_alleles = [round(x, ALLELE_LEN_PRECISION) for x in h.allele_lengths]
    gts = trh.length_genotypes(h, locus.gt)[curr, :-1]
    freq = clean_len_alleles(trh.allele_freqs(h, locus.gt, curr))
    details = [h.motif, str(len(h.motif)), str(round(h.ref_allele_length, ALLELE_LEN_PRECISION)),
               dict_str({k: '{:.2g}'.format(v) for k, v in freq.items()})]
    if len(freq) == 0:
        reason = 'No called samples'
    elif len(freq) == 1:
        reason = 'Only one called allele'
    else:
        af = list(freq.values())
        af.pop(int(np.argmax(af)))
        if np.sum(af) * n_samples * 2 < non_major_cutoff:
            reason = 'non-major allele count<{}'.format(non_major_cutoff)
        else:
            reason = None
    return LoadedLocus(gts=None if reason else gts, unique_alleles=np.unique(len_alleles),
                       chrom=locus.chrom, pos=h.pos, called_samples_filter=called_filter,
                       filter_reason=reason, details=details, n_samples=n_samples)


@dataclass
class Design:
    covars: np.ndarray          # float64 [n, K+2]; col 0 = genotype slot, col 1 = intercept
    outcome: np.ndarray         # float64 [n]
    pheno_std: float
    sample_filter: np.ndarray   # bool over all VCF samples


def prepare_design(trait_arrays: List[np.ndarray], n_vcf_samples: int,
                   sample_mask: Optional[np.ndarray] = None) -> Design:
    """associaTR.py:150-194, ``--same-samples`` branch: hstack, drop NaN rows, standardise every
    column over the retained samples, then outcome = col 1 and col 1 <- 1 (intercept)."""
    covars = np.hstack([np.full((trait_arrays[0].shape[0], 1), -1), *trait_arrays]).astype(float)
    keep = np.ones(n_vcf_samples, dtype=bool) if sample_mask is None else np.array(sample_mask, dtype=bool)
    keep = keep & ~np.any(np.isnan(covars), axis=1)
    covars = covars[keep, :]
    pheno_std = np.std(covars[:, 1])
    with np.errstate(divide='ignore', invalid='ignore'):
        covars = (covars - np.mean(covars, axis=0)) / np.std(covars, axis=0)
    outcome = covars[:, 1].copy()
    covars[:, 1] = 1
    return Design(covars=covars, outcome=outcome, pheno_std=pheno_std, sample_filter=keep)


@dataclass
class OLSResult:
    pvalue: float
    coef: float
    se: float
    rsquared: float


def ols_fit(y: np.ndarray, X: np.ndarray) -> OLSResult:
    """statsmodels ``OLS(y, X, missing='drop').fit()``: pinv parameters,
    ``pinv pinv^T * ssr/(n-rank)`` covariance, two-sided t p-value, centred R^2 when a constant
    column is present.  Index 0 (the genotype column) is what associaTR.py:287-290 reads."""
    keep = ~(np.isnan(y) | np.any(np.isnan(X), axis=1))
    y, X = y[keep], X[keep]
    n = X.shape[0]
    if n == 0:
        return OLSResult(np.nan, np.nan, np.nan, np.nan)
    pinv = np.linalg.pinv(X, rcond=1e-15)
    sv = np.linalg.svd(X, compute_uv=False)
    beta = pinv @ y
    rank = np.linalg.matrix_rank(np.diag(sv))
    df = n - rank
    resid = y - X @ beta
    ssr = float(resid @ resid)
    with np.errstate(divide='ignore', invalid='ignore'):
        cov = (pinv @ pinv.T) * (ssr / df)
        se = np.sqrt(np.diag(cov))
        t = beta / se
    p = 2 * scipy.stats.t.sf(np.abs(t), df)
    ptp = np.ptp(X, axis=0)
    has_const = bool(np.any((ptp == 0) & np.all(X != 0, axis=0)))
    tss = float(np.sum((y - y.mean()) ** 2)) if has_const else float(y @ y)
    return OLSResult(float(p[0]), float(beta[0]), float(se[0]), 1 - ssr / tss)


@dataclass
class AssocRow:
    chrom: str
    pos: int
    alleles: str
    n_samples_tested: int
    locus_filtered: object       # False or reason string
    p: float
    coef: float
    se: float
    r2: float
    details: List[str]

    def to_text(self) -> str:
        """associaTR.py:252-304 (no plotting columns)."""
        head = "{}\t{}\t{}\t{}\t".format(self.chrom, self.pos, self.alleles, self.n_samples_tested)
        if self.locus_filtered:
            return head + '{}\tnan\tnan\tnan\tnan\t'.format(self.locus_filtered) + '\t'.join(self.details) + '\n'
        return (head + 'False\t' + "{:.2e}\t{}\t{}\t{}\t".format(self.p, self.coef, self.se, self.r2)
                + '\t'.join(self.details) + '\n')


def regress_locus(loaded: LoadedLocus, design: Design) -> AssocRow:
    """associaTR.py:246-291."""
    covars = design.covars
    covars[:, 0] = np.nan
    csf = loaded.called_samples_filter
    alleles = ','.join(list(loaded.unique_alleles.astype(str)))
    n_tested = int(np.sum(csf))
    reason = loaded.filter_reason
    if not reason and covars.shape[1] >= n_tested:
        reason = 'n covars >= n samples'
    if reason:
        return AssocRow(loaded.chrom, loaded.pos, alleles, n_tested, reason,
                        np.nan, np.nan, np.nan, np.nan, loaded.details)
    summed = np.sum(loaded.gts, axis=1)
    std = np.std(summed)
    with np.errstate(divide='ignore', invalid='ignore'):
        summed = (summed - np.mean(summed)) / np.std(summed)
    covars[csf, 0] = summed
    res = ols_fit(design.outcome[csf], covars[csf, :])
    with np.errstate(divide='ignore', invalid='ignore'):
        return AssocRow(loaded.chrom, loaded.pos, alleles, n_tested, False,
                        res.pvalue, res.coef / std * design.pheno_std, res.se / std * design.pheno_std,
                        res.rsquared, loaded.details)


HEADER_FIELDS = ['motif', 'period', 'ref_len', 'allele_frequency']


def header_text(phenotype_name: str) -> str:
    """associaTR.py:132-137, 209."""
    return ("chrom\tpos\talleles\tn_samples_tested\tlocus_filtered\tp_{0}\tcoeff_{0}\t".format(phenotype_name)
            + 'se_{}\tregression_R^2\t'.format(phenotype_name) + '\t'.join(HEADER_FIELDS) + '\n')
