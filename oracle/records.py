"""
TEST INFRASTRUCTURE ONLY — input container shared by the oracle, the golden
generator and the tests.

A :class:`Locus` is one VCF record reduced to exactly what the hot path reads
from ``cyvcf2.Variant``: CHROM/POS/ID/REF/ALT/FILTER, the
INFO entries, ``genotype.array()`` (int16 ``[S, P+1]``) and the numeric FORMAT
arrays.  :class:`LocusAsVariant` presents a Locus back to the UNMODIFIED
reference as a ``cyvcf2.Variant`` look-alike (same surface as the reference's own
``DummyCyvcf2Record``, trtools/utils/tests/test_trharmonizer.py:18-50) so the very
same arrays can be pushed through the reference, the oracle and the CUDA path.
"""
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

import numpy as np


@dataclass
class Locus:
    vcftype: str
    chrom: str
    pos: int
    ref: str
    alts: List[str]
    info: Dict[str, Any]
    gt: Optional[np.ndarray]            # int16 [S, P+1] or None (no samples)
    fmt: Dict[str, np.ndarray] = field(default_factory=dict)
    record_id: Optional[str] = None
    vcf_filter: Optional[str] = None    # cyvcf2 semantics: None for '.'/PASS


def locus_from_variant(variant, vcftype: str, numeric_fmt=None, string_fmt=()) -> Locus:
    """Snapshot a cyvcf2-like Variant into a Locus (copies the arrays)."""
    fmt = {}
    keys = list(variant.FORMAT) if variant.FORMAT else []
    for key in keys:
        if key == 'GT':
            continue
        if numeric_fmt is not None and key not in numeric_fmt and key not in string_fmt:
            continue
        try:
            arr = variant.format(key)
        except KeyError:
            continue
        if arr is None:
            continue
        if arr.dtype.kind in 'if' or key in string_fmt or numeric_fmt is None:
            fmt[key] = np.array(arr, copy=True)
    gt = None if variant.genotype is None else np.array(variant.genotype.array(), dtype=np.int16)
    return Locus(vcftype=vcftype, chrom=variant.CHROM, pos=int(variant.POS), ref=variant.REF,
                 alts=list(variant.ALT), info=dict(variant.INFO), gt=gt, fmt=fmt,
                 record_id=variant.ID, vcf_filter=variant.FILTER)


class _Genotype:
    def __init__(self, gt):
        self._gt = gt
        self.n_samples = gt.shape[0]

    def array(self):
        return self._gt.copy()


class _InfoView(dict):
    """dict with cyvcf2.INFO iteration semantics: iterating yields (key, value)."""

    def __iter__(self):
        return iter(list(self.items()))


class LocusAsVariant:
    """cyvcf2.Variant look-alike over a Locus, for driving the unmodified reference."""

    def __init__(self, locus: Locus):
        self._locus = locus
        self.CHROM = locus.chrom
        self.POS = locus.pos
        self.ID = locus.record_id
        self.REF = locus.ref
        self.ALT = list(locus.alts)
        self.FILTER = locus.vcf_filter
        self.INFO = _InfoView(locus.info)
        self._fmt = {k: np.array(v, copy=True) for k, v in locus.fmt.items()}
        self.FORMAT = ['GT'] + list(self._fmt.keys())
        self._gt = None if locus.gt is None else np.array(locus.gt, dtype=np.int16)
        self._genotypes = None

    @property
    def genotype(self):
        return None if self._gt is None else _Genotype(self._gt)

    @property
    def ploidy(self):
        return self._gt.shape[1] - 1

    def format(self, key):
        if key not in self._fmt:
            raise KeyError(key)
        return self._fmt[key]

    def set_format(self, key, data):
        data = np.asarray(data)
        if data.dtype.kind == 'S':
            data = np.char.decode(data)
        if key not in self.FORMAT:
            self.FORMAT.append(key)
        self._fmt[key] = data

    @property
    def genotypes(self):
        if self._genotypes is None:
            p = self._gt.shape[1] - 1
            self._genotypes = [[int(a) for a in row[:p] if a != -2] + [bool(row[p])]
                               for row in self._gt]
        return self._genotypes

    @genotypes.setter
    def genotypes(self, gts):
        p = self._gt.shape[1] - 1
        arr = np.full((len(gts), p + 1), -2, dtype=np.int16)
        for i, g in enumerate(gts):
            for j, a in enumerate(g[:-1]):
                arr[i, j] = a
            arr[i, p] = 1 if g[-1] else 0
        self._gt = arr
        self._genotypes = [list(g) for g in gts]

    def __str__(self):
        return "{}:{} {} {}".format(self.CHROM, self.POS, self.REF, ",".join(self.ALT))


# ----------------------------------------------------------------------------
# fixture (de)serialisation: loci <-> one .npz (no pickles)
# ----------------------------------------------------------------------------
def _jsonable(v):
    if isinstance(v, (np.integer,)):
        return int(v)
    if isinstance(v, (np.floating,)):
        return float(v)
    if isinstance(v, tuple):
        return [_jsonable(x) for x in v]
    return v


def save_loci(path: str, loci: List[Locus], extra: Optional[Dict[str, Any]] = None,
              info_keys=None, sample_names=None):
    """Store loci (+ a JSON-able ``extra`` payload, e.g. reference outputs) in one npz."""
    import json
    L = len(loci)
    with_gt = [l for l in loci if l.gt is not None]
    arrays = {}
    if with_gt:
        S = max(l.gt.shape[0] for l in with_gt)
        pmax = max(l.gt.shape[1] - 1 for l in with_gt)
        gt = np.full((L, S, pmax + 1), -2, dtype=np.int16)
        ploidy = np.zeros(L, dtype=np.int32)
        nsamp = np.full(L, -1, dtype=np.int32)          # -1: record without samples (gt None)
        for i, l in enumerate(loci):
            if l.gt is None:
                continue
            p = l.gt.shape[1] - 1
            n = l.gt.shape[0]
            ploidy[i] = p
            nsamp[i] = n
            gt[i, :n, :p] = l.gt[:, :p]
            gt[i, :n, pmax] = l.gt[:, p]
        arrays['gt'] = gt
        arrays['ploidy'] = ploidy
        arrays['nsamp'] = nsamp
        keys = sorted({k for l in loci for k in l.fmt})
        for k in keys:
            first = next(l.fmt[k] for l in loci if k in l.fmt)
            if first.dtype.kind in 'if':
                ncol = max(l.fmt[k].shape[1] if (k in l.fmt and l.fmt[k].ndim > 1) else 1 for l in loci)
                fill = INT32_MISSING_ if first.dtype.kind == 'i' else np.nan
                arr = np.full((L, S, ncol), fill, dtype=first.dtype)
                has = np.zeros(L, dtype=bool)
                for i, l in enumerate(loci):
                    if k in l.fmt:
                        v = l.fmt[k].reshape(l.gt.shape[0], -1)
                        arr[i, :v.shape[0], :v.shape[1]] = v
                        has[i] = True
                arrays['fmt_' + k] = arr
                arrays['has_' + k] = has
    meta = []
    for l in loci:
        info = {k: _jsonable(v) for k, v in l.info.items() if info_keys is None or k in info_keys}
        meta.append(dict(vcftype=l.vcftype, chrom=l.chrom, pos=int(l.pos), ref=l.ref, alts=list(l.alts),
                         info=info, id=l.record_id, filter=l.vcf_filter))
    arrays['meta'] = np.array(json.dumps(meta))
    arrays['extra'] = np.array(json.dumps(extra if extra is not None else {}))
    arrays['samples'] = np.array(json.dumps(list(sample_names) if sample_names is not None else []))
    np.savez_compressed(path, **arrays)


INT32_MISSING_ = -2147483648


def load_loci(path: str):
    """Inverse of :func:`save_loci` -> (loci, extra, sample_names)."""
    import json
    z = np.load(path, allow_pickle=False)
    meta = json.loads(str(z['meta']))
    extra = json.loads(str(z['extra']))
    samples = json.loads(str(z['samples']))
    fmt_keys = [k[4:] for k in z.files if k.startswith('fmt_')]
    loci = []
    for i, m in enumerate(meta):
        gt = None
        fmt = {}
        if 'gt' in z.files and int(z['nsamp'][i]) >= 0:
            p = int(z['ploidy'][i])
            n = int(z['nsamp'][i])
            g = z['gt'][i][:n]
            gt = np.concatenate([g[:, :p], g[:, -1:]], axis=1).astype(np.int16)
            for k in fmt_keys:
                if z['has_' + k][i]:
                    fmt[k] = np.array(z['fmt_' + k][i][:n])
        loci.append(Locus(vcftype=m['vcftype'], chrom=m['chrom'], pos=m['pos'], ref=m['ref'],
                          alts=m['alts'], info=m['info'], gt=gt, fmt=fmt, record_id=m['id'],
                          vcf_filter=m['filter']))
    return loci, extra, samples


def synth_to_loci(sloci, calls, with_fmt=True) -> List[Locus]:
    """Synthetic block (trtools_b200.synth) -> list of Locus (HipSTR records)."""
    out = []
    for i in range(calls.gt.shape[0]):
        fmt = {}
        if with_fmt:
            fmt = {'DP': calls.dp[i][:, None].copy(), 'DSTUTTER': calls.dstutter[i][:, None].copy(),
                   'DFLANKINDEL': calls.dflankindel[i][:, None].copy(), 'Q': calls.q[i][:, None].copy()}
        out.append(Locus(vcftype='hipstr', chrom=sloci.chrom[i], pos=int(sloci.pos[i]), ref=sloci.ref[i],
                         alts=list(sloci.alts[i]),
                         info={'START': int(sloci.start[i]), 'END': int(sloci.end[i]),
                               'PERIOD': int(sloci.period[i])},
                         gt=calls.gt[i].copy(), fmt=fmt, record_id="STR_%d" % (i + sloci.locus_offset)))
    return out
