"""
TEST INFRASTRUCTURE ONLY — numpy/pure-Python restatement of the reference's
harmonizer layer (``trtools/utils/tr_harmonizer.py`` and the helpers it calls in
``trtools/utils/utils.py``).  Every function cites the reference lines it follows.

Parity pinned: tests/test_oracle_golden.py checks these functions against
(a) golden vectors produced by the UNMODIFIED reference on its own fixtures
(tests/golden/*.npz, generator tests/golden/make_golden.py) and (b) the
reference's in-code known answers (test_trharmonizer.py / test_utils.py).
"""
import math
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from .records import Locus

_NUC = {"A": 0, "C": 1, "G": 2, "T": 3}      # utils.py:17


# ----------------------------------------------------------------------------
# string helpers (trtools/utils/utils.py)
# ----------------------------------------------------------------------------
def canonical_one_strand(repseq: str) -> str:
    """utils.py:396-427 — lexicographically smallest rotation under A<C<G<T.
    A non-ACGT base raises KeyError exactly like the reference's dict lookup."""
    repseq = repseq.upper()
    n = len(repseq)
    best = repseq
    for i in range(n):
        rot = repseq[n - i:] + repseq[:n - i]
        for j in range(n):
            a, b = _NUC[rot[j]], _NUC[best[j]]
            if a < b:
                best = rot
            elif a > b:
                break
    return best


def infer_repeat_sequence(seq: str, period: int) -> str:
    """utils.py:465-508.  The reference rescans the same non-overlapping k-mers
    ``period`` times (its ``offset`` is unused), keeping a running best that is
    replaced only when the current leader's count becomes strictly larger; the
    survivor is the k-mer that first reaches the final maximum count."""
    if period > len(seq):
        return "N" * period
    counts: Dict[str, int] = {}
    best_kmer, best_copies = None, 0
    start = 0
    while start + period <= len(seq):
        kmer = seq[start:start + period]
        counts[kmer] = counts.get(kmer, 0) + 1
        start += period
        # leader = first-inserted k-mer holding the maximum count (python max() tie rule)
        leader = max(counts, key=lambda k: counts[k])
        if counts[leader] > best_copies:
            best_kmer, best_copies = leader, counts[leader]
    return canonical_one_strand(best_kmer)


def fabricate_allele(motif: str, length: float) -> str:
    """utils.py:566-602."""
    fab = math.floor(length) * motif
    idx = 0
    while (len(fab) + 1) / len(motif) < length:
        fab += motif[idx]
        idx += 1
    return fab


def homopolymer_run(seq: str) -> int:
    """utils.py:340-360."""
    if len(seq) == 0:
        return 0
    seq = seq.upper()
    best = run = 1
    for a, b in zip(seq, seq[1:]):
        run = run + 1 if a == b else 1
        best = max(best, run)
    return best


# ----------------------------------------------------------------------------
# HarmonizeRecord / TRRecord.__init__ (tr_harmonizer.py:264-550, 693-808)
# ----------------------------------------------------------------------------
@dataclass
class Harmonized:
    ref_allele: str
    alt_alleles: List[str]
    motif: str
    record_id: Optional[str]
    pos: int
    end_pos: int
    ref_allele_length: float
    alt_allele_lengths: List[float]
    full_alleles: Optional[Tuple[str, List[str]]]
    quality_field: Optional[str]
    fabricated_ref: bool
    fabricated_alts: bool
    min_allele_length: float
    max_allele_length: float

    @property
    def allele_lengths(self) -> List[float]:
        return [self.ref_allele_length] + list(self.alt_allele_lengths)

    @property
    def seq_alleles(self) -> List[str]:
        return [self.ref_allele] + list(self.alt_alleles)


def harmonize(locus: Locus) -> Harmonized:
    """Dispatch of tr_harmonizer.py:264-300 followed by TRRecord.__init__ :693-773."""
    vt = locus.vcftype
    info = locus.info
    where = "{}:{}".format(locus.chrom, locus.pos)
    full_alleles = None
    ref_len = None
    alt_lens = None
    harmonized_pos = None
    if vt in ("hipstr", "longtr"):                                   # :336-408
        if info.get("START") is None or info.get("END") is None or info.get("PERIOD") is None:
            raise TypeError("Record at {} is missing one of the mandatory HipSTR/LongTR info fields "
                            "START, END, PERIOD. ".format(where))
        pos = int(locus.pos)
        start_offset = int(info["START"]) - pos
        neg_end_offset = int(info["END"]) - pos + 1 - len(locus.ref)
        if not (start_offset == 0 and neg_end_offset == 0):
            full_alleles = (locus.ref.upper(), [a.upper() for a in locus.alts])
        stop = neg_end_offset if neg_end_offset != 0 else None
        ref_allele = locus.ref[start_offset:stop].upper()
        alt_alleles = [str(a)[start_offset:stop].upper() for a in locus.alts]
        # NB reference quirk: motif inferred on the ALREADY trimmed allele sliced again (:397)
        motif = infer_repeat_sequence(ref_allele[start_offset:], info["PERIOD"])
        record_id = locus.record_id
        quality = 'Q' if info.get('IMP') is None else None
        harmonized_pos = int(info["START"])
    elif vt == "gangstr":                                             # :303-333
        if info.get('RU') is None:
            raise TypeError("Record at {} is missing mandatory GangSTR info field RU. ".format(where))
        if info.get('VID') is not None:
            raise TypeError("Trying to read an AdVNTR record as a GangSTR record {}".format(where))
        if info.get('VARID') is not None:
            raise TypeError("Trying to read an EH record as a GangSTR record {}".format(where))
        ref_allele = locus.ref.upper()
        alt_alleles = [a.upper() for a in locus.alts]
        motif = info["RU"].upper()
        record_id = None
        quality = 'Q' if info.get('IMP') is None else None
    elif vt == "advntr":                                              # :411-436
        if info.get('RU') is None or info.get('VID') is None:
            raise TypeError("Record at {} is missing one of the mandatory ADVNTR info fields RU, VID. ".format(where))
        ref_allele = locus.ref.upper()
        alt_alleles = [a.upper() for a in locus.alts]
        motif = info["RU"].upper()
        record_id = info["VID"]
        quality = 'ML' if info.get('IMP') is None else None
    elif vt == "popstr":                                              # :473-512
        if info.get('Motif') is None:
            raise TypeError("Record at {} is missing mandatory PopSTR info field MOTIF".format(where))
        ref_allele = locus.ref.upper()
        motif = info["Motif"].upper()
        record_id = locus.record_id
        alt_lens = []
        for alt in locus.alts:
            alt = str(alt)
            if alt[0] != "<" or alt[-1] != ">":
                raise TypeError("This record does not look like a PopSTR record.")
            alt_lens.append(float(alt[1:-1]))
        alt_alleles = None
        quality = None
    elif vt == "eh":                                                  # :515-550
        if info.get('VARID') is None or info.get('RU') is None:
            raise TypeError("Record at {} is missing one of the mandatory ExpansionHunter info fields VARID, RU. ".format(where))
        record_id = info["VARID"]
        motif = info["RU"].upper()
        ref_len = int(info["RL"]) / len(motif)
        alt_lens = []
        for alt in locus.alts:
            alt = str(alt)
            if alt[:4] != "<STR" or alt[-1] != ">":
                raise TypeError("This record does not look like an EH record.")
            alt_lens.append(float(alt[4:-1]))
        ref_allele = None
        alt_alleles = None
        quality = None
    else:
        raise ValueError("{} is not an excepted TR vcf type".format(vt))

    # ---- TRRecord.__init__ :693-773 --------------------------------------
    pos = harmonized_pos if harmonized_pos is not None else locus.pos
    if ref_len is not None:
        fabricated_ref = True
        ref_allele = fabricate_allele(motif, ref_len)
    else:
        fabricated_ref = False
        ref_len = len(ref_allele) / len(motif)                        # :740
    end_pos = round(pos + ref_len * len(motif) - 1)                   # :745
    if alt_lens is not None:
        fabricated_alts = True
        alt_alleles = [fabricate_allele(motif, l) for l in alt_lens]
    else:
        fabricated_alts = False
        alt_lens = [len(a) / len(motif) for a in alt_alleles]         # :757-759
    if len(alt_alleles) > 0:
        mn = min(ref_len, min(alt_lens))
        mx = max(ref_len, max(alt_lens))
    else:
        mn = mx = ref_len
    # _CheckRecord :775-808
    if len(alt_alleles) != len(locus.alts):
        raise ValueError("Underlying record does not have the same number of alt alleles")
    if full_alleles:
        fref, falts = full_alleles
        if ref_allele not in fref:
            raise ValueError("could not find ref allele inside full ref allele")
        for i, (fa, a) in enumerate(zip(falts, alt_alleles)):
            if a not in fa:
                raise ValueError("Could not find alt allele {} inside its full alt allele".format(i))
    return Harmonized(ref_allele=ref_allele, alt_alleles=alt_alleles, motif=motif,
                      record_id=record_id, pos=pos, end_pos=end_pos,
                      ref_allele_length=ref_len, alt_allele_lengths=list(alt_lens),
                      full_alleles=full_alleles, quality_field=quality,
                      fabricated_ref=fabricated_ref, fabricated_alts=fabricated_alts,
                      min_allele_length=mn, max_allele_length=mx)


# ----------------------------------------------------------------------------
# TRRecord accessors (tr_harmonizer.py:810-1575)
# ----------------------------------------------------------------------------
def genotype_indices(gt: Optional[np.ndarray]) -> Optional[np.ndarray]:
    """:829-862 — ``genotype.array().astype(int)``."""
    if gt is None:
        return None
    return np.asarray(gt).astype(int)


def called_samples(gt, strict: bool = True):
    """:864-897."""
    idx = genotype_indices(gt)
    if idx is None:
        return None
    if strict:
        return ~np.any(idx[:, :-1] == -1, axis=1)
    return ~np.all(np.logical_or(idx[:, :-1] == -1, idx[:, :-1] == -2), axis=1)


def sample_ploidies(gt):
    """:899-919."""
    idx = genotype_indices(gt)
    if idx is None:
        return None
    return idx.shape[1] - 1 - np.sum(idx[:, :-1] == -2, axis=1)


def call_rate(gt, strict: bool = True):
    """:921-946."""
    c = called_samples(gt, strict)
    if c is None:
        return None
    return np.sum(c) / c.shape[0]


def length_genotypes(h: Harmonized, gt):
    """:1210-1245 — gather through ``[ref_len, *alt_lens, -2, -1]`` so the negative
    sentinels index the two trailing entries; phase column copied back."""
    idx = genotype_indices(gt)
    if idx is None:
        return None
    table = np.array([h.ref_allele_length, *h.alt_allele_lengths, -2, -1])
    out = table[idx]
    out[:, -1] = idx[:, -1]
    return out


def _string_array(idx, seqs: List[str]):
    """:948-961."""
    width = max(len(a) for a in seqs)
    arr = np.empty(idx.shape, dtype="<U{}".format(width))
    arr[:, -1][idx[:, -1] == 0] = '0'
    arr[:, -1][idx[:, -1] == 1] = '1'
    for k, s in enumerate(seqs):
        arr[:, :-1][idx[:, :-1] == k] = s
    arr[:, :-1][idx[:, :-1] == -1] = '.'
    arr[:, :-1][idx[:, :-1] == -2] = ','
    return arr


def string_genotypes(h: Harmonized, gt):
    """:963-1017."""
    idx = genotype_indices(gt)
    if idx is None:
        return None
    return _string_array(idx, h.seq_alleles)


def full_string_genotypes(h: Harmonized, gt):
    """:1019-1047."""
    if h.full_alleles is None:
        return string_genotypes(h, gt)
    idx = genotype_indices(gt)
    if idx is None:
        return None
    return _string_array(idx, [h.full_alleles[0]] + list(h.full_alleles[1]))


def unique_string_genotype_mapping(h: Harmonized) -> Dict[int, int]:
    """:1049-1082."""
    if h.full_alleles is None:
        return {i: i for i in range(len(h.alt_alleles) + 1)}
    first: Dict[str, int] = {}
    out = {}
    for i, a in enumerate(h.seq_alleles):
        out[i] = first.setdefault(a, i)
    return out


def unique_length_genotype_mapping(h: Harmonized) -> Dict[int, int]:
    """:1247-1273 (keyed on bp length of the allele string)."""
    first: Dict[int, int] = {}
    out = {}
    for i, a in enumerate(h.seq_alleles):
        out[i] = first.setdefault(len(a), i)
    return out


def _select_gts(h, gt, uselength, index, fullgenotypes):
    if uselength and fullgenotypes:
        raise ValueError("Can't specify both uselength and fullgenotypes")
    if index and not uselength:
        raise ValueError("Specified uselength=False and index at the same time")
    if index:
        return genotype_indices(gt), -1, -2
    if uselength:
        return length_genotypes(h, gt), -1, -2
    if not fullgenotypes:
        return string_genotypes(h, gt), '.', ','
    return full_string_genotypes(h, gt), '.', ','


def genotype_counts(h: Harmonized, gt, sample_index=None, uselength=True, index=False,
                    fullgenotypes=False, include_nocalls=False) -> Dict[tuple, int]:
    """:1326-1418 — sort haplotypes within a call, row-unique, drop rows holding the
    no-call sentinel (ploidy pads are kept)."""
    gts, nocall, _ = _select_gts(h, gt, uselength, index, fullgenotypes)
    if gts is None:
        return {}
    gts = np.sort(gts[:, :-1], axis=1)
    if sample_index is not None:
        gts = gts[sample_index, :]
    rows, counts = np.unique(gts, axis=0, return_counts=True)
    out = dict(zip(tuple(map(tuple, rows)), counts))
    if not include_nocalls:
        for g in [g for g in out if nocall in g]:
            del out[g]
    return out


def allele_counts(h: Harmonized, gt, sample_index=None, *, uselength=True, index=False,
                  fullgenotypes=False) -> Dict[Any, int]:
    """:1420-1499 — partial calls contribute their called haplotype."""
    gts, nocall, pad = _select_gts(h, gt, uselength, index, fullgenotypes)
    if gts is None:
        return {}
    gts = gts[:, :-1]
    if sample_index is not None:
        gts = gts[sample_index, :]
    gts = gts[gts != nocall]
    gts = gts[gts != pad]
    keys, counts = np.unique(gts, return_counts=True)
    return dict(zip(keys, counts))


def allele_freqs(h: Harmonized, gt, sample_index=None, *, uselength=True, index=False,
                 fullgenotypes=False) -> Dict[Any, float]:
    """:1501-1540."""
    ac = allele_counts(h, gt, sample_index, uselength=uselength, index=index,
                       fullgenotypes=fullgenotypes)
    total = float(sum(ac.values()))
    return {k: v / total for k, v in ac.items()}


def max_allele(h: Harmonized, gt, sample_index=None) -> float:
    """:1542-1575."""
    keys = allele_counts(h, gt, sample_index, uselength=True).keys()
    if len(keys) == 0:
        return np.nan
    return max(keys)


def dosages_bestguess(h: Harmonized, gt, norm: bool = False):
    """:1141-1153, 1191-1205 (bestguess / bestguess_norm branches of GetDosages)."""
    if gt is None or gt.shape[0] == 0:
        return None
    lengts = length_genotypes(h, gt)
    if norm:
        lengts[lengts == -1] = np.nan
        lengts[lengts == -2] = np.nan
    else:
        lengts[lengts == -1] = 0
        lengts[lengts == -2] = 0
    unnorm = lengts[:, :-1].sum(axis=1).astype(np.float32)
    if not norm:
        return unnorm
    if h.min_allele_length == h.max_allele_length:
        return np.zeros(gt.shape[0], dtype=np.float32)
    d = (unnorm - 2 * h.min_allele_length) / (h.max_allele_length - h.min_allele_length)
    if np.any(d >= 2.1) or np.any(d <= -0.1):
        raise ValueError("Error normalizing dosages: value >=2.1 or <=-0.1 detected")
    return np.clip(d, 0, 2)
