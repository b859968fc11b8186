"""
Drop-in mirror of ``trtools.utils.tr_harmonizer`` (reference trtools/utils/tr_harmonizer.py)
backed by the CUDA library: same names, arguments, return dtypes and exception types, but the
allele-string harmonisation and every counting accessor run on the GPU over blocks of records.

* ``TRRecordHarmonizer`` reads ahead ``block_size`` records, stages them as one block
  (``trtools_b200.block``) and yields ``TRRecord`` views of it.
* ``HarmonizeRecord(vcftype, record)`` (one record) is a block of one locus.
* ``TRRecord(vcfrecord, ref_allele, alt_alleles, motif, ...)`` keeps the reference's
  constructor (tr_harmonizer.py:693-773) and also goes through the GPU.

Per-sample array accessors (``GetGenotypeIndicies``, ``GetLengthGenotypes``, ``GetStringGenotypes``)
only materialise arrays at the API edge from the device-computed allele tables; all counts,
frequencies and statistics come from the scan kernels.  There is no CPU fallback: without the
CUDA library / a device these raise.
"""
import enum
import re
import warnings
from typing import Any, Callable, Dict, Iterator, List, Optional, Set, Tuple, Union

import numpy as np

from . import _lib, block as _block
from . import common

_beagle_error = _block._beagle_error


class VcfTypes(enum.Enum):
    """The different tr callers that tr_harmonizer supports (reference :23-38)."""
    gangstr = "gangstr"
    advntr = "advntr"
    hipstr = "hipstr"
    eh = "eh"
    popstr = "popstr"
    longtr = "longtr"

    def __repr__(self):
        return '<{}.{}>'.format(self.__class__.__name__, self.name)


class TRDosageTypes(enum.Enum):
    """Ways to compute TR dosages (reference :40-47)."""
    bestguess = "bestguess"
    beagleap = "beagleap"
    bestguess_norm = "bestguess_norm"
    beagleap_norm = "beagleap_norm"

    def __repr__(self):
        return '<{}.{}>'.format(self.__class__.__name__, self.name)


def _ToVCFType(vcftype: Union[str, VcfTypes]):
    """reference :49-66."""
    if isinstance(vcftype, str):
        if vcftype not in VcfTypes.__members__:
            raise ValueError(("{} is not an excepted TR vcf type. "
                              "Expected one of {}").format(vcftype, list(VcfTypes.__members__)))
        return VcfTypes[vcftype]
    elif isinstance(vcftype, VcfTypes):
        return vcftype
    else:
        raise TypeError("{} (of type {}) is not a vcftype".format(vcftype, type(vcftype)))


def MayHaveImpureRepeats(vcftype: Union[str, VcfTypes]):
    """reference :69-104."""
    return _ToVCFType(vcftype) in (VcfTypes.hipstr, VcfTypes.longtr, VcfTypes.advntr, VcfTypes.popstr)


def HasLengthRefGenotype(vcftype: Union[str, VcfTypes]):
    """reference :107-139."""
    return _ToVCFType(vcftype) == VcfTypes.eh


def HasLengthAltGenotypes(vcftype: Union[str, VcfTypes]):
    """reference :142-172."""
    return _ToVCFType(vcftype) in (VcfTypes.popstr, VcfTypes.eh)


def InferVCFType(vcffile, vcftype: Union[str, VcfTypes] = "auto") -> VcfTypes:
    """Caller detection from the lower-cased raw header (reference :180-244)."""
    possible = set()
    header = vcffile.raw_header.lower()
    if 'command=' in header and 'gangstr' in header:
        possible.add(VcfTypes.gangstr)
    if 'command=' in header and 'hipstr' in header:
        possible.add(VcfTypes.hipstr)
    if 'command=' in header and 'longtr' in header:
        possible.add(VcfTypes.longtr)
    if 'source=advntr' in header:
        possible.add(VcfTypes.advntr)
    if 'source=popstr' in header:
        possible.add(VcfTypes.popstr)
    if re.search(r'ALT=<ID=STR\d+'.lower(), header):
        possible.add(VcfTypes.eh)
    if len(possible) == 0:
        raise TypeError('Could not identify the type of this vcf')
    if vcftype == 'auto':
        if len(possible) == 1:
            return next(iter(possible))
        raise TypeError(('Confused - this vcf looks like it could have '
                         'been any of the types: {}. Please specify '
                         '--vcftype to choose one of them').format(possible))
    user = _ToVCFType(vcftype)
    if user in possible:
        return user
    raise TypeError(('Confused - this vcf looks like it could have '
                     'been any of the types: {}. But you specified: '
                     '--vcftype {} which is not one of those types.'.format(possible, vcftype)))


def IsBeagleVCF(vcffile) -> bool:
    """reference :246-262."""
    return bool(re.search('##source=(\'|")beagle', vcffile.raw_header.lower()))


def HarmonizeRecord(vcftype: Union[str, VcfTypes], vcfrecord, ctx=None):
    """One record -> TRRecord through a block of one locus (reference :264-300)."""
    vt = _ToVCFType(vcftype)
    ctx = ctx or _lib.default_context()
    blk = _block.build_block(ctx, vt.name, [vcfrecord])
    return TRRecord._from_block(blk, 0, vcfrecord)


class _Cyvcf2FormatDict():
    """dict-like façade over ``record.format(key)`` (reference :561-588)."""

    def __init__(self, record):
        self.record = record

    def __getitem__(self, key: str):
        return self.record.format(key)

    def __len__(self):
        return len(self.record.FORMAT)

    def __iter__(self):
        return iter(self.record.FORMAT)

    def __contains__(self, key: str):
        return key in self.record.FORMAT

    def keys(self):
        return self.record.FORMAT

    def get(self, key: str):
        return self.record.format(key)


class _ExplicitRecord:
    """Adapter: the arguments of the reference's TRRecord constructor as a RecordMeta."""


def _sample_mask(sample_index, n_samples: int) -> Optional[np.ndarray]:
    """numpy-style sample_index (bool mask / index list / None) -> uint8 [S] multiplicity-free mask."""
    if sample_index is None:
        return None
    idx = np.asarray(sample_index)
    if idx.dtype == bool:
        if idx.shape[0] != n_samples:
            raise IndexError("boolean index did not match the number of samples")
        return idx.astype(np.uint8)
    mask = np.zeros(n_samples, dtype=np.int64)
    np.add.at(mask, idx.astype(np.int64), 1)
    if mask.max(initial=0) > 1:
        raise ValueError("trtools_b200: sample_index with repeated samples is not supported")
    return mask.astype(np.uint8)


class TRRecord:
    """
    Caller-agnostic view of a TR VCF record (reference class TRRecord :591-1647); attributes and
    methods keep the reference's names, meanings and return dtypes.
    """

    def __init__(self, vcfrecord, ref_allele: Optional[str], alt_alleles: Optional[List[str]], motif: str,
                 record_id: str, quality_field: Optional[str], *, harmonized_pos: Optional[int] = None,
                 full_alleles: Optional[Tuple[str, List[str]]] = None, ref_allele_length: Optional[float] = None,
                 alt_allele_lengths: Optional[List[float]] = None,
                 quality_score_transform: Optional[Callable[..., float]] = None, ctx=None):
        # argument validation exactly as reference :720-731
        if full_alleles is not None and (alt_alleles is None or ref_allele is None):
            raise ValueError("Cannot set full alleles without setting regular alleles")
        if alt_allele_lengths is not None and alt_alleles is not None:
            raise ValueError("Must specify only the sequences or the lengths of the alt alleles, not both.")
        if ref_allele_length is not None and alt_allele_lengths is None:
            raise ValueError("If the ref allele is specified by length, the alt alleles must be too.")
        m = _block.RecordMeta()
        m.vcftype = "gangstr"      # explicit alleles: no flank trimming, motif supplied
        m.chrom = vcfrecord.CHROM
        m.vcf_pos = int(vcfrecord.POS)
        n_alt = len(alt_alleles) if alt_alleles is not None else len(alt_allele_lengths)
        m.ref = ref_allele if ref_allele is not None else ""
        m.alts = list(alt_alleles) if alt_alleles is not None else [""] * n_alt
        m.start = m.vcf_pos
        m.end = m.vcf_pos + len(m.ref) - 1
        m.period = len(motif)
        m.motif_in = motif
        m.record_id = record_id
        m.quality_field = quality_field
        m.harmonized_pos = harmonized_pos
        m.fabricated_ref = ref_allele_length is not None
        m.fabricated_alts = alt_allele_lengths is not None
        m.given_len = None
        if m.fabricated_ref or m.fabricated_alts:
            m.given_len = [ref_allele_length if m.fabricated_ref else np.nan] + \
                ([float(x) for x in alt_allele_lengths] if m.fabricated_alts else [np.nan] * n_alt)
        ctx = ctx or _lib.default_context()
        gt = None
        if vcfrecord.genotype is not None:
            g = np.asarray(vcfrecord.genotype.array())
            gt = g.astype(np.int16).reshape(1, g.shape[0], g.shape[1]) if g.ndim == 2 and g.shape[0] > 0 else \
                np.zeros((1, 0, 3), dtype=np.int16)
        blk = _block.Block(ctx, "gangstr", [m], gt)
        self._init_from_block(blk, 0, vcfrecord, full_alleles=full_alleles,
                              quality_score_transform=quality_score_transform, explicit=True)

    @classmethod
    def _from_block(cls, blk: "_block.Block", l: int, vcfrecord):
        self = cls.__new__(cls)
        self._init_from_block(blk, l, vcfrecord)
        return self

    def _init_from_block(self, blk, l, vcfrecord, full_alleles=None, quality_score_transform=None, explicit=False):
        m = blk.metas[l]
        self._blk = blk
        self._l = l
        self.vcfrecord = vcfrecord
        sl = blk.allele_slice(l)
        self._sl = sl
        alleles = blk.trimmed_alleles(l)
        flags = int(blk.h["flags"][l])
        if flags & _lib.HF_MOTIF_NONACGT:
            raise KeyError(blk.motif(l))       # GetCanonicalOneStrand utils.py:420 (nucToNumber lookup)
        self.motif = m.motif_in if explicit else blk.motif(l)
        self.ref_allele = alleles[0]
        self.alt_alleles = alleles[1:]
        self.record_id = m.record_id
        self.chrom = vcfrecord.CHROM
        self.pos = m.harmonized_pos if m.harmonized_pos is not None else vcfrecord.POS
        self.info = dict(vcfrecord.INFO)
        self.format = _Cyvcf2FormatDict(vcfrecord)
        if explicit:
            self.full_alleles = full_alleles
        elif flags & _lib.HF_HAS_FULL:
            self.full_alleles = (m.ref.upper(), [a.upper() for a in m.alts])
        else:
            self.full_alleles = None
        self.full_alleles_pos = self.vcfrecord.POS
        lens = blk.h["allele_len"][sl]
        self.ref_allele_length = float(lens[0])
        self.alt_allele_lengths = [float(x) for x in lens[1:]]
        self.quality_field = m.quality_field
        self.quality_score_transform = quality_score_transform
        self.has_fabricated_ref_allele = m.fabricated_ref
        self.has_fabricated_alt_alleles = m.fabricated_alts
        self.end_pos = round(self.pos + self.ref_allele_length * len(self.motif) - 1)
        self.full_alleles_end_pos = self.end_pos if self.full_alleles is None else \
            round(self.full_alleles_pos + len(self.full_alleles[0]) - 1)
        if len(self.alt_alleles) > 0:
            self.min_allele_length = min(self.ref_allele_length, min(self.alt_allele_lengths))
            self.max_allele_length = max(self.ref_allele_length, max(self.alt_allele_lengths))
        else:
            self.min_allele_length = self.ref_allele_length
            self.max_allele_length = self.ref_allele_length
        try:
            self._CheckRecord()
        except ValueError as e:
            raise ValueError(("Invalid TRRecord. TRRecord: {} Original record:"
                              " {}").format(str(self), str(self.vcfrecord)), e)

    def _CheckRecord(self):
        """reference :775-808."""
        if len(self.alt_alleles) != len(self.vcfrecord.ALT):
            raise ValueError("Underlying record does not have the same "
                             "number of alt alleles as given to the TRRecord "
                             "constructor. Underlying alt alleles: {}, "
                             " constructor alt alleles: {}".format(self.vcfrecord.ALT, self.alt_alleles))
        if self.full_alleles:
            if len(self.full_alleles) != 2:
                raise ValueError("full_alleles doesn't have both a ref allele and alt alleles")
            full_ref, full_alts = self.full_alleles
            if len(full_alts) != len(self.alt_alleles):
                raise ValueError("Different number of full alternate alleles than normal alt alleles")
            if self.ref_allele not in full_ref:
                raise ValueError("could not find ref allele inside full ref allele")
            for idx, (full_alt, alt) in enumerate(zip(full_alts, self.alt_alleles)):
                if alt not in full_alt:
                    raise ValueError(("Could not find alt allele {} inside its full alt allele").format(idx))

    # ---- simple accessors ----------------------------------------------------------------------
    def GetMaxPloidy(self) -> int:
        return self.vcfrecord.ploidy

    def GetNumSamples(self) -> int:
        return self.vcfrecord.genotype.n_samples

    def GetGenotypeIndicies(self) -> Optional[np.ndarray]:
        """reference :829-862."""
        if self.vcfrecord.genotype is None:
            return None
        return self.vcfrecord.genotype.array().astype(int)

    def _gpu_counts(self, sample_index=None):
        """(ac by index, n_called, n_called_nonstrict) of this locus from the scan kernels."""
        blk = self._blk
        if sample_index is None:
            st = blk.stats(True)
            g = 0
        else:
            mask = _sample_mask(sample_index, blk.S)
            st = blk.stats(True, group_masks=np.broadcast_to(mask, (1, blk.S)))
            g = 0
        return st, g

    def GetCalledSamples(self, strict: bool = True) -> Optional[np.ndarray]:
        """reference :864-897 (per-sample bool array materialised at the API edge)."""
        gt_idxs = self.GetGenotypeIndicies()
        if gt_idxs is None:
            return None
        if strict:
            return ~np.any(gt_idxs[:, :-1] == -1, axis=1)
        return ~np.all(np.logical_or(gt_idxs[:, :-1] == -1, gt_idxs[:, :-1] == -2), axis=1)

    def GetSamplePloidies(self) -> Optional[np.ndarray]:
        """reference :899-919."""
        gt_idxs = self.GetGenotypeIndicies()
        if gt_idxs is None:
            return None
        return gt_idxs.shape[1] - 1 - np.sum(gt_idxs[:, :-1] == -2, axis=1)

    def GetCallRate(self, strict: bool = True) -> float:
        """reference :921-946 — from the device counters."""
        if self.vcfrecord.genotype is None:
            return None
        st, g = self._gpu_counts()
        n = st["n_called"][g, self._l] if strict else st["n_called_nonstrict"][g, self._l]
        return n / self._blk.S

    def _GetStringGenotypeArray(self, idx_gts: np.ndarray, seq_alleles: List[str]):
        """reference :948-961."""
        max_len = max(len(allele) for allele in seq_alleles)
        seq_array = np.empty(idx_gts.shape, dtype="<U{}".format(max_len))
        seq_array[:, -1][idx_gts[:, -1] == 0] = '0'
        seq_array[:, -1][idx_gts[:, -1] == 1] = '1'
        for allele_idx, seq_allele in enumerate(seq_alleles):
            seq_array[:, :-1][idx_gts[:, :-1] == allele_idx] = seq_allele
        seq_array[:, :-1][idx_gts[:, :-1] == -1] = '.'
        seq_array[:, :-1][idx_gts[:, :-1] == -2] = ','
        return seq_array

    def GetStringGenotypes(self) -> Optional[np.ndarray]:
        """reference :963-1017."""
        idx_gts = self.GetGenotypeIndicies()
        if idx_gts is None:
            return None
        if self.HasFabricatedAltAlleles():
            warnings.warn("String genotypes have been requested for a"
                          " TRRecord generated by a caller which only "
                          "generates length genotypes, not string genotypes"
                          ". Returning a fabricated string genotype. Consider"
                          " requesting length based genotypes instead.")
        return self._GetStringGenotypeArray(idx_gts, [self.ref_allele] + list(self.alt_alleles))

    def GetFullStringGenotypes(self) -> Optional[np.ndarray]:
        """reference :1019-1047."""
        if not self.HasFullStringGenotypes():
            return self.GetStringGenotypes()
        idx_gts = self.GetGenotypeIndicies()
        if idx_gts is None:
            return None
        return self._GetStringGenotypeArray(idx_gts, [self.full_alleles[0]] + list(self.full_alleles[1]))

    def UniqueStringGenotypeMapping(self) -> Dict[int, int]:
        """reference :1049-1082 — from the device sequence classes."""
        n = len(self.alt_alleles) + 1
        if not self.HasFullStringGenotypes():
            return {i: i for i in range(n)}
        cls = self._blk.h["seq_class"][self._sl]
        return {i: int(cls[i]) for i in range(n)}

    def UniqueStringGenotypes(self) -> Set[int]:
        return set(self.UniqueStringGenotypeMapping().values())

    def UniqueLengthGenotypeMapping(self) -> Dict[int, int]:
        """reference :1247-1273 — from the device length classes."""
        cls = self._blk.h["len_class"][self._sl]
        return {i: int(cls[i]) for i in range(len(self.alt_alleles) + 1)}

    def UniqueLengthGenotypes(self) -> Set[int]:
        return set(self.UniqueLengthGenotypeMapping().values())

    def GetLengthGenotypes(self) -> Optional[np.ndarray]:
        """reference :1210-1245: float64 [S, P+1] materialised from the device length table."""
        idx_gts = self.GetGenotypeIndicies()
        if idx_gts is None:
            return None
        allele_lens = np.array([self.ref_allele_length, *self.alt_allele_lengths, -2, -1])
        len_gts = allele_lens[idx_gts]
        len_gts[:, -1] = idx_gts[:, -1]
        return len_gts

    def GetDosages(self, dosagetype: TRDosageTypes = TRDosageTypes.bestguess, strict: bool = True):
        """reference :1098-1208.  One ``trt_dosages`` pass per (block, dosage type) yields the float32 [L][S] tensor;
        this record's row is a view of it.  Record-level validation failures come back as a code per locus and are
        raised (or, when not strict, reported and answered with NaN) here, with the reference's messages."""
        if self.GetNumSamples() == 0:
            return None
        if not isinstance(dosagetype, TRDosageTypes):
            raise ValueError("Unsupported dosagetype")
        n = self.GetNumSamples()

        def fail(msg, raised=None):
            if strict:
                raise ValueError(raised or msg)
            common.WARNING(msg)
            return np.array([np.nan] * n)

        values, codes = self._blk.dosages(dosagetype.value)
        code = int(codes[self._l])
        if code == _lib.DE_NO_AP:
            return fail("Requested Beagle dosages for record at {}:{} but AP1/AP2 fields not found.".format(self.chrom, self.pos))
        if code == _lib.DE_AP_SUM:
            return fail("{}:{} AP1 or AP2 field summing to more than 1 detected".format(self.chrom, self.pos))
        if code == _lib.DE_AP_NEGATIVE:
            return fail("{}:{} Negative AP1 or AP2 fields detected".format(self.chrom, self.pos), "Negative AP1 or AP2 fields detected")
        if code == _lib.DE_NORM_RANGE:
            return fail("{}:{} Error normalizing dosages: value >=2.1 or <=-0.1 detected".format(self.chrom, self.pos))
        return values[self._l].copy()

    def HasFullStringGenotypes(self) -> bool:
        return self.full_alleles is not None

    def HasFabricatedRefAllele(self) -> bool:
        return self.has_fabricated_ref_allele

    def HasFabricatedAltAlleles(self) -> bool:
        return self.has_fabricated_alt_alleles

    # ---- keys of the three allele representations ----------------------------------------------------
    def _allele_keys(self, uselength, index, fullgenotypes):
        if uselength and fullgenotypes:
            raise ValueError("Can't specify both uselength and fullgenotypes")
        if index and not uselength:
            raise ValueError("Specified uselength=False and index at the same"
                             " time, these are mutually exclusive options")
        n = len(self.alt_alleles) + 1
        if index:
            return list(range(n)), -1, -2
        if uselength:
            return [np.float64(self.ref_allele_length)] + [np.float64(x) for x in self.alt_allele_lengths], -1.0, -2.0
        if not fullgenotypes or not self.HasFullStringGenotypes():
            if not self.HasFullStringGenotypes() and self.HasFabricatedAltAlleles():
                warnings.warn("String genotypes have been requested for a TRRecord generated by a caller which "
                              "only generates length genotypes")
            return [np.str_(self.ref_allele)] + [np.str_(a) for a in self.alt_alleles], '.', ','
        return [np.str_(self.full_alleles[0])] + [np.str_(a) for a in self.full_alleles[1]], '.', ','

    def GetGenotypeCounts(self, sample_index: Optional[Any] = None, uselength: bool = True, index: bool = False,
                          fullgenotypes: bool = False, include_nocalls: bool = False) -> Dict[tuple, int]:
        """reference :1326-1418 — the index-genotype table is counted on the GPU
        (``trt_genotype_counts``); keys are mapped to the requested representation here."""
        keys, nocall, pad = self._allele_keys(uselength, index, fullgenotypes)
        if self.vcfrecord.genotype is None:
            return {}
        blk = self._blk
        blk._activate()
        A = len(keys)
        table = blk.ctx.genotype_counts(self._l, A, _sample_mask(sample_index, blk.S))
        allkeys = [pad, nocall] + list(keys)
        out: Dict[tuple, int] = {}
        for cell in np.argwhere(table > 0):
            if not include_nocalls and 1 in cell:
                continue
            gt = tuple(sorted(allkeys[d] for d in cell)) if not index else tuple(int(d) - 2 for d in cell)
            if index:
                gt = tuple(sorted(gt))
            out[gt] = out.get(gt, 0) + int(table[tuple(cell)])
        # np.unique returns rows in sorted order; mirror that ordering of the dict
        items = sorted(out.items(), key=lambda kv: kv[0])
        if index:
            return {tuple(np.int64(x) for x in k): np.int64(v) for k, v in items}
        return {k: np.int64(v) for k, v in items}

    def GetAlleleCounts(self, sample_index: Optional[Any] = None, *, uselength: bool = True, index: bool = False,
                        fullgenotypes: bool = False) -> Dict[Any, int]:
        """reference :1420-1499 — index counts from the scan kernel, folded by representation."""
        keys, _, _ = self._allele_keys(uselength, index, fullgenotypes)
        if self.vcfrecord.genotype is None:
            return {}
        st, g = self._gpu_counts(sample_index)
        ac = st["ac"][g, self._sl]
        out: Dict[Any, int] = {}
        for k, c in zip(keys, ac):
            if c > 0:
                out[k] = out.get(k, 0) + int(c)
        return {k: np.int64(out[k]) for k in sorted(out)}

    def GetAlleleFreqs(self, sample_index: Optional[Any] = None, *, uselength: bool = True, index: bool = False,
                       fullgenotypes: bool = False) -> Dict[Any, float]:
        """reference :1501-1540."""
        allele_counts = self.GetAlleleCounts(uselength=uselength, index=index, fullgenotypes=fullgenotypes,
                                             sample_index=sample_index)
        total = float(sum(allele_counts.values()))
        return {key: value / total for key, value in allele_counts.items()}

    def GetMaxAllele(self, sample_index: Optional[Any] = None) -> float:
        """reference :1542-1575."""
        alleles = self.GetAlleleCounts(uselength=True, sample_index=sample_index).keys()
        if len(alleles) == 0:
            return np.nan
        return max(alleles)

    def HasQualityScores(self) -> bool:
        return (self.quality_field is not None and self.quality_field in self.format)

    def GetQualityScores(self) -> np.ndarray:
        """reference :1592-1615."""
        if not self.HasQualityScores():
            raise TypeError("This TRRecord does not have a corresponding quality score field")
        quality_val = self.format[self.quality_field]
        if self.quality_score_transform is None:
            return quality_val
        return np.apply_along_axis(self.quality_score_transform, 0, quality_val)

    def __str__(self):
        """reference :1617-1647."""
        record_id = self.record_id
        if record_id is None:
            record_id = "{}:{}".format(self.vcfrecord.CHROM, self.vcfrecord.POS)
        if self.HasFullStringGenotypes():
            return "{} {} {} ".format(record_id, self.motif, self.full_alleles[0]) + ",".join(self.full_alleles[1])
        if self.HasFabricatedRefAllele():
            string = "{} {} n_reps:{} ".format(record_id, self.motif, self.ref_allele_length)
        else:
            string = "{} {} {} ".format(record_id, self.motif, self.ref_allele)
        if len(self.alt_alleles) == 0:
            string += '.'
        elif self.HasFabricatedAltAlleles():
            string += ",".join("n_reps:" + str(length) for length in self.alt_allele_lengths)
        else:
            string += ','.join(self.alt_alleles)
        return string


class TRRecordHarmonizer:
    """
    Iterator of TRRecords over a cyvcf2.VCF (reference :1650-1779), reading ahead ``block_size``
    records per GPU block.
    """

    def __init__(self, vcffile, vcftype: Union[str, VcfTypes] = "auto", block_size: int = 512, ctx=None,
                 fmt_keys=()):
        self.vcffile = vcffile
        self.vcftype = InferVCFType(vcffile, vcftype)
        self._record_idx = None
        self._block_size = max(1, int(block_size))
        self._ctx = ctx
        self._fmt_keys = tuple(fmt_keys)
        self._lookahead = None
        self._queue: List[Any] = []
        self._exhausted = False
        self._deferred_error = None
        # the C++ block reader (vcf_ingest.NativeVCF) parses these keys together with GT, and reads
        # runs of the same length as the GPU blocks so that a block is one run
        if hasattr(vcffile, "_prefetch"):
            vcffile._prefetch = tuple(dict.fromkeys(tuple(vcffile._prefetch) + self._fmt_keys))
            vcffile._native_block_loci = self._block_size

    def MayHaveImpureRepeats(self) -> bool:
        return MayHaveImpureRepeats(self.vcftype)

    def HasLengthRefGenotype(self) -> bool:
        return HasLengthRefGenotype(self.vcftype)

    def HasLengthAltGenotypes(self) -> bool:
        return HasLengthAltGenotypes(self.vcftype)

    def HasQualityScore(self) -> bool:
        """reference :1721-1749."""
        if self.vcftype == VcfTypes.gangstr:
            return 'FORMAT=<ID=Q,' in self.vcffile.raw_header
        if self.vcftype in (VcfTypes.hipstr, VcfTypes.longtr, VcfTypes.advntr):
            return not self.IsBeagleVCF()
        return False

    def IsBeagleVCF(self) -> bool:
        return IsBeagleVCF(self.vcffile)

    def __iter__(self) -> Iterator[TRRecord]:
        return self

    @staticmethod
    def _ploidy(rec):
        return rec.ploidy if rec.genotype is not None else 0

    def _read_block(self):
        """Pull up to block_size records of equal ploidy; a parse error is deferred until the
        records before it have been served (same order of events as the reference :1761-1779)."""
        recs = []
        if self._lookahead is not None:
            recs.append(self._lookahead)
            self._lookahead = None
        while len(recs) < self._block_size and not self._exhausted:
            if self._record_idx is None:
                self._record_idx = 1
            self._record_idx += 1
            try:
                rec = next(self.vcffile)
            except StopIteration:
                self._exhausted = True
                break
            except Exception:
                self._deferred_error = ValueError(
                    "Unable to parse the " + str(self._record_idx) + "th tandem "
                    "repeat in the provided VCF. Check that it is properly formatted.")
                self._exhausted = True
                break
            if recs and self._ploidy(rec) != self._ploidy(recs[0]):
                self._lookahead = rec          # starts the next block
                break
            recs.append(rec)
        return recs

    def __next__(self) -> TRRecord:
        if not self._queue:
            recs = self._read_block()
            if not recs:
                if self._deferred_error is not None:
                    err, self._deferred_error = self._deferred_error, None
                    raise err
                raise StopIteration
            ctx = self._ctx or _lib.default_context()
            # validate record by record so that a bad record raises when it is reached, not earlier
            good = []
            for r in recs:
                try:
                    _block.record_meta(self.vcftype.name, r)
                    good.append(r)
                except (TypeError, ValueError) as e:
                    self._deferred_error = e
                    self._exhausted = True
                    self._lookahead = None
                    break
            if not good:
                err, self._deferred_error = self._deferred_error, None
                raise err
            blk = _block.build_block(ctx, self.vcftype.name, good, self._fmt_keys)
            self._queue = [(blk, i, r) for i, r in enumerate(good)]
        blk, i, r = self._queue.pop(0)
        return TRRecord._from_block(blk, i, r)
