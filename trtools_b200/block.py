"""
Blocked ingest: turn a run of cyvcf2 records into the stacked arrays the CUDA kernels consume.

The reference pulls ``genotype.array()`` / ``format(key)`` per record
(trtools/utils/tr_harmonizer.py:561-588, 829-862) and does the allele-string work of
``HarmonizeRecord`` (:264-550) in Python per record.  Here a block of L records is staged once
(GT int16 [L][S][P+1], numeric FORMAT [L][S], one concatenated allele table), copied to HBM and
harmonized by one kernel launch; the per-record Python objects become thin views of the block.

Only the caller-specific INFO validation (mandatory fields, ALT syntax) stays in Python, raising
the same exception types and messages as the reference.
"""
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from . import _lib

_beagle_error = "If this file was imputed by Beagle, did you remember to copy the info fields over?"

# numeric FORMAT fields with a fixed slot in the C-ABI
FIXED_FMT = {"DP": _lib.FMT_DP, "DSTUTTER": _lib.FMT_DSTUTTER, "DFLANKINDEL": _lib.FMT_DFLANKINDEL,
             "Q": _lib.FMT_Q, "QEXP": _lib.FMT_QEXP}
# any other numeric FORMAT field a call filter reads gets one of the AUX slots, in order of appearance


class RecordMeta:
    """Caller-specific scalars of one record, validated like the reference's _Harmonize*Record."""
    __slots__ = ("chrom", "vcf_pos", "record_id", "ref", "alts", "start", "end", "period", "given_len",
                 "motif_in", "quality_field", "harmonized_pos", "vcftype", "fabricated_ref", "fabricated_alts")


def record_meta(vcftype: str, rec) -> RecordMeta:
    """INFO validation + scalars (tr_harmonizer.py:303-550).  Raises TypeError exactly where the
    reference does."""
    m = RecordMeta()
    info = rec.INFO
    m.vcftype = vcftype
    m.chrom = rec.CHROM
    m.vcf_pos = int(rec.POS)
    m.ref = rec.REF
    m.alts = list(rec.ALT) if rec.ALT else []
    m.given_len = None
    m.motif_in = None
    m.harmonized_pos = None
    m.fabricated_ref = False
    m.fabricated_alts = False
    where = "{}:{}".format(rec.CHROM, rec.POS)
    if vcftype in ("hipstr", "longtr"):
        if info.get('START') is None or info.get('END') is None or info.get('PERIOD') is None:
            raise TypeError("Record at {} is missing one of the mandatory HipSTR/LongTR info fields "
                            "START, END, PERIOD. ".format(where) + _beagle_error)
        m.start = int(info['START'])
        m.end = int(info['END'])
        m.period = int(info['PERIOD'])
        m.record_id = rec.ID
        m.quality_field = 'Q' if info.get('IMP') is None else None
        m.harmonized_pos = m.start
    elif vcftype == "gangstr":
        if info.get('RU') is None:
            raise TypeError("Record at {} is missing mandatory GangSTR info field RU. ".format(where) + _beagle_error)
        if info.get('VID') is not None:
            raise TypeError("Trying to read an AdVNTR record as a GangSTR record {}".format(where))
        if info.get('VARID') is not None:
            raise TypeError("Trying to read an EH record as a GangSTR record {}".format(where))
        m.motif_in = info["RU"].upper()
        m.record_id = None
        m.quality_field = 'Q' if info.get('IMP') is None else None
    elif vcftype == "advntr":
        if info.get('RU') is None or info.get('VID') is None:
            raise TypeError("Record at {} is missing one of the mandatory ADVNTR info fields RU, VID. ".format(where)
                            + _beagle_error)
        m.motif_in = info["RU"].upper()
        m.record_id = info["VID"]
        m.quality_field = 'ML' if info.get('IMP') is None else None
    elif vcftype == "popstr":
        if info.get('Motif') is None:
            raise TypeError("Record at {} is missing mandatory PopSTR info field MOTIF".format(where))
        m.motif_in = info["Motif"].upper()
        m.record_id = rec.ID
        m.quality_field = None
        lens = [np.nan]
        for alt in m.alts:
            alt = str(alt)
            if alt[0] != "<" or alt[-1] != ">":
                raise TypeError("This record does not look like a PopSTR record. Alt alleles were not formatted"
                                " as expected")
            lens.append(float(alt[1:-1]))
        m.given_len = lens
        m.fabricated_alts = True
    elif vcftype == "eh":
        if info.get('VARID') is None or info.get('RU') is None:
            raise TypeError("Record at {} is missing one of the mandatory ExpansionHunter info fields VARID, RU. "
                            .format(where) + _beagle_error)
        m.record_id = info["VARID"]
        m.motif_in = info["RU"].upper()
        lens = [int(info["RL"]) / len(m.motif_in)]
        for alt in m.alts:
            alt = str(alt)
            if alt[:4] != "<STR" or alt[-1] != ">":
                raise TypeError("This record does not look like an EH  record. Alt alleles were not formatted"
                                " as expected")
            lens.append(float(alt[4:-1]))
        m.given_len = lens
        m.quality_field = None
        m.fabricated_ref = True
        m.fabricated_alts = True
    else:
        raise ValueError("{} is not an excepted TR vcf type".format(vcftype))
    if vcftype not in ("hipstr", "longtr"):
        m.start = m.vcf_pos
        m.end = m.vcf_pos + len(m.ref) - 1
        m.period = len(m.motif_in)
    return m


def pack_gt(gt: np.ndarray):
    """cyvcf2-layout diploid GT int16 [..., S, 3] -> (uint8 [..., S, 2], phase bits uint8 [..., ceil(S/8)]) or None
    when an allele index does not fit the packed transfer form (see include/trtools_b200.h trt_block_set_gt_packed)."""
    a = gt[..., :2]
    if gt.shape[-1] != 3 or a.size and (int(a.max()) > 252 or int(a.min()) < -2):
        return None
    g2 = np.where(a >= 0, a, 256 + a).astype(np.uint8)
    ph = np.packbits(gt[..., 2] != 0, axis=-1, bitorder="little")
    return g2, ph


def pack_gt4(gt: np.ndarray):
    """cyvcf2-layout diploid GT int16 [..., S, 3] -> (uint8 [..., S] nibble pairs, phase bits) or None when an allele
    index is above 13 (see include/trtools_b200.h trt_block_set_gt_nibble)."""
    a = gt[..., :2]
    if gt.shape[-1] != 3 or a.size and (int(a.max()) > 13 or int(a.min()) < -2):
        return None
    g4 = ((a[..., 0] & 15) | ((a[..., 1] & 15) << 4)).astype(np.uint8)
    ph = np.packbits(gt[..., 2] != 0, axis=-1, bitorder="little")
    return g4, ph


def unpack_gt4(g4: np.ndarray, phase_bits: Optional[np.ndarray] = None) -> np.ndarray:
    """Inverse of :func:`pack_gt4` -> int16 [..., S, 3]."""
    S = g4.shape[-1]
    out = np.empty(g4.shape + (3,), np.int16)
    lo, hi = (g4 & 15).astype(np.int16), (g4 >> 4).astype(np.int16)
    out[..., 0] = np.where(lo >= 14, lo - 16, lo)
    out[..., 1] = np.where(hi >= 14, hi - 16, hi)
    if phase_bits is None:
        out[..., 2] = 0
    else:
        out[..., 2] = np.unpackbits(phase_bits, axis=-1, count=S, bitorder="little")
    return out


def unpack_any(g: np.ndarray, phase_bits: Optional[np.ndarray] = None) -> np.ndarray:
    """Either transfer form of a block ([L][S][2] two-byte or [L][S] nibble) -> int16 [L][S][3]."""
    return unpack_gt(g, phase_bits) if g.ndim == 3 else unpack_gt4(g, phase_bits)


def unpack_gt(gt2: np.ndarray, phase_bits: Optional[np.ndarray] = None) -> np.ndarray:
    """Inverse of :func:`pack_gt` -> int16 [..., S, 3]."""
    S = gt2.shape[-2]
    out = np.empty(gt2.shape[:-1] + (3,), np.int16)
    v = gt2.astype(np.int16)
    out[..., :2] = np.where(v >= 254, v - 256, v)
    if phase_bits is None:
        out[..., 2] = 0
    else:
        out[..., 2] = np.unpackbits(phase_bits, axis=-1, count=S, bitorder="little")
    return out


class Block:
    """L harmonized loci resident on the GPU + the host-side tables describing them."""

    @property
    def gt(self) -> np.ndarray:
        """cyvcf2-layout int16 [L][S][P+1] on the host (expanded on first use when the block was built packed)."""
        if self._gt is None:
            self._gt = unpack_any(*self.gt_packed)
        return self._gt

    def __init__(self, ctx: "_lib.Context", vcftype: str, metas: List[RecordMeta], gt: Optional[np.ndarray],
                 fmt: Optional[Dict[str, np.ndarray]] = None, gt_packed=None):
        """``gt``: cyvcf2-layout int16 [L][S][P+1]; or ``gt_packed`` = (uint8 [L][S][2] or nibble pairs uint8 [L][S],
        phase bits or None), the transfer forms the block reader parses straight from the text (a third / a sixth of
        the host->device bytes)."""
        self.ctx = ctx
        self.vcftype = vcftype
        self.metas = metas
        L = len(metas)
        self.L = L
        self.gt_packed = None
        if gt_packed is not None:
            g2, ph = gt_packed
            self.gt_packed = (np.ascontiguousarray(g2, dtype=np.uint8), None if ph is None else np.ascontiguousarray(ph, dtype=np.uint8))
            self.has_samples = True
            self._gt = None
            self.S = self.gt_packed[0].shape[1]
            self.P = 2
        else:
            self.has_samples = gt is not None
            if gt is None:
                gt = np.zeros((L, 0, 3), dtype=np.int16)
            self._gt = np.ascontiguousarray(gt, dtype=np.int16)
            self.S = self._gt.shape[1]
            self.P = self._gt.shape[2] - 1
        self.rec_ploidy = None       # int32 [L]: GT columns of each record's own array (None: all as wide as the block)
        # ---- allele table ----------------------------------------------------------------------
        seq_parts: List[bytes] = []
        allele_off = [0]
        locus_off = [0]
        given: List[float] = []
        any_given = False
        motifs: List[bytes] = []
        any_motif = False
        for m in metas:
            alleles = [m.ref] + m.alts
            for j, a in enumerate(alleles):
                if m.given_len is not None and not np.isnan(m.given_len[j]):
                    b = b""
                    given.append(m.given_len[j])
                    any_given = True
                else:
                    b = str(a).encode("ascii", "replace")
                    given.append(np.nan)
                seq_parts.append(b)
                allele_off.append(allele_off[-1] + len(b))
            locus_off.append(locus_off[-1] + len(alleles))
            if m.motif_in is not None:
                motifs.append(m.motif_in.encode("ascii", "replace"))
                any_motif = True
            else:
                motifs.append(b"N" * max(m.period, 0))
        self.seqs = b"".join(seq_parts)
        self.allele_off = np.array(allele_off, dtype=np.int64)
        self.locus_off = np.array(locus_off, dtype=np.int32)
        self.pos = np.array([m.vcf_pos for m in metas], dtype=np.int32)
        self.start = np.array([m.start for m in metas], dtype=np.int32)
        self.end = np.array([m.end for m in metas], dtype=np.int32)
        self.period = np.array([m.period for m in metas], dtype=np.int32)
        self.given_len = np.array(given, dtype=np.float64) if any_given else None
        self.motifs_in = b"".join(motifs) if any_motif else None
        # ---- upload + harmonize --------------------------------------------------------------
        ctx.block_begin(L, self.S, self.P, vcftype)
        self._upload_gt()
        ctx.block_set_alleles(self.seqs, self.allele_off, self.locus_off, self.pos, self.start, self.end,
                              self.period, self.given_len, self.motifs_in)
        self.fmt = fmt or {}
        self.fmt_slot: Dict[str, int] = {}
        self._upload_fmt()
        self.h = ctx.harmonize()
        bad = np.nonzero(self.h["flags"] & _lib.HF_BAD_PERIOD)[0]
        if len(bad):
            raise ValueError("Record {}:{} has a non-positive period".format(metas[bad[0]].chrom, metas[bad[0]].vcf_pos))
        self._stats = {}
        self._dosages = {}
        self._records = None          # set by build_block: the AP1 / AP2 arrays are pulled lazily (Beagle dosages only)
        self._ap = None
        ctx._current_block = self

    # ---- cached block-level statistics (one GPU pass per (use_length)) --------------------------
    def stats(self, use_length: bool, nalleles_thresh: float = 0.01, group_masks: Optional[np.ndarray] = None):
        if group_masks is not None:
            return self.ctx_stats(use_length, nalleles_thresh, group_masks)
        key = (bool(use_length), float(nalleles_thresh))
        if key not in self._stats:
            self._stats[key] = self.ctx_stats(use_length, nalleles_thresh, None)
        return self._stats[key]

    def ctx_stats(self, use_length, nalleles_thresh, group_masks):
        self._activate()
        return self.ctx.locus_stats(use_length, group_masks, nalleles_thresh)

    def _upload_gt(self):
        if self.gt_packed is not None and self.gt_packed[0].ndim == 2:
            self.ctx.block_set_gt_nibble(*self.gt_packed)
        elif self.gt_packed is not None:
            self.ctx.block_set_gt_packed(*self.gt_packed)
        else:
            self.ctx.block_set_gt(self._gt)

    def _upload_fmt(self):
        """numeric FORMAT arrays -> device slots (fixed slots for the HipSTR/GangSTR fields, AUX otherwise)"""
        self.fmt_slot = {}
        aux = 0
        for key, arr in self.fmt.items():
            if arr is None or arr.dtype.kind not in "if":
                continue
            if key in FIXED_FMT:
                slot = FIXED_FMT[key]
            else:
                if aux >= _lib.FMT_NAUX:
                    raise ValueError("too many auxiliary FORMAT fields in one block")
                slot = _lib.FMT_AUX0 + aux
                aux += 1
            self.ctx.block_set_format(slot, arr)
            self.fmt_slot[key] = slot

    def add_host_filter_values(self, name: str, values: np.ndarray) -> int:
        """Upload a host-evaluated call filter's output (float [L][S], NaN = keep) as an AUX field."""
        self._activate()
        used = [s for s in self.fmt_slot.values() if s >= _lib.FMT_AUX0]
        slot = (max(used) + 1) if used else _lib.FMT_AUX0
        if slot >= _lib.FMT_AUX0 + _lib.FMT_NAUX:
            raise ValueError("too many auxiliary FORMAT fields in one block")
        arr = np.ascontiguousarray(values, dtype=np.float32)
        self.fmt[name] = arr
        self.ctx.block_set_format(slot, arr)
        self.fmt_slot[name] = slot
        return slot

    def _activate(self):
        """Make this block the context's current block again (another block may have replaced it)."""
        if getattr(self.ctx, "_current_block", None) is not self:
            self.ctx.block_begin(self.L, self.S, self.P, self.vcftype)
            self._upload_gt()
            self.ctx.block_set_alleles(self.seqs, self.allele_off, self.locus_off, self.pos, self.start, self.end,
                                       self.period, self.given_len, self.motifs_in)
            self._upload_fmt()
            self.ctx.harmonize()
            if self._ap is not None:
                self.ctx.block_set_ap(*self._ap)
        self.ctx._current_block = self

    # ---- Beagle allele probabilities / dosages -------------------------------------------------
    def ensure_ap(self):
        """Pull FORMAT AP1 / AP2 (float32 [S][A-1] per record, as cyvcf2 returns them) of every record, stack them and
        upload them.  Records without the fields get zero rows and has_ap = 0 (GetDosages reports them)."""
        self._activate()
        if self._ap is not None:
            return self._ap
        if self._records is None:
            raise ValueError("this block was not built from records: AP1/AP2 are not available")
        parts1, parts2 = [], []
        has = np.ones(self.L, np.uint8)
        for l, rec in enumerate(self._records):
            nalt = int(self.locus_off[l + 1] - self.locus_off[l]) - 1
            ap = []
            for key in ("AP1", "AP2"):
                arr = None
                try:
                    if key in (rec.FORMAT or []):
                        arr = rec.format(key)
                except KeyError:
                    arr = None
                if arr is None:
                    has[l] = 0
                    ap.append(np.zeros((self.S, nalt), np.float32))
                    continue
                arr = np.asarray(arr, dtype=np.float32).reshape(self.S, -1)
                if arr.shape[1] != nalt:
                    fixed = np.zeros((self.S, nalt), np.float32)
                    k = min(nalt, arr.shape[1])
                    fixed[:, :k] = arr[:, :k]
                    arr = fixed
                ap.append(arr)
            parts1.append(ap[0].reshape(-1))
            parts2.append(ap[1].reshape(-1))
        cat = lambda p: np.concatenate(p) if p else np.zeros(0, np.float32)
        self._ap = (cat(parts1), cat(parts2), has)
        self.ctx.block_set_ap(*self._ap)
        return self._ap

    def dosages(self, kind: str):
        """(float32 [L, S], int32 [L] validation codes) of TRRecord.GetDosages(kind) for the whole block, cached."""
        if kind not in self._dosages:
            if kind.startswith("beagleap"):
                self.ensure_ap()
            else:
                self._activate()
            self._dosages[kind] = self.ctx.dosages(kind)
        return self._dosages[kind]

    # ---- per-locus views ---------------------------------------------------------------------------
    def allele_slice(self, l: int) -> slice:
        return slice(int(self.locus_off[l]), int(self.locus_off[l + 1]))

    def trimmed_alleles(self, l: int) -> List[str]:
        """Upper-cased trimmed allele strings of locus l (fabricated for length-only alleles)."""
        sl = self.allele_slice(l)
        out = []
        m = self.metas[l]
        motif = self.motif(l)
        for j, a in enumerate(range(sl.start, sl.stop)):
            tl = int(self.h["trim_len"][a])
            if m.given_len is not None and not np.isnan(m.given_len[j]):
                reps = tl // max(len(motif), 1) + 1
                out.append((motif * reps)[:tl])
            else:
                o = int(self.allele_off[a]) + int(self.h["trim_off"][a])
                out.append(self.seqs[o:o + tl].decode("ascii").upper())
        return out

    def motif(self, l: int) -> str:
        o0, o1 = int(self.h["motif_off"][l]), int(self.h["motif_off"][l + 1])
        return self.h["motif"][o0:o1].decode("ascii")


def _native_run(records):
    """(native block, first index) if the records are a consecutive run of one block of the C++
    reader (vcf_ingest) whose GT arrays are still the reader's own parse; else None."""
    if not records or not hasattr(records[0], "native_slot"):
        return None
    first = records[0].native_slot()
    if first is None:
        return None
    nblk, i0 = first
    for j, r in enumerate(records):
        slot = r.native_slot() if hasattr(r, "native_slot") else None
        if slot is None or slot[0] is not nblk or slot[1] != i0 + j:
            return None
    if nblk.S == 0 or int(nblk.rec_ploidy[i0:i0 + len(records)].max()) != nblk.ploidy:
        return None
    return nblk, i0


def _stack_format(records, key):
    """[L][S][n] stack of record.format(key), or None if a record lacks the key."""
    cols = []
    for r in records:
        try:
            v = r.format(key)
        except KeyError:
            return None
        if v is None:
            return None
        cols.append(v)
    return np.stack(cols, axis=0) if cols else None


def build_block(ctx, vcftype: str, records: Sequence[Any], fmt_keys: Sequence[str] = ()) -> Block:
    """Stage a run of cyvcf2-like records (same ploidy) as one GPU block."""
    metas = [record_meta(vcftype, r) for r in records]
    run = _native_run(records)
    if run is not None:
        # the C++ reader already parsed the whole run into stacked arrays: no per-record pulls
        nblk, i0 = run
        i1 = i0 + len(records)
        nblk.parse(fmt_keys)
        fmt = {}
        for key in fmt_keys:
            if key in nblk.fmt and bool((nblk.present[key][i0:i1] == 1).all()):
                fmt[key] = nblk.fmt[key][i0:i1].reshape(len(records), nblk.S, 1)
                for j, r in enumerate(records):
                    r.share_format(key, fmt[key][j])       # what r.format(key) would have cached
            else:
                stacked = _stack_format(records, key)
                if stacked is not None:
                    fmt[key] = stacked
        if nblk.gt2 is not None:            # parsed straight into the packed transfer form
            blk = Block(ctx, vcftype, metas, None, fmt, gt_packed=(nblk.gt2[i0:i1], nblk.phase[i0:i1]))
        else:
            blk = Block(ctx, vcftype, metas, nblk.gt[i0:i1], fmt)
        blk._records = list(records)
        blk.rec_ploidy = np.ascontiguousarray(nblk.rec_ploidy[i0:i1], dtype=np.int32)
        return blk
    gts = []
    has_samples = True
    for r in records:
        g = r.genotype
        if g is None:
            has_samples = False
            break
        gts.append(g.array())
    gt = None
    if has_samples and gts:
        P = max(g.shape[1] for g in gts)
        S = gts[0].shape[0]
        gt = np.full((len(gts), S, P), -2, dtype=np.int16)
        for i, g in enumerate(gts):
            p = g.shape[1] - 1
            gt[i, :, :p] = g[:, :p]
            gt[i, :, P - 1] = g[:, p]
    fmt = {}
    for key in fmt_keys:
        stacked = _stack_format(records, key)
        if stacked is not None:
            fmt[key] = stacked
    blk = Block(ctx, vcftype, metas, gt, fmt)
    blk._records = list(records)
    if gts and gt is not None:
        blk.rec_ploidy = np.array([g.shape[1] - 1 for g in gts], dtype=np.int32)
    return blk
