"""
Native block VCF ingest (SURVEY.md §8f "next" row 1): the ``cyvcf2.VCF`` surface of
:mod:`trtools_b200.cyvcf2_compat`, with the per-record text work moved into C++
(``csrc/trt_ingest.cpp`` behind ``trt_vcf_*`` in ``include/trtools_b200.h``).

The reference pulls ``genotype.array()`` and ``format(key)`` once per record through cyvcf2/htslib
(trtools/utils/tr_harmonizer.py:829-862, 561-588).  Here the reader inflates BGZF members in
parallel, keeps a run of records as text, and parses GT and the numeric scalar FORMAT keys of the
whole run in one multi-threaded pass into stacked ``[L][S]`` arrays.  ``Variant`` objects are thin
views: the first nine columns are decoded eagerly (INFO validation, allele strings), the sample
columns only if something asks for a string-typed or vector-valued field.  ``block.build_block``
takes the stacked arrays of a run as they are, without per-record copies.

Whatever the C++ pass flags (odd GT tokens, vector-valued values in a scalar field, ragged sample
columns) is re-parsed from the record's text by the pure-Python reader, so results are those of
``cyvcf2_compat.TextVCF`` byte for byte (tests/test_ingest.py).
"""
import ctypes as C
import os
import threading
from typing import Dict, Optional, Sequence

import numpy as np

from . import _lib
from . import cyvcf2_compat as _compat

_DEFAULT_BLOCK_LOCI = 512
_DEFAULT_BLOCK_BYTES = 1 << 30


def _tabix_start(tbi_path: str, chrom: str, start: Optional[int]) -> Optional[int]:
    """Virtual file offset at which records overlapping ``chrom:start-`` can begin, from the linear index of a
    tabix file (one offset per 16 kb window: the smallest offset of any record overlapping the window).
    None: no usable index.  -1: the index shows there is nothing to read (contig absent / window past the end)."""
    import gzip
    import struct
    if not os.path.isfile(tbi_path):
        return None
    with gzip.open(tbi_path, "rb") as f:
        data = f.read()
    if data[:4] != b"TBI\x01":
        return None
    n_ref, fmt, col_seq, col_beg, col_end, meta, skip, l_nm = struct.unpack_from("<8i", data, 4)
    names = data[36:36 + l_nm].split(b"\x00")[:n_ref]
    if (fmt & 0xffff) != 2:              # not a VCF-preset index
        return None
    try:
        tid = names.index(chrom.encode())
    except ValueError:
        return -1
    o = 36 + l_nm
    for r in range(n_ref):
        (n_bin,) = struct.unpack_from("<i", data, o)
        o += 4
        for _ in range(n_bin):
            _bin, n_chunk = struct.unpack_from("<Ii", data, o)
            o += 8 + 16 * n_chunk
        (n_intv,) = struct.unpack_from("<i", data, o)
        o += 4
        if r == tid:
            if n_intv == 0:
                return -1
            w = max((start or 1) - 1, 0) >> 14
            if w >= n_intv:
                return -1
            ioff = struct.unpack_from("<%dQ" % n_intv, data, o)
            # gaps are stored as 0: the next filled window (or any earlier one) bounds the start from below
            k = w
            while k < n_intv and ioff[k] == 0:
                k += 1
            if k == n_intv:
                return -1
            return int(ioff[k])
        o += 8 * n_intv
    return None


def _csi_start(csi_path: str, chrom: str, start: Optional[int], end: Optional[int] = None) -> Optional[int]:
    """The same lower bound from a CSI index (``bcftools index`` / ``tabix -C``; needed for contigs beyond 512 Mb), the
    way htslib's iterator finds it: ``min_off`` = ``loffset`` of the deepest existing bin at or left of the one holding
    ``start`` (leaf bin of the position, then previous siblings, then the parent, up to bin 0); the answer is the
    smallest chunk start among the bins overlapping the region whose chunk ends past ``min_off``.  Same return
    convention as _tabix_start (-1: the index shows the region holds nothing)."""
    import gzip
    import struct
    if not os.path.isfile(csi_path):
        return None
    with gzip.open(csi_path, "rb") as f:
        data = f.read()
    if data[:4] != b"CSI\x01":
        return None
    min_shift, depth, l_aux = struct.unpack_from("<3i", data, 4)
    if l_aux < 28:
        return None                      # no tabix header in the auxiliary block: contig names unknown
    fmt, _cs, _cb, _ce, _meta, _skip, l_nm = struct.unpack_from("<7i", data, 16)
    if (fmt & 0xffff) != 2:
        return None
    names = data[44:44 + l_nm].split(b"\x00")
    o = 16 + l_aux
    (n_ref,) = struct.unpack_from("<i", data, o)
    o += 4
    names = names[:n_ref]
    try:
        tid = names.index(chrom.encode())
    except ValueError:
        return -1
    for r in range(n_ref):
        (n_bin,) = struct.unpack_from("<i", data, o)
        o += 4
        bins = {}
        for _ in range(n_bin):
            b, lo, n_chunk = struct.unpack_from("<IQi", data, o)
            o += 16
            if r == tid:
                bins[b] = (lo, [struct.unpack_from("<QQ", data, o + 16 * i) for i in range(n_chunk)])
            o += 16 * n_chunk
        if r != tid:
            continue
        limit = ((1 << (3 * (depth + 1))) - 1) // 7          # first id past the real bins (pseudo-bins sit above)
        real = {b: v for b, v in bins.items() if b < limit}
        if not real:
            return -1
        span = 1 << (min_shift + 3 * depth)                  # positions the index can address
        beg = max((start or 1) - 1, 0)
        stop = min(end if end is not None else span, span)
        if beg >= span or stop <= beg:
            return -1
        b = ((1 << (3 * depth)) - 1) // 7 + (beg >> min_shift)
        while b:
            if b in real:
                break
            first_sibling = (((b - 1) >> 3) << 3) + 1
            b = b - 1 if b > first_sibling else (b - 1) >> 3
        min_off = real[b][0] if b in real else 0
        best = None
        shift, t = min_shift + 3 * depth, 0
        for level in range(depth + 1):
            for bb in range(t + (beg >> shift), t + ((stop - 1) >> shift) + 1):
                if bb in real:
                    for cb, ce in real[bb][1]:
                        if ce > min_off and (best is None or cb < best):
                            best = cb
            shift -= 3
            t += 1 << (3 * level)
        if best is None:
            return -1
        return int(max(best, min_off))
    return None


class _NativeBlock:
    """One run of records held by the C++ reader (``trt_vcf_block``) and its parsed arrays."""

    def __init__(self, vcf: "NativeVCF", handle, n: int):
        self.vcf = vcf
        self.lib = vcf._lib
        self.h = handle
        self.n = n
        text, off, fixed = C.c_void_p(), C.c_void_p(), C.c_void_p()
        rc = self.lib.trt_vcf_block_text(self.h, C.byref(text), C.byref(off), C.byref(fixed))
        if rc != _lib.TRT_OK:
            raise OSError("trt_vcf_block_text failed")
        self._text = text.value
        self.line_off = np.ctypeslib.as_array((C.c_int64 * (n + 1)).from_address(off.value)).copy()
        self.fixed_len = np.ctypeslib.as_array((C.c_int64 * n).from_address(fixed.value)).copy()
        self.S = len(vcf.samples)
        self._gt: Optional[np.ndarray] = None      # int16 [n][S][P+1] (plain parse, or materialised from the packed form)
        self.gt2: Optional[np.ndarray] = None      # transfer form: nibble pairs uint8 [n][S] or two-byte uint8 [n][S][2] ...
        self.phase: Optional[np.ndarray] = None    # ... + phase bits uint8 [n][ceil(S/8)]
        self.rec_ploidy: Optional[np.ndarray] = None
        self.status: Optional[np.ndarray] = None
        self.fmt: Dict[str, np.ndarray] = {}       # key -> [n][S]
        self.present: Dict[str, np.ndarray] = {}   # key -> uint8 [n]

    def __del__(self):
        try:
            if self.h:
                self.lib.trt_vcf_block_free(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def parsed(self) -> bool:
        return self.status is not None

    @property
    def ploidy(self) -> int:
        return 2 if self.gt2 is not None else self._gt.shape[2] - 1

    @property
    def gt(self) -> Optional[np.ndarray]:
        """cyvcf2-layout int16 [n][S][P+1] of the whole run (materialised on demand when the run was parsed into the
        packed transfer form: the GPU path uploads the packed arrays and never asks for this)."""
        if self._gt is None and self.gt2 is not None:
            from .block import unpack_any
            self._gt = unpack_any(self.gt2, self.phase)
        return self._gt

    # ---- text ----------------------------------------------------------------------------------
    def prefix(self, i: int) -> str:
        return C.string_at(self._text + int(self.line_off[i]), int(self.fixed_len[i])).decode('utf-8')

    def line(self, i: int) -> str:
        o0, o1 = int(self.line_off[i]), int(self.line_off[i + 1])
        return C.string_at(self._text + o0, o1 - o0).decode('utf-8').split('\n', 1)[0]

    # ---- arrays --------------------------------------------------------------------------------
    def _native_key(self, key: str) -> Optional[int]:
        """0 / 1 (is_float) if the header types the key as a numeric scalar, else None."""
        hrec = self.vcf._format_types.get(key)
        if not hrec or str(hrec.get('Number', '1')) != '1':
            return None
        typ = hrec.get('Type', 'String')
        if typ == 'Integer':
            return 0
        if typ == 'Float':
            return 1
        return None

    def parse(self, keys: Sequence[str] = ()):
        """One C++ pass: GT (first time) + the numeric scalar keys not parsed yet."""
        todo = [k for k in dict.fromkeys(keys) if k not in self.fmt and self._native_key(k) is not None][:32]
        want_gt = not self.parsed
        if not want_gt and not todo:
            return
        n, S = self.n, self.S
        outs = [np.empty((n, S), dtype=np.float32 if self._native_key(k) else np.int32) for k in todo]
        nk = len(todo)
        c_keys = (C.c_char_p * max(nk, 1))(*[k.encode() for k in todo])
        c_isf = (C.c_int32 * max(nk, 1))(*[int(self._native_key(k)) for k in todo])
        c_out = (C.c_void_p * max(nk, 1))(*[o.ctypes.data for o in outs])
        present = np.zeros((n, max(nk, 1)), dtype=np.uint8)
        rec_ploidy = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.uint8)
        P = 2
        gt2 = phase = None
        if want_gt and getattr(self.vcf, "_packed_gt", True) and getattr(self.vcf, "_nibble_gt", True) and \
                hasattr(self.lib, "trt_vcf_block_parse_nibble"):
            # one byte per call (alleles <= 13: nearly every TR locus) + a phase bit
            gt2 = np.empty((n, S), dtype=np.uint8)
            phase = np.empty((n, (S + 7) // 8), dtype=np.uint8)
            rc = self.lib.trt_vcf_block_parse_nibble(
                self.h, gt2.ctypes.data_as(C.c_void_p), phase.ctypes.data_as(C.c_void_p), nk,
                C.cast(c_keys, C.c_void_p), C.cast(c_isf, C.c_void_p), C.cast(c_out, C.c_void_p),
                present.ctypes.data_as(C.c_void_p), rec_ploidy.ctypes.data_as(C.c_void_p),
                status.ctypes.data_as(C.c_void_p))
            if rc != _lib.TRT_OK:
                raise OSError("trt_vcf_block_parse_nibble failed ({})".format(rc))
            if (status == 3).any():          # an allele index above 13 (or a polyploid call): the two-byte form below
                gt2 = phase = None
        if gt2 is None and want_gt and getattr(self.vcf, "_packed_gt", True) and hasattr(self.lib, "trt_vcf_block_parse_packed"):
            # straight into the packed transfer form (2 bytes per call + a phase bit): what the GPU block uploads
            gt2 = np.empty((n, S, 2), dtype=np.uint8)
            phase = np.empty((n, (S + 7) // 8), dtype=np.uint8)
            rc = self.lib.trt_vcf_block_parse_packed(
                self.h, gt2.ctypes.data_as(C.c_void_p), phase.ctypes.data_as(C.c_void_p), nk,
                C.cast(c_keys, C.c_void_p), C.cast(c_isf, C.c_void_p), C.cast(c_out, C.c_void_p),
                present.ctypes.data_as(C.c_void_p), rec_ploidy.ctypes.data_as(C.c_void_p),
                status.ctypes.data_as(C.c_void_p))
            if rc != _lib.TRT_OK:
                raise OSError("trt_vcf_block_parse_packed failed ({})".format(rc))
            if (status == 3).any():          # an allele index above 252 or a polyploid call: plain int16 parse below
                gt2 = phase = None
        while gt2 is None:
            gt = np.empty((n, S, P + 1), dtype=np.int16) if want_gt else None
            rc = self.lib.trt_vcf_block_parse(
                self.h, P, None if gt is None else gt.ctypes.data_as(C.c_void_p), nk,
                C.cast(c_keys, C.c_void_p), C.cast(c_isf, C.c_void_p), C.cast(c_out, C.c_void_p),
                present.ctypes.data_as(C.c_void_p), rec_ploidy.ctypes.data_as(C.c_void_p),
                status.ctypes.data_as(C.c_void_p))
            if rc != _lib.TRT_OK:
                raise OSError("trt_vcf_block_parse failed ({})".format(rc))
            if not want_gt:
                break
            ok = status == 0
            pmax = int(rec_ploidy[ok].max()) if ok.any() else 1
            if pmax <= P:
                break
            P = pmax            # a call with more alleles than the array holds: size it and redo the pass
        if want_gt:
            if gt2 is not None:
                self.gt2, self.phase = gt2, phase
            else:
                self._gt = gt
            self.rec_ploidy = np.maximum(rec_ploidy, 1)
            self.status = status
        for j, k in enumerate(todo):
            self.fmt[k] = outs[j]
            self.present[k] = present[:, j].copy()

    def gt_of(self, i: int) -> Optional[np.ndarray]:
        """cyvcf2-layout GT of record i ([S][p+1], p = the record's ploidy) or None if flagged."""
        if not self.parsed:
            self.parse(self.vcf._prefetch)
        if self.status[i] != 0:
            return None
        if self.gt2 is not None and self._gt is None:
            from .block import unpack_any
            g = unpack_any(self.gt2[i][None], self.phase[i][None])[0]
        else:
            g = self.gt[i]
        P = g.shape[1] - 1
        p = int(self.rec_ploidy[i])
        if p == P:
            return g
        return np.concatenate([g[:, :p], g[:, P:]], axis=1)

    def numeric(self, key: str, i: int) -> Optional[np.ndarray]:
        """[S][1] array of a numeric scalar FORMAT key of record i, or None (not handled natively)."""
        if self._native_key(key) is None:
            return None
        if key not in self.fmt:
            self.parse(tuple(self.vcf._prefetch) + (key,))
        if self.present[key][i] != 1 or (self.status is not None and self.status[i] != 0):
            return None
        # a copy: callers own what format() returns (dumpSTR nulls filtered calls in place)
        return self.fmt[key][i].reshape(self.S, 1).copy()


class NativeVariant(_compat.Variant):
    """A record of a native block: fixed columns decoded, sample columns parsed by the block."""

    def __init__(self, blk: _NativeBlock, i: int, vcf: "NativeVCF"):
        self._vcf = vcf
        self._nblk = blk
        self._nidx = i
        self._cols_cache = None
        self._gt_native = False
        self._shared = set()
        self._init_fixed(blk.prefix(i).split('\t'))

    @property
    def _sample_cols(self):
        if self._cols_cache is None:
            cols = self._nblk.line(self._nidx).rstrip('\r').split('\t')[9:]
            if self._vcf._sample_idx is not None:
                cols = [cols[j] for j in self._vcf._sample_idx]
            self._cols_cache = cols
        return self._cols_cache

    def _gts(self):
        if self._gt_arr is None:
            arr = self._nblk.gt_of(self._nidx) if 'GT' in self.FORMAT else None
            if arr is None:
                return super()._gts()
            self._gt_arr = arr
            self._gt_native = True
        return self._gt_arr

    @property
    def genotypes(self):
        return _compat.Variant.genotypes.fget(self)

    @genotypes.setter
    def genotypes(self, gts):
        _compat.Variant.genotypes.fset(self, gts)
        self._gt_native = False

    def share_format(self, key, view):
        """build_block took this key of the whole run from the block's arrays; keep the record's state as if
        format(key) had been called (the writer re-serialises the keys a record has decoded), without a copy."""
        if key not in self._fmt_cache:
            self._fmt_cache[key] = view
            self._shared.add(key)

    def format(self, key, vtype=None):
        if key in self._fmt_cache:
            if key in self._shared:
                # still the block's storage: the caller owns what format() returns (dumpSTR nulls filtered
                # calls in place), so hand out a private copy from here on
                self._shared.discard(key)
                self._fmt_cache[key] = self._fmt_cache[key].copy()
            return self._fmt_cache[key]
        if key not in self.FORMAT:
            raise KeyError(key)
        arr = self._nblk.numeric(key, self._nidx)
        if arr is None:
            return super().format(key, vtype)
        self._fmt_cache[key] = arr
        return arr

    def _raw_field_bytes(self, key):
        """Raw tokens of one FORMAT key for all samples as a numpy 'S' array (C++ walk of the record's text), or
        None when the record is one the C++ reader flagged."""
        if key not in self.FORMAT:
            return None
        blk = self._nblk
        if not blk.parsed:
            blk.parse(self._vcf._prefetch)
        if blk.status[self._nidx] != 0:
            return None
        idx = self.FORMAT.index(key)
        longest = C.c_int32(0)
        rc = blk.lib.trt_vcf_block_field(blk.h, self._nidx, idx, 0, None, C.byref(longest))
        if rc != _lib.TRT_OK:
            return None
        out = np.zeros(blk.S, dtype='S%d' % max(longest.value, 1))
        rc = blk.lib.trt_vcf_block_field(blk.h, self._nidx, idx, out.dtype.itemsize, out.ctypes.data_as(C.c_void_p),
                                         C.byref(longest))
        return out if rc == _lib.TRT_OK else None

    def set_format(self, key, data):
        self._shared.discard(key)
        super().set_format(key, data)

    def native_slot(self):
        """(block, index) if this record's GT is still the block's own parse, else None."""
        if self._gt_arr is not None and not self._gt_native:
            return None
        if not self._nblk.parsed:
            self._nblk.parse(self._vcf._prefetch)
        if self._nblk.status[self._nidx] != 0 or 'GT' not in self.FORMAT:
            return None
        return self._nblk, self._nidx


class NativeVCF(_compat.TextVCF):
    """``cyvcf2.VCF`` surface over the C++ block reader."""

    def __init__(self, fname, mode='r', gts012=False, lazy=False, strict_gt=False, samples=None, threads=None):
        self.fname = str(fname)
        self._lib = _lib.load()
        self._h = None
        h = C.c_void_p()
        rc = self._lib.trt_vcf_open(os.fsencode(self.fname), int(threads or 0), C.byref(h))
        if rc != _lib.TRT_OK:
            raise OSError("Error opening %s" % fname)
        self._h = h
        text, n = C.c_void_p(), C.c_int64()
        self._lib.trt_vcf_header(self._h, C.byref(text), C.byref(n))
        try:
            raw = C.string_at(text.value, n.value).decode('utf-8') if n.value else ''
        except UnicodeDecodeError:
            raise OSError("Error reading %s" % fname)
        header_lines = [l + '\n' for l in raw.split('\n') if l != ''] if raw else []
        if raw and not raw.endswith('\n') and header_lines:
            header_lines[-1] = header_lines[-1][:-1]
        self._init_header(header_lines, samples)
        if self._sample_idx is not None:
            idx = np.asarray(self._sample_idx, dtype=np.int64)
            rc = self._lib.trt_vcf_set_samples(self._h, idx.ctypes.data_as(C.c_void_p), len(idx))
            if rc != _lib.TRT_OK:
                raise OSError(self._err())
        self._blk: Optional[_NativeBlock] = None
        self._blk_i = 0
        self._prefetch: Sequence[str] = ()
        self._native_block_loci = _DEFAULT_BLOCK_LOCI
        self._native_block_bytes = _DEFAULT_BLOCK_BYTES
        # one run of read-ahead (TRTOOLS_B200_INGEST_READAHEAD=0 turns it off)
        self._readahead = os.environ.get("TRTOOLS_B200_INGEST_READAHEAD", "1") != "0"
        self._ra_thread = None
        self._ra_result = None
        self._region_stop = False        # region query served through a tabix index: stop at the region's end
        self._region_empty = False
        self._seen_region_chrom = False

    def _err(self):
        msg = self._lib.trt_vcf_last_error(self._h)
        return msg.decode() if msg else "native VCF reader error"

    def _read_block(self) -> Optional[_NativeBlock]:
        blk, n = C.c_void_p(), C.c_int64()
        rc = self._lib.trt_vcf_read_block(self._h, int(self._native_block_loci), int(self._native_block_bytes),
                                          C.byref(blk), C.byref(n))
        if rc != _lib.TRT_OK:
            raise OSError("Error reading {}: {}".format(self.fname, self._err()))
        self._started = True
        if n.value == 0:
            return None
        return _NativeBlock(self, blk, n.value)

    def _readahead_main(self):
        """Worker thread: read and parse the next run while the caller works on the current one (the C++
        calls release the GIL, so this overlaps with Python-side record handling and with GPU calls)."""
        try:
            blk = self._read_block()
            if blk is not None:
                blk.parse(self._prefetch)
            self._ra_result = (blk, None)
        except BaseException as e:      # handed to the consumer when it asks for this run
            self._ra_result = (None, e)

    def _next_block(self) -> bool:
        if self._ra_thread is not None:
            self._ra_thread.join()
            self._ra_thread = None
            blk, err = self._ra_result
            self._ra_result = None
            if err is not None:
                raise err
        else:
            blk = self._read_block()
        self._blk = blk
        self._blk_i = 0
        if blk is None:
            return False
        if self._readahead:
            self._ra_thread = threading.Thread(target=self._readahead_main, name="trt-vcf-readahead", daemon=True)
            self._ra_thread.start()
        return True

    def _past_region(self, var) -> bool:
        """Indexed (sorted) file: nothing after this record can be in the region."""
        chrom, _, end = self._region
        if var.CHROM != chrom:
            return self._seen_region_chrom
        self._seen_region_chrom = True
        return end is not None and var.POS > end

    def __next__(self):
        while True:
            if self._region_empty:
                raise StopIteration
            if self._blk is None or self._blk_i >= self._blk.n:
                if self._h is None or not self._next_block():
                    raise StopIteration
            blk, i = self._blk, self._blk_i
            self._blk_i += 1
            if blk.fixed_len[i] < 0:
                raise ValueError("malformed VCF line")
            if self._sample_idx is not None:
                # a sample subset of a ragged record fails in the text reader's constructor: same here
                if not blk.parsed:
                    blk.parse(self._prefetch)
                if blk.status[i] == 2:
                    var = _compat.Variant(blk.line(i), self)
                    if self._in_region(var):
                        return var
                    continue
            var = NativeVariant(blk, i, self)
            if self._in_region(var):
                return var
            if self._region_stop and self._past_region(var):
                self._region_empty = True
                raise StopIteration

    # ---- region queries ------------------------------------------------------------------------------
    def __call__(self, region: str):
        """``vcf(region)``: with a tabix index next to a BGZF file, seek to the first 16 kb window of the region
        and stop at its end (the file is sorted, or tabix would not have indexed it); otherwise the same linear
        scan as the text reader."""
        super().__call__(region)
        self._region_stop = False
        self._region_empty = False          # a reader serves any number of region queries, like cyvcf2's
        self._seen_region_chrom = False
        chrom, start, _ = self._region
        try:
            voff = _tabix_start(self.fname + ".tbi", chrom, start)
            if voff is None:
                voff = _csi_start(self.fname + ".csi", chrom, start, self._region[2])
        except Exception:
            voff = None          # no / unreadable index: linear scan
        if voff is None:
            return self
        if self._ra_thread is not None:
            self._ra_thread.join()
            self._ra_thread, self._ra_result = None, None
        self._blk = None
        if voff < 0:             # the index knows the contig is absent or the window is past its last record
            self._region_empty = True
            return self
        if voff > 0:
            rc = self._lib.trt_vcf_seek(self._h, voff >> 16, voff & 0xffff)
            if rc != _lib.TRT_OK:
                raise OSError("Error reading {}: {}".format(self.fname, self._err()))
            self._started = True
        elif getattr(self, "_started", False):
            # 0 = before any record of the file, but this reader has already moved: start over behind the header
            self._lib.trt_vcf_close(self._h)
            h = C.c_void_p()
            if self._lib.trt_vcf_open(os.fsencode(self.fname), 0, C.byref(h)) != _lib.TRT_OK:
                raise OSError("Error opening %s" % self.fname)
            self._h = h
            if self._sample_idx is not None:
                idx = np.asarray(self._sample_idx, dtype=np.int64)
                self._lib.trt_vcf_set_samples(self._h, idx.ctypes.data_as(C.c_void_p), len(idx))
        self._region_stop = True
        return self

    def close(self):
        if getattr(self, "_ra_thread", None) is not None:
            self._ra_thread.join()
            self._ra_thread = None
            self._ra_result = None
        if getattr(self, "_h", None):
            self._lib.trt_vcf_close(self._h)
            self._h = None
        self._blk = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
