"""
Text-VCF reader/writer exposing the slice of the ``cyvcf2`` API that the
TRTools hot path touches.

The reference delegates VCF decoding to cyvcf2/htslib, which is not part of
the reference tree and is not installed in this image.  This module is the
host-side stand-in used by the ``statSTR``/``dumpSTR``/``associaTR`` drop-ins
when the real ``cyvcf2`` cannot be imported; it follows the conventions the
reference relies on:

* ``Variant.genotype.array()`` -> int16 ``[S, P+1]``: allele indices, ``-1`` for
  ``.``, ``-2`` padding for lower-ploidy calls, last column phased 0/1
  (reference docstring ``trtools/utils/tr_harmonizer.py:829-859``).
* ``Variant.format(key)``: Integer -> int32 ``[S, n]`` with ``INT32_MIN`` for a
  missing value; Float -> float32 with NaN; String -> ``<U`` array ``[S]`` with
  ``'.'`` (``trtools/dumpSTR/dumpSTR.py:610,736-742``,
  ``trtools/dumpSTR/filters.py:365,446``).
* ``Variant.FILTER`` is ``None`` for ``.`` and ``PASS``
  (``trtools/statSTR/statSTR.py:580``).
* ``Variant.INFO``: ``get(key)`` / ``[key]`` / iteration yielding ``(key, value)``
  (``trtools/utils/tr_harmonizer.py:316,349-351,713``).

It is plain host code (no GPU work); the blocked ingest stage in
``trtools_b200.block`` turns its per-record arrays into the pinned ``[L][S]``
blocks the CUDA kernels consume.
"""
import gzip
import io
import re
from typing import Dict, Iterator, List, Optional

import numpy as np

INT32_MISSING = -2147483648
_VEC_MIN_SAMPLES = 64      # below this many samples the per-sample loop is as fast as the np.char path
INT32_VECTOR_END = -2147483647
# BCF's float vector-end marker (a NaN with payload 2; np.isnan is true for it, so "missing" logic is unchanged).  Copies
# keep the bits; float32 <-> float64 conversions may set the quiet bit, hence the payload test below.
FLOAT_VECTOR_END = np.array([0x7F800002], dtype=np.uint32).view(np.float32)[0]


def is_float_vector_end(a: np.ndarray) -> np.ndarray:
    """Boolean mask of BCF float vector-end entries of a float32 / float64 array."""
    a = np.asarray(a)
    if a.dtype == np.float32:
        u = a.view(np.uint32)
        return ((u & np.uint32(0x7F800000)) == np.uint32(0x7F800000)) & ((u & np.uint32(0x003FFFFF)) == np.uint32(2))
    if a.dtype == np.float64:
        u = a.view(np.uint64)
        return ((u & np.uint64(0x7FF0000000000000)) == np.uint64(0x7FF0000000000000)) & \
               ((u & np.uint64(0x0007FFFFFFFFFFFF)) == np.uint64(2 << 29))
    return np.zeros(a.shape, bool)

_header_kv = re.compile(r'([A-Za-z0-9_]+)=("(?:[^"\\]|\\.)*"|[^,]*)')


class HeaderRecord(dict):
    """One structured ``##KEY=<...>`` header line; ``rec['ID']`` style access."""

    def info(self, extra=False):
        return dict(self)


def _parse_header_line(line: str) -> Optional[HeaderRecord]:
    m = re.match(r'##([^=]+)=<(.*)>\s*$', line)
    if not m:
        return None
    rec = HeaderRecord()
    rec['HeaderType'] = m.group(1)
    for k, v in _header_kv.findall(m.group(2)):
        if len(v) >= 2 and v[0] == '"' and v[-1] == '"':
            v = v[1:-1]
        rec[k] = v
    return rec


def _open_text(path: str):
    with open(path, 'rb') as f:
        magic = f.read(2)
    if magic == b'\x1f\x8b':
        # BGZF is a valid multi-member gzip stream
        return io.TextIOWrapper(gzip.open(path, 'rb'), encoding='utf-8', newline='\n')
    return open(path, 'r', encoding='utf-8', newline='\n')


class _Info:
    """INFO column façade: typed by the header's Type/Number."""

    def __init__(self, text: str, types: Dict[str, HeaderRecord]):
        self._types = types
        self._d = {}
        self._order = []
        if text != '.' and text != '':
            for item in text.split(';'):
                if item == '':
                    continue
                if '=' in item:
                    k, v = item.split('=', 1)
                else:
                    k, v = item, True
                if k not in self._d:
                    self._order.append(k)
                self._d[k] = v
        self._cache = {}

    def _convert(self, key, raw):
        if raw is True:
            return True
        hrec = self._types.get(key)
        typ = hrec.get('Type', 'String') if hrec else 'String'
        number = hrec.get('Number', '1') if hrec else '1'
        if typ == 'Flag':
            return True
        if typ == 'String' or typ == 'Character':
            return raw
        conv = int if typ == 'Integer' else (lambda x: float(np.float32(x)))
        parts = raw.split(',')

        def one(x):
            if x == '.':
                return None
            try:
                return conv(x)
            except ValueError:
                return float(x) if typ == 'Integer' else x
        if number == '1' and len(parts) == 1:
            return one(parts[0])
        vals = tuple(one(x) for x in parts)
        if len(vals) == 1 and number not in ('A', 'R', 'G', '.'):
            return vals[0]
        return vals

    def get(self, key, default=None):
        if key not in self._d:
            return default
        return self[key]

    def __getitem__(self, key):
        if key not in self._d:
            raise KeyError(key)
        if key not in self._cache:
            self._cache[key] = self._convert(key, self._d[key])
        return self._cache[key]

    def __setitem__(self, key, value):
        if key not in self._d:
            self._order.append(key)
        if isinstance(value, bool):
            raw = True
        elif isinstance(value, (float, np.floating)):
            raw = '%g' % value
        elif isinstance(value, (tuple, list)):
            raw = ','.join(str(v) for v in value)
        else:
            raw = str(value)
        self._d[key] = raw
        self._cache[key] = value

    def __delitem__(self, key):
        del self._d[key]
        self._order.remove(key)
        self._cache.pop(key, None)

    def __contains__(self, key):
        return key in self._d

    def __iter__(self):
        for k in self._order:
            yield (k, self[k])

    def to_text(self):
        if not self._order:
            return '.'
        out = []
        for k in self._order:
            v = self._d[k]
            out.append(k if v is True else '{}={}'.format(k, v))
        return ';'.join(out)


class Genotypes:
    """Mirror of ``cyvcf2.Genotypes`` (only ``array()``/``n_samples``/``ploidy``)."""

    def __init__(self, arr: np.ndarray):
        self._arr = arr
        self.n_samples = arr.shape[0]
        self.ploidy = arr.shape[1] - 1

    def array(self):
        return self._arr.copy()


_gt_split = re.compile(r'[/|]')


def parse_gt_column(gt_strings: List[str]) -> np.ndarray:
    """GT strings of one record -> int16 [S, P+1] in cyvcf2's layout."""
    n = len(gt_strings)
    # fast path: every call is diploid (or a lone '.'); ploidy 2
    ploidy = 1
    split = []
    for g in gt_strings:
        parts = _gt_split.split(g)
        split.append(parts)
        if len(parts) > ploidy:
            ploidy = len(parts)
    arr = np.full((n, ploidy + 1), -2, dtype=np.int16)
    arr[:, ploidy] = 0
    for i, (g, parts) in enumerate(zip(gt_strings, split)):
        for j, a in enumerate(parts):
            arr[i, j] = -1 if (a == '.' or a == '') else int(a)
        if len(parts) > 1:
            # cyvcf2 reports the phase bit carried by the second allele
            sep_pos = len(parts[0])
            arr[i, ploidy] = 1 if g[sep_pos] == '|' else 0
    return arr


class Variant:
    """One VCF data line."""

    def __init__(self, line: str, vcf: 'VCF'):
        self._vcf = vcf
        cols = line.rstrip('\n').rstrip('\r').split('\t')
        self._init_fixed(cols)
        if len(cols) > 8:
            self._sample_cols = cols[9:]
            if vcf._sample_idx is not None:
                self._sample_cols = [self._sample_cols[i] for i in vcf._sample_idx]
        else:
            self._sample_cols = []

    def _init_fixed(self, cols):
        """CHROM..FORMAT from the first (up to) nine columns."""
        if len(cols) < 8:
            raise ValueError("malformed VCF line")
        self.CHROM = cols[0]
        self.POS = int(cols[1])
        self.ID = None if cols[2] == '.' else cols[2]
        self.REF = cols[3]
        self.ALT = [] if cols[4] == '.' else cols[4].split(',')
        self.QUAL = None if cols[5] == '.' else float(cols[5])
        self._filter_raw = cols[6]
        self.INFO = _Info(cols[7], self._vcf._info_types)
        self.FORMAT = cols[8].split(':') if len(cols) > 8 else []
        self._split = None
        self._fmt_cache = {}
        self._gt_arr = None
        self._genotypes = None

    # -- FILTER ----------------------------------------------------------
    @property
    def FILTER(self):
        if self._filter_raw in ('.', 'PASS'):
            return None
        return self._filter_raw

    @FILTER.setter
    def FILTER(self, value):
        if value is None:
            self._filter_raw = '.'
        elif isinstance(value, (list, tuple)):
            self._filter_raw = ';'.join(value)
        else:
            self._filter_raw = str(value)

    @property
    def FILTERS(self):
        return self._filter_raw.split(';')

    # -- samples ---------------------------------------------------------
    def _fields(self):
        if self._split is None:
            self._split = [s.split(':') for s in self._sample_cols]
        return self._split

    def _raw_field(self, key) -> Optional[List[str]]:
        if key not in self.FORMAT:
            return None
        idx = self.FORMAT.index(key)
        return [f[idx] if idx < len(f) else '.' for f in self._fields()]

    def _gts(self):
        if self._gt_arr is None:
            raw = self._raw_field('GT')
            if raw is None:
                self._gt_arr = np.full((len(self._sample_cols), 3), -1, dtype=np.int16)
                self._gt_arr[:, -1] = 0
            else:
                self._gt_arr = parse_gt_column(raw)
        return self._gt_arr

    @property
    def genotype(self):
        if len(self._vcf.samples) == 0:
            return None
        return Genotypes(self._gts())

    @property
    def ploidy(self):
        return self._gts().shape[1] - 1

    @property
    def genotypes(self):
        if self._genotypes is None:
            arr = self._gts()
            p = arr.shape[1] - 1
            out = []
            for row in arr:
                alleles = [int(a) for a in row[:p] if a != -2]
                out.append(alleles + [bool(row[p])])
            self._genotypes = out
        return self._genotypes

    @genotypes.setter
    def genotypes(self, gts):
        p = max((len(g) - 1 for g in gts), default=2)
        arr = np.full((len(gts), p + 1), -2, dtype=np.int16)
        for i, g in enumerate(gts):
            for j, a in enumerate(g[:-1]):
                arr[i, j] = a
            arr[i, p] = 1 if g[-1] else 0
        self._gt_arr = arr
        self._genotypes = [list(g) for g in gts]

    def format(self, key, vtype=None):
        if key in self._fmt_cache:
            return self._fmt_cache[key]
        raw = self._raw_field(key)
        if raw is None:
            raise KeyError(key)
        hrec = self._vcf._format_types.get(key)
        typ = hrec.get('Type', 'String') if hrec else 'String'
        if typ == 'String' or typ == 'Character':
            out = np.array(raw, dtype=str) if len(raw) else np.empty((0,), dtype='<U1')
        else:
            parts = [r.split(',') for r in raw]
            ncol = max((len(p) for p in parts), default=1)
            number = hrec.get('Number', '1') if hrec else '1'
            if str(number).isdigit():      # htslib sizes fixed-Number fields from the header
                ncol = max(ncol, int(number))
            if typ == 'Integer':
                out = np.full((len(raw), ncol), INT32_VECTOR_END, dtype=np.int32)
                for i, p in enumerate(parts):
                    for j, x in enumerate(p):
                        out[i, j] = INT32_MISSING if x in ('.', '') else int(x)
            else:
                out = np.full((len(raw), ncol), FLOAT_VECTOR_END, dtype=np.float32)      # ragged rows end like htslib's
                for i, p in enumerate(parts):
                    for j, x in enumerate(p):
                        out[i, j] = np.nan if x in ('.', '') else np.float32(x)
        self._fmt_cache[key] = out
        return out

    def set_format(self, key, data):
        data = np.asarray(data)
        if key not in self.FORMAT:
            self.FORMAT.append(key)
        if data.dtype.kind == 'S':
            data = np.char.decode(data)
        self._fmt_cache[key] = data

    # -- text ------------------------------------------------------------
    @staticmethod
    def _fmt_scalar(x):
        if isinstance(x, (np.integer, int)):
            if x == INT32_MISSING:
                return '.'
            return str(int(x))
        if isinstance(x, (np.floating, float)):
            if np.isnan(x):
                return '.'
            return '%g' % x
        return str(x)

    # -- vectorised serialisation of the sample columns (np.char over all samples of a field at once; the loop
    #    version below is the definition and the checker: tests/test_ingest.py compares the two byte for byte)
    @staticmethod
    def _gt_text_vec(arr):
        p = arr.shape[1] - 1
        if p < 1 or (arr[:, 0] == -2).any():
            return None
        sep = np.where(arr[:, p] != 0, '|', '/')
        out = np.where(arr[:, 0] == -1, '.', arr[:, 0].astype(str))
        for j in range(1, p):
            a = arr[:, j]
            nxt = np.char.add(np.char.add(out, sep), np.where(a == -1, '.', a.astype(str)))
            out = np.where(a != -2, nxt, out)
        return out

    @staticmethod
    def _column_text_vec(col):
        """One column of a FORMAT array -> (strings, skip mask); skip = INT32_VECTOR_END entries."""
        if col.dtype.kind in 'iu':
            skip = (col == INT32_VECTOR_END) if col.dtype.itemsize >= 4 else np.zeros(col.shape, bool)
            return np.where(col == INT32_MISSING, '.', col.astype(str)), skip
        if col.dtype.kind == 'f':
            nan = np.isnan(col)
            txt = np.char.mod('%g', np.where(nan, 0, col).astype(col.dtype))
            return np.where(nan, '.', txt), is_float_vector_end(np.ascontiguousarray(col))
        if col.dtype.kind == 'U':
            return col, np.zeros(col.shape, bool)
        return None, None

    @classmethod
    def _array_text_vec(cls, data):
        if data.dtype.kind not in 'iufU':
            return None
        if data.ndim == 1:
            if data.dtype.kind in 'iu' and data.dtype.itemsize < 4:
                return None
            txt, _ = cls._column_text_vec(data)
            return txt
        if data.ndim != 2:
            return None
        if data.shape[1] == 0:
            return np.full(data.shape[0], '.')
        out, have = None, None
        for j in range(data.shape[1]):
            txt, skip = cls._column_text_vec(data[:, j])
            if txt is None:
                return None
            if out is None:
                out, have = np.where(skip, '', txt), ~skip
            else:
                joined = np.where(have, np.char.add(np.char.add(out, ','), txt), txt)
                out = np.where(skip, out, joined)
                have = have | ~skip
        return np.where(have, out, '.')

    def _sample_text(self):
        n = len(self._sample_cols)
        if n == 0:
            return []
        if n < _VEC_MIN_SAMPLES:
            return self._sample_text_loop()
        fields = []
        for key in self.FORMAT:
            if key == 'GT':
                txt = self._gt_text_vec(self._gts())
            elif key in self._fmt_cache:
                txt = self._array_text_vec(np.asarray(self._fmt_cache[key]))
            else:
                txt = np.array(self._raw_field(key), dtype=str)
            if txt is None or txt.shape != (n,):
                return self._sample_text_loop()
            fields.append(txt)
        out = fields[0]
        for txt in fields[1:]:
            out = np.char.add(np.char.add(out, ':'), txt)
        return out.tolist()

    def _sample_text_loop(self):
        n = len(self._sample_cols)
        if n == 0:
            return []
        per_field = []
        for key in self.FORMAT:
            if key == 'GT':
                arr = self._gts()
                p = arr.shape[1] - 1
                vals = []
                for row in arr:
                    sep = '|' if row[p] else '/'
                    al = ['.' if a == -1 else str(int(a)) for a in row[:p] if a != -2]
                    vals.append(sep.join(al) if al else '.')
                per_field.append(vals)
            elif key in self._fmt_cache:
                data = self._fmt_cache[key]
                if data.ndim == 1:
                    per_field.append([self._fmt_scalar(x) for x in data])
                else:
                    vals = []
                    for row in data:
                        ve = is_float_vector_end(np.ascontiguousarray(row)) if row.dtype.kind == 'f' else None
                        items = [self._fmt_scalar(x) for j, x in enumerate(row)
                                 if not (isinstance(x, np.integer) and x == INT32_VECTOR_END) and not (ve is not None and ve[j])]
                        vals.append(','.join(items) if items else '.')
                    per_field.append(vals)
            else:
                per_field.append(self._raw_field(key))
        return [':'.join(f[i] for f in per_field) for i in range(n)]

    def __str__(self):
        cols = [self.CHROM, str(self.POS), self.ID if self.ID is not None else '.',
                self.REF, ','.join(self.ALT) if self.ALT else '.',
                '.' if self.QUAL is None else '%g' % self.QUAL,
                self._filter_raw, self.INFO.to_text()]
        if self.FORMAT:
            cols.append(':'.join(self.FORMAT))
            if len(self._sample_cols) >= _VEC_MIN_SAMPLES:
                joined = self._sample_text_native()
                if joined is not None:
                    return '\t'.join(cols) + '\t' + joined + '\n'
            cols.extend(self._sample_text())
        return '\t'.join(cols) + '\n'

    def _sample_text_native(self):
        """All sample columns as one TAB-joined string through trt_vcf_join_samples (csrc/trt_ingest.cpp), or
        None when a field is of a kind the C++ serialiser does not take (the np.char / loop versions do)."""
        import ctypes as C
        try:
            from . import _lib
            lib = _lib.load()
        except (ImportError, OSError, AttributeError):
            return None
        n = len(self._sample_cols)
        kinds, ncols, arrays = [], [], []
        est = 0
        for key in self.FORMAT:
            if key == 'GT':
                a = np.ascontiguousarray(self._gts(), dtype=np.int16)
                kind, nc, width = 1, a.shape[1], 7 * a.shape[1]
            elif key in self._fmt_cache:
                a = np.asarray(self._fmt_cache[key])
                if a.ndim == 2 and a.shape[1] >= 1 and a.dtype in (np.int32, np.float32, np.float64):
                    a = np.ascontiguousarray(a)
                    kind = 2 if a.dtype == np.int32 else (3 if a.dtype == np.float32 else 4)
                    nc, width = a.shape[1], (12 if kind == 2 else 26) * a.shape[1]
                elif a.ndim == 1 and a.dtype.kind in 'US':
                    try:
                        a = np.ascontiguousarray(np.char.encode(a, 'utf-8') if a.dtype.kind == 'U' else a)
                    except UnicodeError:
                        return None
                    kind, nc, width = 0, a.dtype.itemsize, a.dtype.itemsize
                else:
                    return None
            else:
                raw_bytes = getattr(self, '_raw_field_bytes', None)      # records of the C++ reader
                a = raw_bytes(key) if raw_bytes is not None else None
                if a is None:
                    try:
                        a = np.array(self._raw_field(key), dtype='S')
                    except UnicodeError:
                        return None
                kind, nc, width = 0, a.dtype.itemsize, a.dtype.itemsize
            if a.shape[0] != n:
                return None
            kinds.append(kind)
            ncols.append(nc)
            arrays.append(a)
            est += width + 1
        nf = len(kinds)
        c_kind = (C.c_int32 * nf)(*kinds)
        c_ncol = (C.c_int32 * nf)(*ncols)
        c_data = (C.c_void_p * nf)(*[a.ctypes.data for a in arrays])
        cap = n * est + 16
        for _ in range(2):
            buf = C.create_string_buffer(cap)
            w = lib.trt_vcf_join_samples(n, nf, c_kind, c_data, c_ncol, buf, cap)
            if w > 0:
                try:
                    return buf.raw[:w].decode('utf-8')
                except UnicodeDecodeError:
                    return None
            if w == 0:
                return None
            cap = -w + 16
        return None


class TextVCF:
    """Sequential text-VCF reader in pure Python (``cyvcf2.VCF`` surface used by TRTools).

    The drop-ins read through :class:`trtools_b200.vcf_ingest.NativeVCF` (exported below as ``VCF``),
    which parses blocks of records in C++; this class is what it re-parses flagged records with, and
    the reader the ingest parity tests compare against."""

    def __init__(self, fname, mode='r', gts012=False, lazy=False, strict_gt=False,
                 samples=None, threads=None):
        self.fname = str(fname)
        try:
            self._fh = _open_text(self.fname)
        except (OSError, IOError):
            raise OSError("Error opening %s" % fname)
        header_lines = []
        self._first_data = None
        try:
            for line in self._fh:
                if line.startswith('#'):
                    header_lines.append(line)
                else:
                    self._first_data = line
                    break
        except (OSError, UnicodeDecodeError, EOFError):
            raise OSError("Error reading %s" % fname)
        self._init_header(header_lines, samples)

    def _init_header(self, header_lines, samples):
        if not header_lines or not header_lines[0].startswith('##fileformat'):
            if not any(l.startswith('#CHROM') for l in header_lines):
                raise OSError("%s is not a VCF" % self.fname)
        self._header_lines = header_lines
        self._index_types()
        chrom_line = [l for l in header_lines if l.startswith('#CHROM')]
        cols = chrom_line[-1].rstrip('\n').rstrip('\r').split('\t') if chrom_line else []
        all_samples = cols[9:]
        self._sample_idx = None
        if samples is not None:
            if isinstance(samples, str):
                samples = samples.split(',')
            keep = set(samples)
            self._sample_idx = [i for i, s in enumerate(all_samples) if s in keep]
            all_samples = [all_samples[i] for i in self._sample_idx]
        self.samples = all_samples
        self._region = None

    def _index_types(self):
        self._info_types = {}
        self._format_types = {}
        self._hrecs = []
        for line in self._header_lines:
            rec = _parse_header_line(line)
            if rec is None:
                continue
            self._hrecs.append(rec)
            if rec['HeaderType'] == 'INFO' and 'ID' in rec:
                self._info_types[rec['ID']] = rec
            elif rec['HeaderType'] == 'FORMAT' and 'ID' in rec:
                self._format_types[rec['ID']] = rec

    @property
    def raw_header(self):
        return ''.join(self._header_lines)

    def header_iter(self):
        return iter(self._hrecs)

    @property
    def seqnames(self):
        return [r['ID'] for r in self._hrecs if r['HeaderType'].lower() == 'contig']

    def add_to_header(self, line):
        line = line.rstrip('\n') + '\n'
        self._header_lines.insert(len(self._header_lines) - 1, line)
        self._index_types()

    def _add_structured(self, kind, adict, keys):
        parts = []
        for k in keys:
            if k in adict:
                v = adict[k]
                if k == 'Description':
                    v = '"{}"'.format(v)
                parts.append('{}={}'.format(k, v))
        self.add_to_header('##{}=<{}>'.format(kind, ','.join(parts)))

    def add_info_to_header(self, adict):
        self._add_structured('INFO', adict, ['ID', 'Number', 'Type', 'Description'])

    def add_format_to_header(self, adict):
        self._add_structured('FORMAT', adict, ['ID', 'Number', 'Type', 'Description'])

    def add_filter_to_header(self, adict):
        self._add_structured('FILTER', adict, ['ID', 'Description'])

    def __iter__(self) -> Iterator[Variant]:
        return self

    def _next_line(self):
        if self._first_data is not None:
            line, self._first_data = self._first_data, None
            return line
        line = self._fh.readline()
        while line == '\n':
            line = self._fh.readline()
        return line

    def _in_region(self, var) -> bool:
        if self._region is None:
            return True
        chrom, start, end = self._region
        if var.CHROM != chrom:
            return False
        vend = var.POS + len(var.REF) - 1
        if start is not None and vend < start:
            return False
        if end is not None and var.POS > end:
            return False
        return True

    def __next__(self) -> Variant:
        while True:
            line = self._next_line()
            if not line:
                raise StopIteration
            var = Variant(line, self)
            if self._in_region(var):
                return var

    def __call__(self, region: str):
        """Region query by linear scan (real cyvcf2 uses the tabix index)."""
        chrom, start, end = region, None, None
        if ':' in region:
            chrom, rng = region.rsplit(':', 1)
            rng = rng.replace(',', '')
            if '-' in rng:
                s, e = rng.split('-', 1)
                start = int(s)
                end = int(e) if e else None
            else:
                start = int(rng)
        self._region = (chrom, start, end)
        return self

    def close(self):
        try:
            self._fh.close()
        except Exception:
            pass


_BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
_BGZF_MAX = 0xff00          # uncompressed bytes per member (htslib's block size)


class BgzfWriter:
    """Blocked gzip as htslib / bgzip write it: members of <= 65 280 inflated bytes, each carrying its compressed
    size in the ``BC`` extra field, then the empty EOF member — what ``tabix`` (dumpSTR ``--zip``, reference
    dumpSTR.py:1347-1352) and the region seeks of csrc/trt_ingest.cpp need.  Plain ``gzip.open`` output is rejected
    by both.  ``tell()`` is the BGZF virtual offset of the next byte (member offset << 16 | offset within it)."""

    def __init__(self, path, level: int = 6):
        self._fh = open(path, "wb")
        self._buf = bytearray()
        self._level = level
        self._coffset = 0

    def write(self, data: bytes):
        self._buf += data
        while len(self._buf) >= _BGZF_MAX:
            self._flush_block(bytes(self._buf[:_BGZF_MAX]))
            del self._buf[:_BGZF_MAX]

    def tell(self) -> int:
        return (self._coffset << 16) | len(self._buf)

    def _flush_block(self, data: bytes):
        import struct
        import zlib
        c = zlib.compressobj(self._level, zlib.DEFLATED, -15)
        comp = c.compress(data) + c.flush()
        if len(comp) + 26 > 65536:                      # incompressible input: store
            c = zlib.compressobj(0, zlib.DEFLATED, -15)
            comp = c.compress(data) + c.flush()
        block = (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(comp) + 25) + comp +
                 struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data)))
        self._fh.write(block)
        self._coffset += len(block)

    def close(self):
        if self._fh is None:
            return
        if self._buf:
            self._flush_block(bytes(self._buf))
            self._buf = bytearray()
        self._fh.write(_BGZF_EOF)
        self._fh.close()
        self._fh = None


class Writer:
    """Text VCF writer (``cyvcf2.Writer`` surface used by dumpSTR); ``.gz`` names are written as BGZF."""

    def __init__(self, fname, tmpl: 'TextVCF', mode=None):
        self.fname = str(fname)
        self._bgzf = self.fname.endswith('.gz')
        if self._bgzf:
            self._fh = BgzfWriter(self.fname)
        else:
            self._fh = open(self.fname, 'w', encoding='utf-8', newline='\n')
        self._tmpl = tmpl
        self._header_written = False

    def _write(self, text: str):
        self._fh.write(text.encode('utf-8') if self._bgzf else text)

    def write_header(self):
        if not self._header_written:
            self._write(self._tmpl.raw_header)
            self._header_written = True

    def write_record(self, variant: Variant):
        self.write_header()
        self._write(str(variant))

    def write_text(self, text: str):
        """Records already serialised (the multi-GPU dumpSTR gathers every rank's records as text on rank 0)."""
        self.write_header()
        self._write(text)

    def close(self):
        self.write_header()
        self._fh.close()


# ``VCF`` is what the drop-ins open: the block reader of vcf_ingest (C++ inflate + FORMAT parse,
# csrc/trt_ingest.cpp).  TRTOOLS_B200_INGEST=python selects the pure-Python reader instead.
def __getattr__(name):
    if name == "VCF":
        import os
        if os.environ.get("TRTOOLS_B200_INGEST", "native") == "python":
            return TextVCF
        from .vcf_ingest import NativeVCF
        return NativeVCF
    raise AttributeError(name)
