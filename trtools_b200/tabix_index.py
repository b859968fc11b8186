"""
Tabix (.tbi) index writer for BGZF-compressed VCFs: what ``tabix -p vcf`` produces for dumpSTR ``--zip`` (reference
trtools/dumpSTR/dumpSTR.py:1347-1352 shells out to the ``tabix`` binary, which this image does not have).  The index
follows the tabix specification: per contig a binning index (UCSC bins, 14-bit minimum shift, 5 levels; chunks of
virtual offsets) and a linear index (smallest virtual offset of any record overlapping each 16 kb window), the whole
thing BGZF-compressed.  ``vcf_ingest._tabix_start`` / ``trt_vcf_seek`` read it back for ``vcf(region)``.
"""
import os
import struct
import zlib
from typing import Dict, List, Tuple

from .cyvcf2_compat import BgzfWriter

_MIN_SHIFT = 14


def _reg2bin(beg: int, end: int) -> int:
    """Bin of the zero-based half-open interval [beg, end) (tabix specification, section 'reg2bin')."""
    end -= 1
    if beg >> 14 == end >> 14:
        return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def _bgzf_members(path: str):
    """(compressed offset, inflated bytes) of every BGZF member, read one member at a time."""
    with open(path, "rb") as f:
        off = 0
        while True:
            head = f.read(12)
            if not head:
                return
            if len(head) < 12 or head[:4] != b"\x1f\x8b\x08\x04":
                raise OSError("{} is not BGZF".format(path))
            xlen = struct.unpack_from("<H", head, 10)[0]
            extra = f.read(xlen)
            bsize, p = None, 0
            while p + 4 <= len(extra):
                si, slen = extra[p:p + 2], struct.unpack_from("<H", extra, p + 2)[0]
                if si == b"BC" and slen == 2:
                    bsize = struct.unpack_from("<H", extra, p + 4)[0] + 1
                p += 4 + slen
            if bsize is None:
                raise OSError("{} is not BGZF (no BC field)".format(path))
            body = f.read(bsize - 12 - xlen)
            if len(body) != bsize - 12 - xlen:
                raise OSError("{}: truncated BGZF member".format(path))
            yield off, zlib.decompress(body[:-8], -15)
            off += bsize


def _bgzf_lines(path: str):
    """(virtual offset of the line's first byte, virtual offset just past its newline, line bytes) for every line."""
    pending, start_voff = b"", None
    held = None                         # a line that ended with its member: its end offset is the next member's start
    for coff, data in _bgzf_members(path):
        if held is not None:
            yield held[0], coff << 16, held[1]
            held = None
        pos = 0
        while pos < len(data):
            nl = data.find(b"\n", pos)
            if nl < 0:
                if not pending:
                    start_voff = (coff << 16) | pos
                pending += data[pos:]
                break
            if pending:
                line, beg, pending = pending + data[pos:nl], start_voff, b""
            else:
                line, beg = data[pos:nl], (coff << 16) | pos
            nxt = nl + 1
            if nxt < len(data):
                yield beg, (coff << 16) | nxt, line
            else:
                held = (beg, line)      # the next byte is the first of the following member
            pos = nxt
    eof_voff = os.path.getsize(path) << 16
    if held is not None:
        yield held[0], eof_voff, held[1]
    if pending:
        yield start_voff, eof_voff, pending


def write_tbi(vcf_gz_path: str, tbi_path: str = None) -> str:
    """Index a BGZF VCF sorted by contig and position; returns the index path."""
    tbi_path = tbi_path or vcf_gz_path + ".tbi"
    names: List[bytes] = []
    bins: List[Dict[int, List[List[int]]]] = []
    linear: List[List[int]] = []
    last_pos = -1
    for beg_v, end_v, line in _bgzf_lines(vcf_gz_path):
        if not line or line[:1] == b"#":
            continue
        f = line.split(b"\t", 8)
        if len(f) < 8:
            raise ValueError("malformed VCF record while indexing {}".format(vcf_gz_path))
        chrom, pos, ref, info = f[0], int(f[1]), f[3], f[7]
        beg0 = pos - 1
        end0 = beg0 + max(len(ref), 1)
        for kv in info.split(b";"):
            if kv.startswith(b"END="):          # like htslib's VCF preset: INFO END replaces the REF-length end when it lies past POS
                try:
                    e = int(kv[4:])
                    if e > beg0:
                        end0 = e
                except ValueError:
                    pass
                break
        if not names or names[-1] != chrom:
            if chrom in names:
                raise ValueError("{}: records of contig {} are not contiguous".format(vcf_gz_path, chrom.decode()))
            names.append(chrom)
            bins.append({})
            linear.append([])
            last_pos = -1
        if beg0 < last_pos:
            raise ValueError("{}: records are not sorted by position".format(vcf_gz_path))
        last_pos = beg0
        chunks = bins[-1].setdefault(_reg2bin(beg0, end0), [])
        if chunks and chunks[-1][1] == beg_v:
            chunks[-1][1] = end_v
        else:
            chunks.append([beg_v, end_v])
        lin = linear[-1]
        w0, w1 = beg0 >> _MIN_SHIFT, (end0 - 1) >> _MIN_SHIFT
        if len(lin) <= w1:
            lin.extend([0] * (w1 + 1 - len(lin)))
        for w in range(w0, w1 + 1):
            if lin[w] == 0:
                lin[w] = beg_v
    nm = b"".join(n + b"\x00" for n in names)
    out = bytearray(b"TBI\x01")
    # n_ref, format (2 = VCF), col_seq, col_beg, col_end, meta char, skip, l_nm
    out += struct.pack("<8i", len(names), 2, 1, 2, 0, ord("#"), 0, len(nm))
    out += nm
    for b, lin in zip(bins, linear):
        out += struct.pack("<i", len(b))
        for bin_id in sorted(b):
            out += struct.pack("<Ii", bin_id, len(b[bin_id]))
            for cb, ce in b[bin_id]:
                out += struct.pack("<QQ", cb, ce)
        # like htslib: windows before the first record take its offset, later windows without a record of their own
        # the previous window's
        filled = list(lin)
        first = next((i for i, v in enumerate(filled) if v), len(filled))
        for i in range(first):
            filled[i] = filled[first]
        for i in range(first + 1, len(filled)):
            if filled[i] == 0:
                filled[i] = filled[i - 1]
        out += struct.pack("<i", len(filled))
        out += struct.pack("<%dQ" % len(filled), *filled)
    w = BgzfWriter(tbi_path)
    w.write(bytes(out))
    w.close()
    return tbi_path
