"""
CPU tests of the native block VCF ingest (csrc/trt_ingest.cpp behind trt_vcf_*).

The checker is the pure-Python text reader ``cyvcf2_compat.TextVCF`` — the reader every golden-output CLI test
of rounds 1 was pinned with (it follows cyvcf2's conventions).  The C++ reader must hand
out the same arrays bit for bit (GT int16, Integer int32, Float float32 compared as raw bits), the same fixed
columns, the same errors at the same record, and the same re-serialised text, on the reference's own example
files (copies under tests/golden/data; the whole reference tree when it is present) and on edge cases built
here: ploidy mixes, half calls, dropped trailing fields, vector values in scalar fields, odd tokens, ragged
columns, CRLF, blank lines, unterminated last line, sample subsets, gzip / BGZF / plain, corrupt members.
"""
import glob
import gzip
import os
import struct
import zlib

import numpy as np
import pytest

from trtools_b200 import cyvcf2_compat as cc
from trtools_b200.vcf_ingest import NativeVCF, NativeVariant

from conftest import DATA

REF_ROOT = "/root/reference"


def _same(a, b):
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    if a.dtype.kind == 'f':
        return np.array_equal(a.view(np.uint32), b.view(np.uint32))
    return np.array_equal(a, b)


def _next(it):
    try:
        return next(it)
    except StopIteration:
        return None
    except Exception as e:   # the error itself is compared
        return e


def _try(fn):
    try:
        return fn()
    except Exception as e:
        return (type(e).__name__, str(e))


def _compare(path, max_records=None, block_loci=5, samples=None, threads=None, block_bytes=None):
    """Walk both readers in lock step; returns the number of records compared."""
    try:
        t = cc.TextVCF(path, samples=samples)
    except OSError:
        with pytest.raises(OSError):
            NativeVCF(path, samples=samples)
        return 0
    n = NativeVCF(path, samples=samples, threads=threads)
    n._native_block_loci = block_loci
    if block_bytes:
        n._native_block_bytes = block_bytes
    assert t.raw_header == n.raw_header
    assert t.samples == n.samples
    assert t.seqnames == n.seqnames
    keys = list(t._format_types)
    it, inn = iter(t), iter(n)
    count = 0
    while max_records is None or count < max_records:
        a, b = _next(it), _next(inn)
        if a is None or b is None:
            assert a is None and b is None, (path, count, a, b)
            break
        if isinstance(a, Exception) or isinstance(b, Exception):
            assert type(a) is type(b) and str(a) == str(b), (path, count, a, b)
            break
        where = (path, count, a.CHROM, a.POS)
        assert (a.CHROM, a.POS, a.ID, a.REF, a.ALT, a.QUAL, a.FILTER, a.FORMAT) == \
               (b.CHROM, b.POS, b.ID, b.REF, b.ALT, b.QUAL, b.FILTER, b.FORMAT), where
        assert list(a.INFO) == list(b.INFO), where
        if not t.samples:
            assert a.genotype is None and b.genotype is None, where
        else:
            ga, gb = _try(lambda: a.genotype.array()), _try(lambda: b.genotype.array())
            if isinstance(ga, tuple) or isinstance(gb, tuple):      # both readers fail alike (e.g. GT "0/x")
                assert ga == gb, where + (ga, gb)
                count += 1
                continue
            assert _same(ga, gb), where
            assert a.ploidy == b.ploidy, where
            assert a.genotypes == b.genotypes, where
        for k in keys:
            fa, fb = _try(lambda: a.format(k)), _try(lambda: b.format(k))
            if isinstance(fa, tuple) or isinstance(fb, tuple):
                assert fa == fb, where + (k, fa, fb)
                continue
            assert _same(fa, fb), where + (k, fa[:4], fb[:4])
        assert str(a) == str(b), where
        count += 1
    t.close()
    n.close()
    return count


# ---- the reference's own files ---------------------------------------------------------------

FIXTURES = sorted(glob.glob(os.path.join(DATA, "*.vcf")) + glob.glob(os.path.join(DATA, "*.vcf.gz")))


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_native_reader_matches_text_reader_on_fixtures(path):
    assert _compare(path, max_records=150, block_loci=16) > 0


def test_fixture_trio_all_records_one_thread_vs_many():
    path = os.path.join(DATA, "trio_chr21_hipstr.sorted.vcf.gz")
    outs = []
    for threads in (1, 8):
        v = NativeVCF(path, threads=threads)
        v._prefetch = ("DP", "Q", "DSTUTTER", "DFLANKINDEL")
        gts, dps, qs = [], [], []
        for r in v:
            gts.append(r.genotype.array())
            dps.append(r.format("DP"))
            qs.append(r.format("Q"))
        outs.append(gts + dps + qs)
    assert len(outs[0]) == 3 * 9532          # BASELINE config 1
    for a, b in zip(*outs):
        assert _same(a, b)


@pytest.mark.needs_reference
def test_native_reader_matches_text_reader_on_reference_tree():
    files = sorted(glob.glob(REF_ROOT + "/example-files/*.vcf*") +
                   glob.glob(REF_ROOT + "/trtools/testsupport/sample_vcfs/**/*.vcf*", recursive=True))
    files = [f for f in files if not f.endswith((".tbi", ".csi"))]
    assert len(files) > 100
    total = sum(_compare(f, max_records=12, block_loci=5) for f in files)
    assert total > 1000


# ---- edge cases ------------------------------------------------------------------------------------

HEADER = (
    "##fileformat=VCFv4.1\n"
    "##command=HipSTR-test\n"
    "##contig=<ID=1,length=1000000>\n"
    '##INFO=<ID=START,Number=1,Type=Integer,Description="s">\n'
    '##INFO=<ID=END,Number=1,Type=Integer,Description="e">\n'
    '##INFO=<ID=PERIOD,Number=1,Type=Integer,Description="p">\n'
    '##INFO=<ID=AF,Number=A,Type=Float,Description="vector info">\n'
    '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n'
    '##FORMAT=<ID=Q,Number=1,Type=Float,Description="q">\n'
    '##FORMAT=<ID=DP,Number=1,Type=Integer,Description="dp">\n'
    '##FORMAT=<ID=DFLANKINDEL,Number=1,Type=Integer,Description="dfi">\n'
    '##FORMAT=<ID=GB,Number=1,Type=String,Description="bp diffs">\n'
    '##FORMAT=<ID=PL,Number=G,Type=Integer,Description="vector">\n'
    '##FORMAT=<ID=QEXP,Number=3,Type=Float,Description="fixed-size vector">\n'
)


def _vcf_text(records, samples=("A", "B", "C", "D"), eol="\n", header=HEADER):
    h = header + "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(samples) + "\n"
    if not samples:
        h = header + "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n"
    return h.replace("\n", eol) + "".join(r + eol for r in records)


def _rec(pos, fmt, cols, ref="ACACAC", alt="ACAC,ACACACAC", info="START=%d;END=%d;PERIOD=2"):
    fixed = ["1", str(pos), "STR_%d" % pos, ref, alt, ".", ".", info % (pos, pos + len(ref) - 1) if "%d" in info else info]
    if fmt is None:
        return "\t".join(fixed)
    return "\t".join(fixed + [fmt] + list(cols))


EDGE_RECORDS = [
    _rec(100, "GT:Q:DP:DFLANKINDEL:GB", ["0|1:0.99:30:1:0|-2", "1/2:0.5:12:0:-2/2", ".:.:.:.:.", "./1:1:7:2:.|-2"]),
    _rec(200, "GT:DP", ["0", "1", ".", "2"]),                                  # haploid record
    _rec(300, "GT:DP", ["0/1/2", "1|1|0", "0/1", "."]),                        # triploid + diploid + missing
    _rec(400, "GT:Q:DP", ["0/1", "1/1:0.25", "0/0:0.5:3", "2|0:.:4"]),         # trailing fields dropped
    _rec(500, "GT:Q:DP", ["0/1:1e-3:5", "0/0:nan:6", "1/1:inf:-7", "0|2:-0.0:+8"]),
    _rec(600, "GT:Q:DP", ["0/1:0.1234567890123456789:5", "0/0:.5:6", "1/1:5.:7", "0|2:123456789012345678901234567890:8"]),
    _rec(700, "GT:DP:PL:QEXP", ["0/1:5:0,10,20:0.1,0.2,0.7", "0/0:6:.:.", "1/1:7:1,2,3:.,.,.", "0|2:8:5,6:1,0,0"]),
    _rec(800, "GT:DP", ["0/1:1,2", "0/0:6", "1/1:7", "0|2:8"]),                # vector value in a scalar field
    _rec(900, "GT:DP", ["0/x:1", "0/0:6", "1/1:7", "0|2:8"]),                  # odd GT token: Python int() decides
    _rec(1000, "DP:Q", ["5:0.5", "6:0.25", "7:.", ".:."]),                     # no GT key
    _rec(1100, "DP:GT:Q", ["5:0|1:0.5", "6:1/1:0.25", "7:.:.", ".:./.:1"]),    # GT not first
    _rec(1200, "GT:DP", ["0/1:3", "0/0:6", "1/1:7"]),                          # a sample column short
    _rec(1300, "GT:DP:DP", ["0/1:3:4", "0/0:6:7", "1/1:7:8", "0|2:8:9"]),      # duplicated key: first wins
    _rec(1400, "GT:Q", ["0/1:1E+2", "0/0:-1.5e-10", "1/1:1e", "0|2:0x10"]),    # exponents; tokens float() rejects
    _rec(1500, "GT:DP", ["40000/1:3", "0/0:6", "1/1:7", "0|2:8"]),             # allele index beyond int16
    _rec(1600, "GT:DP", ["0/1:99999999999", "0/0:2147483647", "1/1:-2147483648", "0|2:-2147483647"]),
    _rec(1700, "GT", ["0|0", "1|1", "2|2", "0|1"], info="START=1700;END=1705;PERIOD=2;AF=0.5,0.25;IMP"),
    _rec(1800, "GT:DP", ["|1:3", "0/:6", "/:7", "0|2:"]),                      # empty allele / empty value tokens
    _rec(1900, "GT:DP", ["0/1:3", "0/0:6", "1/1:7", "0|2:8", "0/0:1"]),        # a sample column too many
]


def _write(tmp_path, name, text, mode="plain"):
    p = str(tmp_path / name)
    data = text.encode()
    if mode == "plain":
        open(p, "wb").write(data)
    elif mode == "gzip":
        with gzip.open(p, "wb") as f:
            f.write(data)
    else:
        open(p, "wb").write(_bgzf(data, member=int(mode)))
    return p


def _bgzf(data: bytes, member: int = 300) -> bytes:
    """BGZF (SAM spec §4.1): gzip members with a 'BC' extra field carrying the member size, then the EOF member."""
    out = []
    for o in list(range(0, len(data), member)) + [None]:
        chunk = b"" if o is None else data[o:o + member]
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = co.compress(chunk) + co.flush()
        bsize = 12 + 6 + len(body) + 8 - 1
        out.append(struct.pack("<4BIBBH2BHH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6, ord('B'), ord('C'), 2, bsize))
        out.append(body)
        out.append(struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk)))
    return b"".join(out)


@pytest.mark.parametrize("mode", ["plain", "gzip", "300", "65280"])
@pytest.mark.parametrize("block_loci", [1, 3, 512])
def test_edge_records_match_text_reader(tmp_path, mode, block_loci):
    p = _write(tmp_path, "edge.vcf" + ("" if mode == "plain" else ".gz"), _vcf_text(EDGE_RECORDS), mode)
    assert _compare(p, block_loci=block_loci) == len(EDGE_RECORDS)


def test_bgzf_written_here_is_valid_gzip(tmp_path):
    p = _write(tmp_path, "x.vcf.gz", _vcf_text(EDGE_RECORDS), "300")
    assert gzip.open(p, "rb").read().decode() == _vcf_text(EDGE_RECORDS)


def test_flagged_records_are_reparsed_not_guessed(tmp_path):
    p = _write(tmp_path, "edge.vcf", _vcf_text(EDGE_RECORDS))
    v = NativeVCF(p)
    recs = list(v)
    blk = recs[0]._nblk
    blk.parse(("DP", "Q"))
    by_pos = {r.POS: i for i, r in enumerate(recs)}
    assert blk.status[by_pos[100]] == 0 and blk.status[by_pos[300]] == 0
    for pos in (900, 1000, 1200, 1500, 1900):       # odd token, no GT, ragged columns, int16 overflow
        assert blk.status[by_pos[pos]] == 2, pos
    assert blk.present["DP"][by_pos[800]] == 2       # "1,2" in a scalar field
    assert blk.present["DP"][by_pos[1600]] == 2      # beyond int32
    assert blk.present["Q"][by_pos[1400]] == 2       # "1e", "0x10"
    assert blk.present["Q"][by_pos[200]] == 0        # key absent from FORMAT
    assert blk.gt.shape == (len(recs), 4, 4)         # sized by the triploid record
    assert list(blk.rec_ploidy[:3]) == [2, 1, 3]


def test_crlf_blank_lines_and_unterminated_last_record(tmp_path):
    text = _vcf_text(EDGE_RECORDS[:4], eol="\r\n")
    p = _write(tmp_path, "crlf.vcf", text)
    assert _compare(p) == 4
    text = _vcf_text(EDGE_RECORDS[:2]) + "\n\n" + EDGE_RECORDS[2] + "\n\n" + EDGE_RECORDS[3]    # no final newline
    p = _write(tmp_path, "blank.vcf", text)
    assert _compare(p, block_loci=2) == 4


def test_no_samples_and_header_only(tmp_path):
    p = _write(tmp_path, "nosamp.vcf", _vcf_text([_rec(100, None, []), _rec(200, None, [])], samples=()))
    assert _compare(p) == 2
    p = _write(tmp_path, "empty.vcf", _vcf_text([]))
    assert _compare(p) == 0
    p = _write(tmp_path, "notvcf.txt", "hello\nworld\n")
    with pytest.raises(OSError):
        NativeVCF(p)
    with pytest.raises(OSError):
        NativeVCF(str(tmp_path / "does_not_exist.vcf"))


def test_malformed_record_raises_when_reached(tmp_path):
    text = _vcf_text(EDGE_RECORDS[:3] + ["1\t5000\tonly_three_columns"] + EDGE_RECORDS[3:5])
    p = _write(tmp_path, "broken.vcf", text)
    assert _compare(p, block_loci=512) == 3         # both readers raise ValueError at the 4th record
    v = NativeVCF(p)
    got = [next(v) for _ in range(3)]
    assert [r.POS for r in got] == [100, 200, 300]
    with pytest.raises(ValueError):
        next(v)


@pytest.mark.parametrize("keep", [["B"], ["A", "D"], ["D", "A", "zzz"], []])
def test_sample_subsets(tmp_path, keep):
    p = _write(tmp_path, "edge.vcf.gz", _vcf_text(EDGE_RECORDS), "300")
    # record 1200 lacks its 4th sample column: picking "D" out of it fails (IndexError) in both readers
    want = 11 if "D" in keep else len(EDGE_RECORDS)
    assert _compare(p, samples=keep, block_loci=4) == want


def test_region_query(tmp_path):
    p = _write(tmp_path, "edge.vcf", _vcf_text(EDGE_RECORDS))
    a = [r.POS for r in cc.TextVCF(p)("1:300-805")]
    b = [r.POS for r in NativeVCF(p)("1:300-805")]
    assert a == b and len(a) > 3


def test_corrupt_and_truncated_bgzf_fail_loudly(tmp_path):
    data = bytearray(_bgzf(_vcf_text(EDGE_RECORDS * 20).encode(), 4000))
    bad = bytearray(data)
    bad[len(bad) // 2] ^= 0x55
    p = str(tmp_path / "bad.vcf.gz")
    open(p, "wb").write(bytes(bad))
    with pytest.raises(OSError):
        list(NativeVCF(p))
    p = str(tmp_path / "trunc.vcf.gz")
    open(p, "wb").write(bytes(data[:len(data) // 2]))
    with pytest.raises(OSError):
        list(NativeVCF(p))


def test_float_tokens_round_like_numpy(tmp_path):
    """Float FORMAT values: decimal string -> nearest double -> float32, bit for bit as np.float32(str)."""
    rng = np.random.default_rng(20261017)
    toks = []
    for i in range(6000):
        kind = i % 6
        if kind == 0:
            toks.append("%.*f" % (int(rng.integers(0, 12)), rng.random()))
        elif kind == 1:
            toks.append(repr(float(np.float32(rng.normal() * 10 ** int(rng.integers(-8, 8))))))
        elif kind == 2:
            toks.append("%.17g" % (rng.random() * 10 ** int(rng.integers(-30, 30))))
        elif kind == 3:
            toks.append("%d" % rng.integers(-10 ** 9, 10 ** 9))
        elif kind == 4:
            # halfway cases between adjacent float32 values (double rounding matters here)
            f = np.float32(rng.random())
            mid = (float(f) + float(np.nextafter(f, np.float32(2)))) / 2
            toks.append("%.25f" % mid)
        else:
            toks.append("%.*e" % (int(rng.integers(0, 20)), rng.normal() * 10 ** int(rng.integers(-40, 40))))
    samples = ["S%d" % i for i in range(len(toks))]
    rec = _rec(100, "GT:Q", ["0/1:" + t for t in toks])
    p = _write(tmp_path, "floats.vcf", _vcf_text([rec], samples=samples))
    v = NativeVCF(p)
    r = next(v)
    got = r.format("Q")[:, 0]
    assert r._nblk.present["Q"][0] == 1              # all of them parsed natively
    want = np.array([np.float32(t) for t in toks], dtype=np.float32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_wide_records_many_threads(tmp_path):
    """A block of wide records (S = 3000) in several byte-capped runs, threads 1 vs 8, vs the text reader."""
    rng = np.random.default_rng(7)
    S, L = 3000, 24
    samples = ["S%d" % i for i in range(S)]
    recs = []
    for l in range(L):
        a = rng.integers(-1, 3, size=(S, 2))
        dp = rng.integers(0, 80, size=S)
        q = rng.random(S)
        cols = []
        for s in range(S):
            if rng.random() < 0.02:
                cols.append(".")
                continue
            g = "|".join("." if x < 0 else str(x) for x in a[s])
            cols.append("%s:%d:%.6g" % (g, dp[s], q[s]))
        recs.append(_rec(100 + 10 * l, "GT:DP:Q", cols))
    p = _write(tmp_path, "wide.vcf.gz", _vcf_text(recs, samples=samples), "65280")
    assert _compare(p, block_loci=7, block_bytes=200000, threads=8) == L
    stacks = []
    for threads in (1, 8):
        v = NativeVCF(p, threads=threads)
        v._prefetch = ("DP", "Q")
        rr = list(v)
        assert all(isinstance(r, NativeVariant) for r in rr)
        stacks.append((np.stack([r.genotype.array() for r in rr]), np.stack([r.format("DP") for r in rr]),
                       np.stack([r.format("Q") for r in rr])))
    for a, b in zip(*stacks):
        assert _same(a, b)


# ---- the block hand-off: build_block takes the reader's stacked arrays as they are ----------------------


class _RecordingCtx:
    """Stands in for _lib.Context: records what a Block uploads (no GPU in the CPU suite)."""

    def __init__(self):
        self.calls = {}

    def block_begin(self, L, S, P, vcftype):
        self.calls["begin"] = (L, S, P, vcftype)

    def block_set_gt(self, gt):
        self.calls["gt"] = np.array(gt)
        self.calls["gt_is_view"] = gt.base is not None

    def block_set_gt_packed(self, gt2, phase=None):
        from trtools_b200.block import unpack_gt
        self.calls["gt"] = unpack_gt(np.array(gt2), None if phase is None else np.array(phase))
        self.calls["gt_is_view"] = gt2.base is not None
        self.calls["packed"] = True

    def block_set_gt_nibble(self, g4, phase=None):
        from trtools_b200.block import unpack_gt4
        self.calls["gt"] = unpack_gt4(np.array(g4), None if phase is None else np.array(phase))
        self.calls["gt_is_view"] = g4.base is not None
        self.calls["packed"] = True

    def block_set_alleles(self, *a):
        self.calls["alleles"] = a

    def block_set_format(self, slot, arr):
        self.calls[("fmt", slot)] = np.array(arr).reshape(self.calls["begin"][0], self.calls["begin"][1], -1)

    def harmonize(self):
        return {"flags": np.zeros(self.calls["begin"][0], dtype=np.int32)}


@pytest.mark.parametrize("path,vcftype,keys", [
    (os.path.join(DATA, "trio_chr21_hipstr.sorted.vcf.gz"), "hipstr", ("DP", "Q", "DFLANKINDEL", "DSTUTTER")),
    (os.path.join(DATA, "many_samples.vcf.gz"), "hipstr", ("DP", "Q")),
    (os.path.join(DATA, "test_gangstr_head.vcf"), "gangstr", ("DP", "Q", "QEXP")),
])
def test_build_block_uploads_identical_arrays_from_both_readers(path, vcftype, keys):
    from trtools_b200 import block as _block
    uploads = []
    for cls in (cc.TextVCF, NativeVCF):
        v = cls(path)
        if cls is NativeVCF:
            v._prefetch = keys
            v._native_block_loci = 64
        recs = [r for _, r in zip(range(64), v)]
        ctx = _RecordingCtx()
        blk = _block.build_block(ctx, vcftype, recs, keys)
        uploads.append((ctx.calls, blk))
    (a, blk_a), (b, blk_b) = uploads
    assert b["gt_is_view"]                                   # the reader's own slab, no per-record copy
    assert a["begin"] == b["begin"]
    assert _same(a["gt"], b["gt"])
    fa = sorted(k for k in a if isinstance(k, tuple))
    assert fa == sorted(k for k in b if isinstance(k, tuple)) and len(fa) == len(keys)
    for k in fa:
        assert _same(a[k], b[k]), k
    assert blk_a.seqs == blk_b.seqs and np.array_equal(blk_a.allele_off, blk_b.allele_off)


def test_format_arrays_are_private_copies():
    """dumpSTR nulls filtered calls in the array format() returned; the block's stacked arrays must not change."""
    v = NativeVCF(os.path.join(DATA, "many_samples.vcf.gz"))
    v._prefetch = ("DP",)
    r = next(v)
    r.genotype.array()                       # first access parses the run (GT + the prefetch keys)
    before = r._nblk.fmt["DP"].copy()
    a = r.format("DP")
    a[:] = -7
    g = r.genotype.array()
    g[:] = 9
    assert np.array_equal(r._nblk.fmt["DP"], before)
    assert not (r._nblk.gt[0] == 9).any()


@pytest.mark.parametrize("fname,limit", [("trio_chr21_hipstr.sorted.vcf.gz", 1500), ("many_samples.vcf.gz", 400)])
def test_dumpstr_write_back_text_identical_for_both_readers(fname, limit):
    """dumpSTR's record write-back (FORMAT:FILTER, nulling of filtered calls; reference dumpSTR.py:684-746) on
    blocks cut like the CLI cuts them (runs of equal ploidy): the re-serialised record text must not depend on
    the reader — including which untouched fields the writer passes through verbatim."""
    import itertools
    import types
    from trtools_b200 import block as _block, dumpSTR, _lib

    class MinDP:
        name, gpu_kind, field, threshold = "HipSTRCallMinDepth10", _lib.CF_MIN, "DP", 10

    keys = ["DP", "Q", "DFLANKINDEL", "DSTUTTER"]
    outs = []
    for cls in (cc.TextVCF, NativeVCF):
        v = cls(os.path.join(DATA, fname))
        if cls is NativeVCF:
            v._prefetch = tuple(keys) + ("LC",)
            v._native_block_loci = 300
        it, texts, done = iter(v), [], False
        rng = np.random.default_rng(1)
        while not done and len(texts) < limit:
            recs = []
            while len(recs) < 300:
                try:
                    rec = next(it)
                except StopIteration:
                    done = True
                    break
                if recs and rec.ploidy != recs[0].ploidy:
                    it = itertools.chain([rec], it)
                    break
                recs.append(rec)
            if not recs:
                break
            blk = _block.build_block(_RecordingCtx(), "hipstr", recs, keys)
            res = types.SimpleNamespace(call_mask=(rng.random((len(recs), blk.S)) < 0.3).astype(np.uint32),
                                        host_values={})
            for l, r in enumerate(recs):
                dumpSTR._apply_to_record(r, blk, res, l, [MinDP()])
                texts.append(str(r))
        outs.append(texts)
    assert len(outs[0]) >= limit and outs[0] == outs[1]


def _fuzz_token(rng, kind):
    """A FORMAT value: mostly well-formed, sometimes the odd spellings real files (and broken ones) contain."""
    r = rng.random()
    if kind == "int":
        if r < 0.70:
            return str(int(rng.integers(-50, 5000)))
        return rng.choice([".", "", "+7", "-0", "007", "1,2", "3.5", "1e3", "2147483647", "2147483648", "-2147483648",
                           "-2147483649", "99999999999", "x", "1_0", " 1", "--1", "+", "-"])
    if kind == "float":
        if r < 0.70:
            return rng.choice(["%.3f", "%.6g", "%.10e", "%.17g"]) % (rng.normal() * 10 ** int(rng.integers(-6, 6)))
        return rng.choice([".", "", "nan", "NaN", "-nan", "inf", "-inf", "+Infinity", "1e", "e5", ".e5", "1.2.3", "0x1p3",
                           "1,2", "5.", ".5", "-.5", "+.", "1e400", "-1e-400", "00012.500", "1_0.5", "１"])
    if kind == "gt":
        if r < 0.75:
            a, b = int(rng.integers(0, 12)), int(rng.integers(0, 12))
            return "%d%s%d" % (a, rng.choice(["|", "/"]), b)
        return rng.choice([".", "./.", ".|.", "./1", "1|.", "0", "3", "0/1/2", "1|2|3|4", "", "/", "|", "0/", "|1", "12/345",
                           "32767/0", "32768/0", "0/x", "a|b", "0\\1", "1/ 2"])
    return rng.choice([".", "abc", "1|2;3|4", "-2|4", "x,y", ""])


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_random_records_match_text_reader(tmp_path, seed):
    """Seeded fuzz: random FORMAT layouts (key order, duplicates, unknown keys, GT anywhere or absent), random
    tokens including malformed ones, dropped trailing fields, ragged columns.  Whatever the text reader returns or
    raises, the C++ reader must return or raise the same (by parsing it, or by flagging it for the Python parser)."""
    rng = np.random.default_rng(seed)
    samples = ["S%d" % i for i in range(int(rng.integers(1, 9)))]
    kinds = {"GT": "gt", "DP": "int", "DFLANKINDEL": "int", "Q": "float", "GB": "str", "PL": "int", "QEXP": "float",
             "ZZ": "str"}
    recs = []
    for i in range(160):
        keys = list(rng.permutation(list(kinds))[:int(rng.integers(1, 7))])
        if rng.random() < 0.85 and "GT" in keys:
            keys.remove("GT")
            keys.insert(0, "GT")
        if rng.random() < 0.05:
            keys.append(keys[0])
        cols = []
        ncol = len(samples) if rng.random() < 0.95 else int(rng.integers(0, len(samples) + 3))
        for s in range(ncol):
            toks = [_fuzz_token(rng, kinds[k]) for k in keys]
            if rng.random() < 0.15:
                toks = toks[:int(rng.integers(1, len(toks) + 1))]
            if rng.random() < 0.03:
                toks.append("extra")
            cols.append(":".join(toks))
        recs.append(_rec(100 + 10 * i, ":".join(keys), cols))
    mode = ["plain", "gzip", "300", "65280"][seed % 4]
    p = _write(tmp_path, "fuzz.vcf" + ("" if mode == "plain" else ".gz"), _vcf_text(recs, samples=samples), mode)
    assert _compare(p, block_loci=int(rng.integers(1, 40))) == len(recs)


@pytest.mark.parametrize("fname,regions", [
    ("trio_chr21_hipstr.sorted.vcf.gz", ["chr21:9500000-9700000", "chr21:9489666-9489666", "chr21:1-9000000", "chr21",
                                         "chr21:47000000-48000000", "chr21:48100000-48200000", "chr21:34000000-",
                                         "chr21:20000000-20050000", "chr20:1-100", "chr21:9546720-9546780"]),
    ("many_samples.vcf.gz", ["1:3000000-3200000", "1:1-1000", "1", "2:1-5", "1:3045000-3045500", "1:3060000-"]),
])
def test_region_queries_through_tabix_index_match_linear_scan(data_dir, fname, regions):
    """vcf(region) with a .tbi (the reference's own index files): seek + early stop must serve exactly the records
    the text reader's linear scan serves."""
    from trtools_b200 import vcf_ingest
    path = os.path.join(data_dir, fname)
    assert os.path.isfile(path + ".tbi")
    used_index = 0
    for region in regions:
        want = [(r.CHROM, r.POS, r.ID) for r in cc.TextVCF(path)(region)]
        v = NativeVCF(path)
        got_recs = list(v(region))
        got = [(r.CHROM, r.POS, r.ID) for r in got_recs]
        assert got == want, (region, len(got), len(want))
        used_index += bool(v._region_stop or v._region_empty)
        if got_recs:                               # arrays of a record reached through a seek
            t = next(iter(cc.TextVCF(path)(region)))
            assert _same(got_recs[0].genotype.array(), t.genotype.array()), region
            assert _same(got_recs[0].format("DP"), t.format("DP")), region
    assert used_index == len(regions)
    # ONE reader serving every region in turn (and once more in reverse order): cyvcf2 readers can be queried repeatedly
    v = NativeVCF(path)
    for region in list(regions) + list(regions)[::-1]:
        want = [(r.CHROM, r.POS, r.ID) for r in cc.TextVCF(path)(region)]
        got = [(r.CHROM, r.POS, r.ID) for r in v(region)]
        assert got == want, ("repeated query", region, len(got), len(want))
    # without the index: same answers by linear scan
    assert vcf_ingest._tabix_start(path + ".nope.tbi", "1", 5) is None


def test_refill_boundaries_small_chunks(tmp_path):
    """Members, records and pread pieces straddling refill boundaries: the reader in a subprocess with 128 KiB fills
    (TRTOOLS_B200_INGEST_CHUNK) over ~1.5 MB files, plain / gzip / BGZF, against the default 16 MiB fills."""
    import subprocess
    import sys
    rng = np.random.default_rng(3)
    S, L = 1500, 60
    samples = ["S%d" % i for i in range(S)]
    recs = []
    for l in range(L):
        a = rng.integers(0, 9, size=(S, 2))
        dp = rng.integers(0, 99, size=S)
        q = rng.random(S)
        cols = ["%d|%d:%d:%.5g" % (a[s, 0], a[s, 1], dp[s], q[s]) for s in range(S)]
        recs.append(_rec(100 + 10 * l, "GT:DP:Q", cols))
    text = _vcf_text(recs, samples=samples)
    script = (
        "import sys, hashlib, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "from trtools_b200.vcf_ingest import NativeVCF\n"
        "v = NativeVCF(sys.argv[1], threads=4); v._prefetch = ('DP', 'Q'); v._native_block_loci = 7\n"
        "h = hashlib.sha256()\n"
        "n = 0\n"
        "for r in v:\n"
        "    h.update(r.genotype.array().tobytes()); h.update(r.format('DP').tobytes()); h.update(r.format('Q').tobytes())\n"
        "    h.update(str(r.POS).encode()); n += 1\n"
        "print(n, h.hexdigest())\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    outs = set()
    for mode in ("plain", "gzip", "65280", "3000"):
        p = _write(tmp_path, "chunks_%s.vcf%s" % (mode, "" if mode == "plain" else ".gz"), text, mode)
        for chunk in ("131072", ""):
            env = dict(os.environ, TRTOOLS_B200_INGEST_CHUNK=chunk)
            res = subprocess.run([sys.executable, "-c", script, p], env=env, capture_output=True, text=True, timeout=300)
            assert res.returncode == 0, res.stderr[-2000:]
            outs.add(res.stdout.strip())
    assert len(outs) == 1 and next(iter(outs)).startswith("%d " % L), outs


def test_vectorised_sample_text_equals_the_loop(tmp_path, monkeypatch):
    """Record serialisation (the dumpSTR writer's path): the np.char version must produce the per-sample loop's text
    byte for byte — untouched records, records whose numeric / string / vector fields were decoded, and records
    after dumpSTR-style write-back (FILTER strings set, filtered calls nulled)."""
    monkeypatch.setattr(cc, "_VEC_MIN_SAMPLES", 1)
    rng = np.random.default_rng(5)
    paths = list(FIXTURES) + [_write(tmp_path, "edge.vcf", _vcf_text(EDGE_RECORDS))]
    checked = vectorised = native = 0
    for path in paths:
        for cls in (cc.TextVCF, NativeVCF):
            v = cls(path)
            it = iter(v)
            for _ in range(60):
                r = _next(it)
                if r is None or isinstance(r, Exception):
                    break
                S = len(v.samples)
                if S == 0 or isinstance(_try(lambda: r.genotype.array()), tuple) or len(r._sample_cols) != S:
                    continue             # no samples / GT the readers reject / ragged record
                for stage in range(3):
                    if stage == 1:       # decode every field (numeric ones get re-serialised from their arrays)
                        for k in list(r.FORMAT):
                            _try(lambda: r.format(k))
                    if stage == 2:       # dumpSTR write-back
                        filtered = rng.random(S) < 0.4
                        r.set_format("FILTER", np.char.encode(np.where(filtered, "HipSTRCallMinDepth10_3", "PASS")))
                        gts = r.genotypes
                        for i in np.nonzero(filtered)[0]:
                            gts[i] = [-1] * r.ploidy + [False]
                        r.genotypes = gts
                        for k in list(r.FORMAT):
                            if k in ("GT", "FILTER"):
                                continue
                            vals = _try(lambda: r.format(k))
                            if isinstance(vals, tuple):
                                continue
                            if vals.dtype.kind == "U":
                                vals[filtered] = "."
                                vals = np.char.encode(vals)
                            elif vals.dtype.kind == "f":
                                vals[filtered] = np.nan
                            else:
                                vals[filtered] = cc.INT32_MISSING
                            r.set_format(k, vals)
                    want = _try(r._sample_text_loop)
                    got = _try(r._sample_text)
                    assert got == want, (path, cls.__name__, r.POS, stage)
                    checked += 1
                    nat = _try(r._sample_text_native)          # the C++ serialiser (trt_vcf_join_samples)
                    if isinstance(want, list) and nat is not None:
                        assert nat == "\t".join(want), (path, cls.__name__, r.POS, stage, "native")
                        native += 1
                    # how often the np.char path itself (not its fallback to the loop) produced the text
                    monkeypatch.setattr(cc.Variant, "_sample_text_loop", lambda self: None)
                    vectorised += _try(r._sample_text) is not None
                    monkeypatch.undo()
                    monkeypatch.setattr(cc, "_VEC_MIN_SAMPLES", 1)
    assert checked > 1500 and vectorised > 0.9 * checked and native > 0.9 * checked, (checked, vectorised, native)


def test_bench_ingest_leg_small():
    """bench.py's `ingest` key: the bounded VCF-text sample must come back equal to the generator's arrays."""
    import bench
    out = bench.ingest_leg(700, n_loci=6)
    assert "error" not in out, out
    assert out["arrays_equal_generator"] is True and out["value"] > 0 and out["unit"] == "loci/s"


def test_many_short_records_small_blocks_is_linear(tmp_path):
    """Few samples, many loci (trio-shaped files): handing out 512-record blocks of a large fill must not copy the
    whole unread tail per block.  150 000 three-sample records through 64 threads and 128-locus blocks finish in
    seconds (this was quadratic: ~1 MB/s on hosts with many threads), and every record still arrives, in order."""
    import time
    n = 150000
    lines = [HEADER + "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tA\tB\tC\n"]
    body = "".join("1\t%d\t.\tACACAC\tACAC\t.\t.\tSTART=%d;END=%d;PERIOD=2\tGT:DP\t0|1:%d\t1|1:%d\t.:.\n" % (
        100 + 10 * i, 100 + 10 * i, 105 + 10 * i, i % 60, (i * 7) % 60) for i in range(n))
    path = tmp_path / "short.vcf"
    path.write_text(lines[0] + body)
    from trtools_b200.vcf_ingest import NativeVCF
    t0 = time.time()
    v = NativeVCF(str(path), threads=64)
    v._native_block_loci = 128
    v._prefetch = ("DP",)
    count, last_pos, dp_sum = 0, 0, 0
    for rec in v:
        assert rec.POS > last_pos
        last_pos = rec.POS
        if count % 997 == 0:
            dp_sum += int(rec.format("DP")[0, 0])
        count += 1
    v.close()
    dt = time.time() - t0
    assert count == n
    assert dt < 60, "reading {} short records took {:.1f} s".format(n, dt)


def test_packed_parse_equals_packing_the_plain_parse(tmp_path):
    """trt_vcf_block_parse_packed writes the transfer form of trt_block_set_gt_packed straight from the text: it must
    equal numpy's packing of the plain int16 parse (phase bits included), and flag records that do not fit."""
    from trtools_b200.block import pack_gt, pack_gt4
    for path in (os.path.join(DATA, "many_samples.vcf.gz"), os.path.join(DATA, "trio_chr21_hipstr.sorted.vcf.gz")):
        outs = []
        for packed, nibble in ((True, True), (True, False), (False, False)):
            v = NativeVCF(path)
            v._packed_gt = packed
            v._nibble_gt = nibble
            v._native_block_loci = 200
            recs = [r for _, r in zip(range(200), v)]
            nblk = recs[0]._nblk
            nblk.parse(())
            outs.append(nblk)
        n4, p, q = outs
        assert p.gt2 is not None and p.gt2.ndim == 3 and q.gt2 is None and q.gt is not None
        ok = q.status == 0
        assert np.array_equal(p.status == 0, ok)
        g2, ph = pack_gt(q.gt[ok])
        assert np.array_equal(p.gt2[ok], g2) and np.array_equal(p.phase[ok], ph)
        assert np.array_equal(p.gt[ok], q.gt[ok])                 # materialised on demand
        for i in np.nonzero(ok)[0][:20]:
            assert np.array_equal(p.gt_of(int(i)), q.gt_of(int(i)))
        # the nibble form (one byte per call) when no allele index of the run exceeds 13, else the two-byte form
        fits = pack_gt4(q.gt[ok])
        if fits is not None and (q.rec_ploidy[ok] <= 2).all():
            assert n4.gt2 is not None and n4.gt2.ndim == 2
            assert np.array_equal(n4.gt2[ok], fits[0]) and np.array_equal(n4.phase[ok], fits[1])
            assert np.array_equal(n4.gt[ok], q.gt[ok])
            for i in np.nonzero(ok)[0][:20]:
                assert np.array_equal(n4.gt_of(int(i)), q.gt_of(int(i)))
        else:
            assert n4.gt2 is not None and n4.gt2.ndim == 3 and np.array_equal(n4.gt2[ok], g2)
    # an allele index above 252 falls back to the plain form for the whole run
    alts = ",".join("AC" * (k + 2) for k in range(260))
    text = _vcf_text(["1\t100\t.\tAC\t%s\t.\t.\tSTART=100;END=101;PERIOD=2\tGT\t0|1\t255|3\t.\t259/0" % alts])
    path = _write(tmp_path, "wide.vcf", text)
    v = NativeVCF(path)
    rec = next(iter(v))
    assert rec._nblk.gt2 is None or not rec._nblk.parsed
    assert rec.genotype.array().tolist() == [[0, 1, 1], [255, 3, 1], [-1, -2, 0], [259, 0, 0]]
    assert rec._nblk.gt2 is None
    # an allele index of 14 .. 252: the nibble form does not fit, the two-byte form does
    alts = ",".join("AC" * (k + 2) for k in range(20))
    text = _vcf_text(["1\t100\t.\tAC\t%s\t.\t.\tSTART=100;END=101;PERIOD=2\tGT\t0|1\t14|3\t.\t13/0" % alts])
    path = _write(tmp_path, "mid.vcf", text)
    rec = next(iter(NativeVCF(path)))
    assert rec.genotype.array().tolist() == [[0, 1, 1], [14, 3, 1], [-1, -2, 0], [13, 0, 0]]
    assert rec._nblk.gt2 is not None and rec._nblk.gt2.ndim == 3


def test_writer_emits_bgzf_that_both_readers_and_gzip_accept(tmp_path):
    """Writer('x.vcf.gz') writes BGZF (dumpSTR --zip must stay indexable): every member carries the BC size field and
    inflates to <= 65 280 bytes, the file ends with the EOF member, gzip reads it as a multi-member stream, and the C++
    block reader (which rejects plain gzip for seeks) reads the records back byte for byte."""
    import gzip
    import struct
    from trtools_b200.cyvcf2_compat import TextVCF, Writer
    src = os.path.join(DATA, "many_samples.vcf.gz")
    tmpl = TextVCF(src)
    recs = [r for _, r in zip(range(120), tmpl)]
    out = str(tmp_path / "out.vcf.gz")
    w = Writer(out, tmpl)
    for r in recs:
        w.write_record(r)
    w.close()
    raw = open(out, "rb").read()
    off, members, total = 0, 0, 0
    while off < len(raw):
        assert raw[off:off + 4] == b"\x1f\x8b\x08\x04" and raw[off + 12:off + 16] == b"BC\x02\x00"
        bsize = struct.unpack("<H", raw[off + 16:off + 18])[0] + 1
        isize = struct.unpack("<I", raw[off + bsize - 4:off + bsize])[0]
        assert isize <= 65280
        total += isize
        off += bsize
        members += 1
    assert off == len(raw) and members >= 3
    assert raw[-28:] == bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")   # the BGZF EOF marker
    text = gzip.open(out, "rt").read()
    assert len(text.encode()) == total
    want = tmpl.raw_header + "".join(str(r) for r in recs)
    assert text == want
    back = [str(r) for r in NativeVCF(out)]
    assert back == [str(r) for r in recs]


def _parse_tbi(path):
    data = gzip.open(path, "rb").read()
    assert data[:4] == b"TBI\x01"
    n_ref, fmt, cs, cb, ce, meta, skip, l_nm = struct.unpack_from("<8i", data, 4)
    names = data[36:36 + l_nm].split(b"\x00")[:n_ref]
    o, refs = 36 + l_nm, []
    for _ in range(n_ref):
        (n_bin,) = struct.unpack_from("<i", data, o)
        o += 4
        bins = {}
        for _ in range(n_bin):
            b, nc = struct.unpack_from("<Ii", data, o)
            o += 8
            bins[b] = [struct.unpack_from("<QQ", data, o + 16 * i) for i in range(nc)]
            o += 16 * nc
        (n_intv,) = struct.unpack_from("<i", data, o)
        o += 4
        refs.append((bins, list(struct.unpack_from("<%dQ" % n_intv, data, o))))
        o += 8 * n_intv
    return (fmt, cs, cb, ce, meta, skip), names, refs


def test_native_tabix_index_writer(tmp_path, data_dir):
    """tabix_index.write_tbi (dumpSTR --zip without the tabix binary): header, contig names and the whole linear index
    equal the reference's own tabix-made .tbi files; every chunk of the real index lies inside a chunk of ours at the
    same or a parent bin (htslib additionally merges small bins into their parents); and a file written by Writer,
    indexed here, serves region queries through the index exactly like a linear scan."""
    from trtools_b200.cyvcf2_compat import TextVCF, Writer
    from trtools_b200.tabix_index import write_tbi
    n_checked = 0
    for fname in sorted(os.listdir(data_dir)):
        path = os.path.join(data_dir, fname)
        if not fname.endswith(".vcf.gz") or not os.path.isfile(path + ".tbi"):
            continue
        mine = write_tbi(path, str(tmp_path / (fname + ".tbi")))
        a, b = _parse_tbi(path + ".tbi"), _parse_tbi(mine)
        assert a[0] == b[0] and a[1] == b[1], fname
        for (bins_a, lin_a), (bins_b, lin_b) in zip(a[2], b[2]):
            assert lin_a == lin_b, fname
            ours = sorted(c for cs in bins_b.values() for c in cs)
            for bin_id, chunks in bins_a.items():
                if bin_id == 37450:                         # htslib's pseudo-bin with mapped/unmapped counts
                    continue
                for cb, ce in chunks:
                    assert any(ob <= cb and ce <= oe for ob, oe in _merged(ours)), (fname, bin_id)
        n_checked += 1
    assert n_checked >= 2
    # Writer -> BGZF -> our index -> region queries
    src = os.path.join(data_dir, "many_samples.vcf.gz")
    tmpl = TextVCF(src)
    out = str(tmp_path / "written.vcf.gz")
    w = Writer(out, tmpl)
    for r in tmpl:
        w.write_record(r)
    w.close()
    write_tbi(out)
    v = NativeVCF(out)
    for region in ("1:1-20000", "1:164000-230000", "1:3000000-3100000", "1:900000", "2:1-100"):
        want = [(r.CHROM, r.POS) for r in TextVCF(src)(region)]
        got = [(r.CHROM, r.POS) for r in v(region)]
        assert got == want, region
        assert v._region_stop or v._region_empty, region      # served through the index, not by a linear scan


def _merged(chunks):
    out = []
    for b, e in chunks:
        if out and b <= out[-1][1]:
            out[-1][1] = max(out[-1][1], e)
        else:
            out.append([b, e])
    return out


def test_ragged_float_vectors_end_like_htslib_and_reserialise_unchanged(tmp_path):
    """A Number=. Float FORMAT field with rows of different lengths: the padding is BCF's float vector-end marker (a NaN,
    so 'missing' tests still hold), not a missing value — after a write-back of ANOTHER field re-serialises the record,
    '1.5' must not become '1.5,.' (the loop writer, the vectorised writer and the C++ serialiser all skip the marker)."""
    header = HEADER + '##FORMAT=<ID=XF,Number=.,Type=Float,Description="ragged float vector">\n'
    cols = ["0/1:5:1.5", "0/0:6:2,3", "1/1:7:.", "0|2:8:.,4"]
    for n_copies in (1, 40):               # 4 samples: the loop writer; 160 samples: the vectorised / C++ writers
        samples = tuple("S%d" % i for i in range(4 * n_copies))
        text = _vcf_text([_rec(100, "GT:DP:XF", cols * n_copies)], samples=samples, header=header)
        path = _write(tmp_path, "ragged%d.vcf" % n_copies, text)
        for reader in (cc.TextVCF, NativeVCF):
            rec = next(iter(reader(path)))
            xf = rec.format("XF")
            assert xf.shape == (4 * n_copies, 2) and xf.dtype == np.float32
            assert xf[0, 0] == 1.5 and np.isnan(xf[0, 1]) and cc.is_float_vector_end(xf)[0].tolist() == [False, True]
            assert cc.is_float_vector_end(xf)[2].tolist() == [False, True] and np.isnan(xf[2, 0])     # '.' then the end marker
            assert not cc.is_float_vector_end(xf)[3].any() and np.isnan(xf[3, 0]) and xf[3, 1] == 4
            dp = rec.format("DP").copy()
            dp[0] = 9
            rec.set_format("DP", dp)        # forces every decoded field to be re-serialised
            got = str(rec).rstrip("\n").split("\t")[9:]
            assert got[0] == "0/1:9:1.5" and got[1:4] == ["0/0:6:2,3", "1/1:7:.", "0|2:8:.,4"], (reader.__name__, got[:4])


def test_region_queries_through_csi_index_match_linear_scan(data_dir):
    """vcf(region) with a .csi next to the file (the reference's example-files/test_sample1.vcf.gz and its bcftools-made
    index, min_shift 14 / depth 6): htslib's min_off walk + overlapping-bin chunks give a start offset from which the
    reader serves exactly the records of the text reader's linear scan, and empty regions are answered from the index."""
    from trtools_b200 import vcf_ingest
    path = os.path.join(data_dir, "csi_indexed_sample.vcf.gz")
    assert os.path.isfile(path + ".csi") and not os.path.isfile(path + ".tbi")
    v = NativeVCF(path)
    for region in ("chr21:9483511-9490000", "chr21:20000000-20100000", "chr21:48000000", "chr21:1-100",
                   "chr21:30000000-30000500", "chr22:1-100", "chr21:14000000-14500000", "chr21:9483522-9483522", "chr21"):
        want = [(r.CHROM, r.POS, r.REF) for r in cc.TextVCF(path)(region)]
        got = [(r.CHROM, r.POS, r.REF) for r in v(region)]
        assert got == want, (region, len(got), len(want))
        assert v._region_stop or v._region_empty, region
    assert vcf_ingest._csi_start(path + ".csi", "chrNope", 1) == -1
    assert vcf_ingest._csi_start(path + ".nope.csi", "chr21", 1) is None


def test_native_tabix_index_multi_contig_and_unsorted(tmp_path):
    """Two contigs in one file: per-contig bins and linear index, region queries on both; an unsorted or interleaved
    file is refused the way tabix refuses it."""
    from trtools_b200.cyvcf2_compat import BgzfWriter, TextVCF
    from trtools_b200.tabix_index import write_tbi
    header = HEADER.replace("##contig=<ID=1,length=1000000>\n", "##contig=<ID=1,length=1000000>\n##contig=<ID=2,length=1000000>\n")

    def rec(chrom, pos):
        return _rec(pos, "GT:DP", ["0/1:5", "0/0:6", "1/1:7", "0|2:8"]).replace("1\t", chrom + "\t", 1)

    recs = [rec("1", p) for p in (100, 20000, 40000, 700000)] + [rec("2", p) for p in (50, 16500, 900000)]
    path = str(tmp_path / "two.vcf.gz")
    w = BgzfWriter(path)
    w.write(_vcf_text(recs, header=header).encode())
    w.close()
    write_tbi(path)
    a = _parse_tbi(path + ".tbi")
    assert a[1] == [b"1", b"2"] and len(a[2][0][1]) == (700000 >> 14) + 1 and len(a[2][1][1]) == (900000 >> 14) + 1
    v = NativeVCF(path)
    for region in ("1:1-30000", "1:600000", "2:1-17000", "2:800000-950000", "2:100-200", "3:1-10"):
        want = [(r.CHROM, r.POS) for r in TextVCF(path)(region)]
        got = [(r.CHROM, r.POS) for r in v(region)]
        assert got == want, region
    for bad in ([rec("1", 500), rec("1", 100)], [rec("1", 100), rec("2", 100), rec("1", 200)]):
        p2 = str(tmp_path / "bad.vcf.gz")
        w = BgzfWriter(p2)
        w.write(_vcf_text(bad, header=header).encode())
        w.close()
        with pytest.raises(ValueError):
            write_tbi(p2)
