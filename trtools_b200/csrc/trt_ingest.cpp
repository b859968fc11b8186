// trt_ingest.cpp — native block VCF ingest (host side of the ingest stage).
//
// The reference pulls one record at a time through cyvcf2/htslib:  vcfrecord.genotype.array()
// (trtools/utils/tr_harmonizer.py:829-862) and vcfrecord.format(key) (:561-588), each a text parse
// of one FORMAT column over all samples.  Here BGZF blocks are inflated in parallel, a run of
// records ("block") is kept as text, and GT plus the requested numeric FORMAT keys of every record
// of the block are parsed in one multi-threaded pass straight into the stacked [L][S] arrays (pinned
// host memory when the caller passes trt_host_alloc buffers) that trt_block_set_gt /
// trt_block_set_format_* copy to HBM.
//
// Conventions of the produced arrays (cyvcf2's, as the reference relies on them):
//   GT  int16 [L][S][P+1]  allele index, -1 for '.', -2 ploidy pad, last column = phased 0/1
//                          (tr_harmonizer.py:829-859)
//   Integer FORMAT  int32 [L][S], INT32_MIN for '.'        (dumpSTR/dumpSTR.py:736-742)
//   Float   FORMAT  float32 [L][S], NaN for '.'            (dumpSTR/filters.py:365,446)
// Anything this parser does not handle byte-for-byte like the text reader in cyvcf2_compat.py (odd
// allele tokens, vector-valued fields, ragged sample columns) is flagged per record / per key and the
// Python side re-parses that record from its text: flagged, never guessed.
//
// Region queries: trt_vcf_seek continues at a BGZF virtual offset taken from a tabix index (the Python side reads
// the .tbi); everything else about a region (overlap test, stopping) stays with the caller.
//
// Host code only: no CUDA calls, usable (and tested) without a GPU.  It feeds the CUDA path; it is
// not an alternative to it.
#include <sys/types.h>
#include <unistd.h>
#include <zlib.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <atomic>
#include <cerrno>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <functional>
#include <limits>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/trtools_b200.h"

namespace {

// file bytes pulled per fill (16 MiB; TRTOOLS_B200_INGEST_CHUNK overrides it so that the tests can put many
// refill boundaries — members and records straddling them — into small files).  A BGZF member is at most 64 KiB.
const size_t kCompressedChunk = [] {
    const char* e = getenv("TRTOOLS_B200_INGEST_CHUNK");
    long long n = e ? atoll(e) : 0;
    return n >= (128 << 10) ? (size_t)n : size_t(16) << 20;
}();
constexpr int    kMaxFmtKeys = 256;                     // FORMAT keys per record handled natively

void parallel_for(int64_t n, int n_threads, const std::function<void(int64_t)>& fn) {
    if (n <= 0) return;
    int nt = (int)std::min<int64_t>(n, std::max(1, n_threads));
    if (nt == 1) {
        for (int64_t i = 0; i < n; ++i) fn(i);
        return;
    }
    std::atomic<int64_t> next{0};
    std::vector<std::thread> th;
    th.reserve(nt);
    for (int t = 0; t < nt; ++t)
        th.emplace_back([&] {
            for (;;) {
                int64_t i = next.fetch_add(1, std::memory_order_relaxed);
                if (i >= n) break;
                fn(i);
            }
        });
    for (auto& t : th) t.join();
}

// Byte buffer without value-initialisation (std::vector<char>::resize would memset gigabytes of text
// that inflate / fread overwrite anyway); growth is realloc, which the allocator serves by mremap.
class RawBuf {
public:
    RawBuf() = default;
    explicit RawBuf(size_t n) { resize(n); }
    RawBuf(const RawBuf&) = delete;
    RawBuf& operator=(const RawBuf&) = delete;
    ~RawBuf() { free(p_); }
    char* data() { return p_; }
    const char* data() const { return p_; }
    size_t size() const { return n_; }
    char& operator[](size_t i) { return p_[i]; }
    const char& operator[](size_t i) const { return p_[i]; }
    void resize(size_t n) {
        if (n > cap_) {
            char* q = static_cast<char*>(realloc(p_, n));
            if (!q) throw std::bad_alloc();
            p_ = q;
            cap_ = n;
        }
        n_ = n;
    }
    void assign(const unsigned char* b, const unsigned char* e) {
        resize((size_t)(e - b));
        if (e > b) memcpy(p_, b, (size_t)(e - b));
    }
    void swap(RawBuf& o) {
        std::swap(p_, o.p_);
        std::swap(n_, o.n_);
        std::swap(cap_, o.cap_);
    }
private:
    char* p_ = nullptr;
    size_t n_ = 0, cap_ = 0;
};

struct BgzfBlock {
    size_t c_off, c_len;   // deflate payload inside the compressed chunk
    size_t o_off, o_len;   // where it lands in the text buffer
    uint32_t crc;
};

}  // namespace

struct trt_vcf_block {
    RawBuf text;                       // the block's records, '\n'-terminated, verbatim
    std::vector<int64_t> line_off;     // [n+1] start of each record in text
    std::vector<int64_t> fixed_len;    // [n] bytes of the first 9 columns (no trailing tab); -1 = malformed (<8 columns)
    std::vector<int64_t> samp_off;     // [n] offset (from line start) of the first sample column, -1 if none
    std::vector<int64_t> keep;         // kept sample columns (copy of the reader's at read time)
    int64_t n_file_samples = 0;
    int n_threads = 1;
};

struct trt_vcf {
    FILE* fh = nullptr;
    enum Kind { PLAIN, GZIP, BGZF } kind = PLAIN;
    std::vector<unsigned char> cbuf;   // compressed bytes not yet inflated
    size_t c_begin = 0, c_end = 0;
    bool file_eof = false;
    z_stream zs;                       // GZIP (non-BGZF) streaming state
    bool zs_live = false;
    RawBuf text;                       // inflated bytes; [t_pos, t_end) not yet handed out
    size_t t_pos = 0, t_end = 0;
    bool eof = false;
    int64_t plain_off = -1;            // PLAIN: file offset of the next unread byte when pread is usable
    std::string header;
    int64_t n_file_samples = 0;
    std::vector<int64_t> keep;
    int n_threads = 1;
    std::string err;
};

namespace {

thread_local std::string g_open_error;

int fail(trt_vcf* v, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (v) v->err = buf; else g_open_error = buf;
    return code;
}

void text_reserve(trt_vcf* v, size_t more) {
    if (v->t_end + more > v->text.size()) {
        size_t want = std::max(v->text.size() * 2, v->t_end + more);
        v->text.resize(want);
    }
}

// pull more compressed bytes from the file behind whatever is left in cbuf
size_t refill_compressed(trt_vcf* v, size_t chunk = kCompressedChunk) {
    if (v->c_begin > 0) {
        memmove(v->cbuf.data(), v->cbuf.data() + v->c_begin, v->c_end - v->c_begin);
        v->c_end -= v->c_begin;
        v->c_begin = 0;
    }
    if (v->cbuf.size() < v->c_end + chunk) v->cbuf.resize(v->c_end + chunk);
    size_t got = fread(v->cbuf.data() + v->c_end, 1, chunk, v->fh);
    v->c_end += got;
    if (got == 0) v->file_eof = true;
    return got;
}

// BGZF: gzip members with an extra subfield 'BC' carrying the member size (SAM spec §4.1)
bool bgzf_member_size(const unsigned char* p, size_t avail, size_t* total, size_t* hdr) {
    if (avail < 18) return false;
    if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return false;
    size_t xlen = p[10] | (size_t(p[11]) << 8);
    if (avail < 12 + xlen) return false;
    size_t q = 12, qe = 12 + xlen;
    while (q + 4 <= qe) {
        size_t slen = p[q + 2] | (size_t(p[q + 3]) << 8);
        if (p[q] == 'B' && p[q + 1] == 'C' && slen == 2 && q + 6 <= qe) {
            *total = (p[q + 4] | (size_t(p[q + 5]) << 8)) + 1;
            *hdr = 12 + xlen;
            return true;
        }
        q += 4 + slen;
    }
    return false;
}

int fill_bgzf(trt_vcf* v) {
    // gather the complete members available in cbuf (refilling once if there is none)
    for (int attempt = 0; attempt < 2; ++attempt) {
        std::vector<BgzfBlock> blocks;
        size_t c = v->c_begin, o = v->t_end;
        while (c < v->c_end) {
            size_t total, hdr;
            size_t avail = v->c_end - c;
            if (avail < 18) break;
            const unsigned char* m = v->cbuf.data() + c;
            if (m[0] != 0x1f || m[1] != 0x8b) return fail(v, TRT_ERECORD, "corrupt BGZF stream (bad member magic)");
            if (avail < 12 + (size_t(m[10]) | (size_t(m[11]) << 8))) break;     // header not buffered yet
            if (!bgzf_member_size(m, avail, &total, &hdr))
                return fail(v, TRT_ERECORD, "gzip member without a BGZF size field inside a BGZF file");
            if (avail < total) break;
            if (total < hdr + 8) return fail(v, TRT_ERECORD, "corrupt BGZF member (size field too small)");
            const unsigned char* t = v->cbuf.data() + c + total - 8;
            uint32_t crc = t[0] | (uint32_t(t[1]) << 8) | (uint32_t(t[2]) << 16) | (uint32_t(t[3]) << 24);
            uint32_t isz = t[4] | (uint32_t(t[5]) << 8) | (uint32_t(t[6]) << 16) | (uint32_t(t[7]) << 24);
            blocks.push_back({c + hdr, total - hdr - 8, o, isz, crc});
            o += isz;
            c += total;
        }
        if (blocks.empty()) {
            if (v->file_eof) {
                if (v->c_end > v->c_begin) return fail(v, TRT_ERECORD, "truncated BGZF member at end of file");
                v->eof = true;
                return TRT_OK;
            }
            refill_compressed(v);
            if (v->file_eof && v->c_end == v->c_begin) { v->eof = true; return TRT_OK; }
            continue;
        }
        text_reserve(v, o - v->t_end);
        std::atomic<int> bad{0};
        const unsigned char* cb = v->cbuf.data();
        char* tb = v->text.data();
        parallel_for((int64_t)blocks.size(), v->n_threads, [&](int64_t i) {
            const BgzfBlock& b = blocks[i];
            if (b.o_len == 0) return;
            z_stream zs;
            memset(&zs, 0, sizeof zs);
            if (inflateInit2(&zs, -15) != Z_OK) { bad = 1; return; }
            zs.next_in = const_cast<unsigned char*>(cb + b.c_off);
            zs.avail_in = (uInt)b.c_len;
            zs.next_out = reinterpret_cast<unsigned char*>(tb + b.o_off);
            zs.avail_out = (uInt)b.o_len;
            int rc = inflate(&zs, Z_FINISH);
            bool ok = (rc == Z_STREAM_END && zs.avail_out == 0);
            inflateEnd(&zs);
            if (ok) ok = (uint32_t)crc32(crc32(0L, Z_NULL, 0), reinterpret_cast<const unsigned char*>(tb + b.o_off),
                                         (uInt)b.o_len) == b.crc;
            if (!ok) bad = 1;
        });
        if (bad) return fail(v, TRT_ERECORD, "corrupt BGZF member (inflate / CRC mismatch)");
        v->c_begin = c;
        v->t_end = o;
        return TRT_OK;
    }
    return TRT_OK;
}

int fill_gzip(trt_vcf* v) {
    // plain (possibly multi-member) gzip: one sequential inflate stream
    if (v->c_begin == v->c_end) {
        if (!v->file_eof) refill_compressed(v);
        if (v->c_begin == v->c_end) {
            if (v->zs_live) return fail(v, TRT_ERECORD, "truncated gzip stream");
            v->eof = true;
            return TRT_OK;
        }
    }
    if (!v->zs_live) {
        memset(&v->zs, 0, sizeof v->zs);
        if (inflateInit2(&v->zs, 15 + 32) != Z_OK) return fail(v, TRT_ENOMEM, "inflateInit2 failed");
        v->zs_live = true;
    }
    const size_t out_chunk = size_t(8) << 20;
    text_reserve(v, out_chunk);
    v->zs.next_in = v->cbuf.data() + v->c_begin;
    v->zs.avail_in = (uInt)(v->c_end - v->c_begin);
    v->zs.next_out = reinterpret_cast<unsigned char*>(v->text.data() + v->t_end);
    v->zs.avail_out = (uInt)out_chunk;
    int rc = inflate(&v->zs, Z_NO_FLUSH);
    if (rc != Z_OK && rc != Z_STREAM_END && rc != Z_BUF_ERROR)
        return fail(v, TRT_ERECORD, "corrupt gzip stream (zlib rc %d)", rc);
    v->c_begin = v->c_end - v->zs.avail_in;
    v->t_end += out_chunk - v->zs.avail_out;
    if (rc == Z_STREAM_END) {            // next member, if any
        inflateEnd(&v->zs);
        v->zs_live = false;
    }
    return TRT_OK;
}

int fill_plain(trt_vcf* v) {
    // regular files: every host thread preads its own piece (page-cache copies scale with threads);
    // anything pread refuses (pipes) goes through the FILE
    // (at most 64 MiB per fill whatever the thread count: a fill is handed out block by block afterwards)
    const int np_cap = std::max(1, v->n_threads);
    const size_t kPiece = std::max(size_t(256) << 10, std::min(std::min(size_t(8) << 20, kCompressedChunk), (size_t(64) << 20) / (size_t)np_cap));
    if (v->plain_off >= 0 && v->n_threads > 1) {
        const int np = v->n_threads;
        text_reserve(v, kPiece * np);
        std::vector<ssize_t> got((size_t)np, 0);
        char* base = v->text.data() + v->t_end;
        const int fd = fileno(v->fh);
        const off_t off0 = (off_t)v->plain_off;
        parallel_for(np, np, [&](int64_t i) {
            size_t done = 0;
            while (done < kPiece) {
                ssize_t r = pread(fd, base + i * kPiece + done, kPiece - done, off0 + (off_t)(i * kPiece + done));
                if (r < 0 && errno == EINTR) continue;
                if (r <= 0) { if (r < 0 && done == 0) got[i] = -1; break; }
                done += (size_t)r;
            }
            if (got[i] == 0) got[i] = (ssize_t)done;
        });
        if (got[0] >= 0) {
            size_t total = 0;
            for (int i = 0; i < np; ++i) {
                if (got[i] < 0) return fail(v, TRT_ERECORD, "read error: %s", strerror(errno));
                total += (size_t)got[i];
                if ((size_t)got[i] < kPiece) break;          // end of file inside this piece
            }
            v->plain_off += (int64_t)total;
            v->t_end += total;
            if (total == 0) v->eof = true;
            return TRT_OK;
        }
        if (fseeko(v->fh, off0, SEEK_SET) != 0) v->plain_off = -1;   // not seekable: the FILE is where we left it
        v->plain_off = -1;
    }
    text_reserve(v, kCompressedChunk);
    size_t got = fread(v->text.data() + v->t_end, 1, kCompressedChunk, v->fh);
    v->t_end += got;
    if (got == 0) v->eof = true;
    return TRT_OK;
}

int fill(trt_vcf* v) {
    switch (v->kind) {
        case trt_vcf::BGZF: return fill_bgzf(v);
        case trt_vcf::GZIP: return fill_gzip(v);
        default: return fill_plain(v);
    }
}

// ---- record text -> arrays ------------------------------------------------------------------

struct KeySpec {
    std::string name;
    int is_float;
    void* out;          // int32 / float [n][S]
};

inline bool is_delim(const char* q, const char* le) { return q >= le || *q == ':' || *q == '\t'; }

// first ':' or '\t' at or after c (le if none), 16 bytes per step (SSE2 is baseline x86-64)
inline const char* skip_field(const char* c, const char* le) {
#if defined(__SSE2__)
    const __m128i colon = _mm_set1_epi8(':'), tab = _mm_set1_epi8('\t');
    while (le - c >= 16) {
        __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(c));
        int m = _mm_movemask_epi8(_mm_or_si128(_mm_cmpeq_epi8(v, colon), _mm_cmpeq_epi8(v, tab)));
        if (m) return c + __builtin_ctz((unsigned)m);
        c += 16;
    }
#endif
    while (c < le && *c != ':' && *c != '\t') ++c;
    return c;
}

// The field parsers scan and convert in one walk.  They return 1 = value parsed, 0 = missing ('.' or ''),
// -1 = a token this parser does not take (the caller flags the record's key for the Python parser);
// *end is the field's delimiter in every case.

// Python int() on the token, restricted to [-+]digits within int32
inline int field_i32(const char* a, const char* le, const char** end, int32_t* out) {
    const char* q = a;
    bool neg = false;
    if (q < le && (*q == '-' || *q == '+')) { neg = (*q == '-'); ++q; }
    const char* d0 = q;
    uint64_t x = 0;
    while (q < le && (unsigned)((unsigned char)*q - '0') <= 9u) { x = x * 10 + (unsigned)(*q - '0'); ++q; }
    if (is_delim(q, le)) {
        *end = q;
        size_t nd = (size_t)(q - d0);
        if (nd == 0) return q == a ? 0 : -1;
        if (nd > 10) return -1;
        int64_t v = neg ? -(int64_t)x : (int64_t)x;
        if (v < INT32_MIN || v > INT32_MAX) return -1;
        *out = (int32_t)v;
        return 1;
    }
    *end = skip_field(q, le);
    return (*end - a == 1 && *a == '.') ? 0 : -1;
}

const double kPow10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11,
                           1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

// exponents, long mantissas, inf/nan spellings: strtod (correctly rounded), on exactly the tokens both
// Python float() and strtod accept
bool parse_f32_general(const char* p, const char* e, float* out) {
    char tmp[64];
    size_t n = (size_t)(e - p);
    if (n == 0 || n >= sizeof tmp) return false;
    memcpy(tmp, p, n);
    tmp[n] = 0;
    for (size_t i = 0; i < n; ++i) {
        char ch = tmp[i];
        bool ok = (ch >= '0' && ch <= '9') || ch == '.' || ch == '-' || ch == '+' || ch == 'e' || ch == 'E';
        if (!ok) {
            const char* t = tmp + ((tmp[0] == '-' || tmp[0] == '+') ? 1 : 0);
            if (!strcasecmp(t, "nan") || !strcasecmp(t, "inf") || !strcasecmp(t, "infinity")) break;
            return false;
        }
    }
    char* endp = nullptr;
    double d = strtod(tmp, &endp);
    if (endp != tmp + n) return false;
    *out = (float)d;
    return true;
}

// np.float32(str): the decimal string rounds to the nearest double, then to float32.  Plain decimals with
// <= 15 significant digits and <= 22 fractional digits are exact by Clinger's argument (mantissa < 2^53, one
// correctly rounded division by an exactly representable power of ten).
inline int field_f32(const char* a, const char* le, const char** end, float* out) {
    const char* s = a;
    bool neg = false;
    if (s < le && (*s == '-' || *s == '+')) { neg = (*s == '-'); ++s; }
    uint64_t m = 0;
    const char* q = s;
    while (q < le && (unsigned)((unsigned char)*q - '0') <= 9u) { m = m * 10 + (unsigned)(*q - '0'); ++q; }
    long nt = q - s, frac = 0;
    if (q < le && *q == '.') {
        ++q;
        const char* f0 = q;
        while (q < le && (unsigned)((unsigned char)*q - '0') <= 9u) { m = m * 10 + (unsigned)(*q - '0'); ++q; }
        frac = q - f0;
        nt += frac;
    }
    if (is_delim(q, le)) {
        *end = q;
        if (nt > 0 && nt <= 15) {          // <= 15 digits in all: mantissa < 2^53, frac <= 15 <= 22
            double d = (double)m;
            if (frac) d /= kPow10[frac];
            *out = (float)(neg ? -d : d);
            return 1;
        }
        if (q == a || (q - a == 1 && *a == '.')) return 0;
        return parse_f32_general(a, q, out) ? 1 : -1;
    }
    *end = skip_field(q, le);
    return parse_f32_general(a, *end, out) ? 1 : -1;
}

struct LineResult {
    int ploidy;        // max number of alleles in a GT of this record (0: no samples)
    int status;        // 0 ok, 2 = re-parse this record in Python
};

// Parse one record's sample columns.  gt may be null (skip GT).  present[k]: 0 key absent, 1 parsed,
// 2 = this (record, key) needs the Python parser.
LineResult parse_line(const char* ls, const char* le, int64_t samp_off, const std::vector<int64_t>& keep,
                      int64_t n_file_samples, int P, int16_t* gt, const std::vector<KeySpec>& keys, int64_t rec,
                      int64_t S, uint8_t* present) {
    LineResult r{0, 0};
    const size_t nk = keys.size();
    for (size_t k = 0; k < nk; ++k) present[k] = 0;
    if (samp_off < 0) {
        if (n_file_samples > 0) r.status = 2;     // header names samples, record has none
        return r;
    }
    // FORMAT column = the 9th; it ends right before samp_off
    const char* fe = ls + samp_off - 1;
    const char* fs = fe;
    while (fs > ls && fs[-1] != '\t') --fs;
    int want[kMaxFmtKeys];
    int nfmt = 0, gt_idx = -1;
    {
        const char* a = fs;
        for (const char* c = fs;; ++c) {
            if (c == fe || *c == ':') {
                if (nfmt >= kMaxFmtKeys) { r.status = 2; return r; }
                want[nfmt] = -1;
                size_t len = (size_t)(c - a);
                if (len == 2 && a[0] == 'G' && a[1] == 'T') { if (gt_idx < 0) gt_idx = nfmt; }
                for (size_t k = 0; k < nk; ++k)
                    if (keys[k].name.size() == len && !memcmp(keys[k].name.data(), a, len) && present[k] == 0) {
                        // list.index(key): the first occurrence wins
                        want[nfmt] = (int)k;
                        present[k] = 1;
                    }
                ++nfmt;
                a = c + 1;
                if (c == fe) break;
            }
        }
    }
    const bool do_gt = (gt != nullptr);
    const int last_needed = [&] {
        int m = do_gt ? gt_idx : -1;
        for (int f = 0; f < nfmt; ++f) if (want[f] >= 0) m = std::max(m, f);
        return m;
    }();
    int16_t* gt_rec = do_gt ? gt + (size_t)rec * S * (P + 1) : nullptr;
    const bool all = keep.empty();
    size_t kp = 0;
    int64_t col = 0, outi = 0;
    const char* p = ls + samp_off;
    int maxparts = 0;
    while (p <= le) {
        // [p, ce) is sample column `col`
        bool kept = all ? true : (kp < keep.size() && keep[kp] == col);
        const char* c = p;
        if (!kept) {
            const void* t = memchr(p, '\t', (size_t)(le - p));
            c = t ? (const char*)t : le;
        } else {
            if (outi >= S) { r.status = 2; return r; }
            int f = 0;
            bool gt_seen = false;
            uint32_t seen_mask_lo = 0;   // keys parsed for this sample (the ABI caps a pass at 32 keys)
            for (;;) {
                const char* a = c;
                const bool is_gt = do_gt && f == gt_idx;
                const int k = f < nfmt ? want[f] : -1;
                if (is_gt) {
                    gt_seen = true;
                    int16_t* g = gt_rec + (size_t)outi * (P + 1);
                    int parts = 0;
                    int phased = 0;
                    if (P == 2 && le - c >= 4 && (unsigned)((unsigned char)c[0] - '0') <= 9u && (c[1] == '|' || c[1] == '/') &&
                        (unsigned)((unsigned char)c[2] - '0') <= 9u && (c[3] == ':' || c[3] == '\t')) {
                        g[0] = (int16_t)(c[0] - '0');
                        g[1] = (int16_t)(c[2] - '0');
                        g[2] = (int16_t)(c[1] == '|');
                        if (maxparts < 2) maxparts = 2;
                        c += 3;
                        goto gt_done;
                    }
                    for (;;) {
                        const char* ta = c;
                        int32_t x = 0;
                        bool digits = true;
                        while (c < le && *c != '/' && *c != '|' && *c != ':' && *c != '\t') {
                            unsigned d = (unsigned char)*c - '0';
                            if (d > 9) digits = false; else x = x * 10 + (int32_t)d;
                            if (x > 32767) digits = false, x = 0;
                            ++c;
                        }
                        int16_t val;
                        if (c == ta || (c - ta == 1 && *ta == '.')) val = -1;
                        else if (digits) val = (int16_t)x;
                        else { r.status = 2; return r; }
                        if (parts < P) g[parts] = val;
                        ++parts;
                        if (is_delim(c, le)) break;
                        if (parts == 1) phased = (*c == '|');
                        ++c;
                    }
                    for (int j = parts; j < P; ++j) g[j] = -2;
                    g[P] = (int16_t)phased;
                    if (parts > maxparts) maxparts = parts;
                gt_done:
                    if (k >= 0) { seen_mask_lo |= (1u << k); present[k] = 2; }    // GT asked for as a number
                } else if (k >= 0) {
                    seen_mask_lo |= (1u << k);
                    const KeySpec& ks = keys[k];
                    size_t o = (size_t)rec * S + outi;
                    if (ks.is_float) {
                        float x = std::numeric_limits<float>::quiet_NaN();
                        if (field_f32(a, le, &c, &x) < 0) present[k] = 2;
                        ((float*)ks.out)[o] = x;
                    } else {
                        int32_t x = INT32_MIN;
                        if (field_i32(a, le, &c, &x) < 0) present[k] = 2;
                        ((int32_t*)ks.out)[o] = x;
                    }
                } else {
                    c = skip_field(c, le);
                }
                ++f;
                if (c >= le || *c == '\t') break;
                ++c;
                if (f > last_needed) {   // nothing more to read in this column
                    const void* t = memchr(c, '\t', (size_t)(le - c));
                    c = t ? (const char*)t : le;
                    break;
                }
            }
            // trailing fields dropped by the caller (VCF allows it): they read as '.'
            if (do_gt && !gt_seen) {
                int16_t* g = gt_rec + (size_t)outi * (P + 1);
                if (gt_idx >= 0) {
                    g[0] = -1;
                    for (int j = 1; j < P; ++j) g[j] = -2;
                    g[P] = 0;
                    if (maxparts < 1) maxparts = 1;
                } else {
                    r.status = 2;      // record without GT: served by the Python reader
                    return r;
                }
            }
            if (f <= last_needed) for (size_t k = 0; k < nk; ++k) {
                if (present[k] == 0) continue;
                if (!((seen_mask_lo >> k) & 1u)) {
                    size_t o = (size_t)rec * S + outi;
                    if (keys[k].is_float) ((float*)keys[k].out)[o] = std::numeric_limits<float>::quiet_NaN();
                    else ((int32_t*)keys[k].out)[o] = INT32_MIN;
                }
            }
            ++outi;
            ++kp;
        }
        ++col;
        if (c >= le) break;
        p = c + 1;
    }
    if (col != n_file_samples || outi != S) r.status = 2;
    r.ploidy = maxparts;
    return r;
}

}  // namespace

// ---- C-ABI ------------------------------------------------------------------------------------

extern "C" {

const char* trt_vcf_last_error(const trt_vcf* v) { return v ? v->err.c_str() : g_open_error.c_str(); }

static int vcf_open_impl(const char* path, int n_threads, trt_vcf** out) {
    if (!path || !out) return fail(nullptr, TRT_EINVAL, "trt_vcf_open: null argument");
    *out = nullptr;
    FILE* fh = fopen(path, "rb");
    if (!fh) return fail(nullptr, TRT_EINVAL, "Error opening %s: %s", path, strerror(errno));
    trt_vcf* v = new trt_vcf();
    v->fh = fh;
    if (n_threads <= 0) {
        unsigned hc = std::thread::hardware_concurrency();
        n_threads = hc ? (int)hc : 1;
    }
    v->n_threads = n_threads;
    refill_compressed(v, size_t(1) << 20);      // enough for the header: opening a file only for its samples stays cheap
    size_t avail = v->c_end - v->c_begin;
    if (avail >= 2 && v->cbuf[0] == 0x1f && v->cbuf[1] == 0x8b) {
        size_t total, hdr;
        v->kind = bgzf_member_size(v->cbuf.data(), avail, &total, &hdr) ? trt_vcf::BGZF : trt_vcf::GZIP;
    } else {
        v->kind = trt_vcf::PLAIN;
        v->text.assign(v->cbuf.data() + v->c_begin, v->cbuf.data() + v->c_end);
        v->t_end = v->text.size();
        v->c_begin = v->c_end = 0;
        if (v->file_eof && v->t_end == 0) v->eof = true;
        off_t at = ftello(fh);
        v->plain_off = (at >= 0 && (size_t)at == v->t_end) ? (int64_t)at : -1;
    }
    // header: every leading line that starts with '#'
    for (;;) {
        bool done = false;
        while (v->t_pos < v->t_end) {
            if (v->text[v->t_pos] != '#') { done = true; break; }
            const void* nl = memchr(v->text.data() + v->t_pos, '\n', v->t_end - v->t_pos);
            if (!nl) {
                if (!v->eof) break;
                v->header.append(v->text.data() + v->t_pos, v->t_end - v->t_pos);   // unterminated last line
                v->t_pos = v->t_end;
                break;
            }
            size_t e = (const char*)nl - v->text.data() + 1;
            v->header.append(v->text.data() + v->t_pos, e - v->t_pos);
            v->t_pos = e;
        }
        if (done || v->eof) break;
        int rc = fill(v);
        if (rc != TRT_OK) {
            g_open_error = v->err;
            fclose(fh);
            delete v;
            return rc;
        }
    }
    // sample count from the last #CHROM line
    size_t at = v->header.rfind("#CHROM");
    while (at != std::string::npos && at != 0 && v->header[at - 1] != '\n') at = at ? v->header.rfind("#CHROM", at - 1) : std::string::npos;
    if (at != std::string::npos) {
        size_t e = v->header.find('\n', at);
        if (e == std::string::npos) e = v->header.size();
        int64_t tabs = 0;
        for (size_t i = at; i < e; ++i) tabs += (v->header[i] == '\t');
        v->n_file_samples = tabs >= 9 ? tabs - 8 : 0;
    }
    *out = v;
    return TRT_OK;
}

void trt_vcf_close(trt_vcf* v) {
    if (!v) return;
    if (v->zs_live) inflateEnd(&v->zs);
    if (v->fh) fclose(v->fh);
    delete v;
}

int trt_vcf_header(trt_vcf* v, const char** text, int64_t* len) {
    if (!v || !text || !len) return TRT_EINVAL;
    *text = v->header.data();
    *len = (int64_t)v->header.size();
    return TRT_OK;
}

int64_t trt_vcf_n_samples(const trt_vcf* v) { return v ? v->n_file_samples : 0; }

int trt_vcf_set_samples(trt_vcf* v, const int64_t* cols, int64_t n) {
    if (!v || (n > 0 && !cols)) return TRT_EINVAL;
    v->keep.assign(cols, cols + n);
    for (int64_t i = 0; i < n; ++i)
        if (cols[i] < 0 || cols[i] >= v->n_file_samples || (i && cols[i] <= cols[i - 1]))
            return fail(v, TRT_EINVAL, "trt_vcf_set_samples: columns must be strictly increasing and in range");
    if (n == 0) v->keep.assign(1, -1);   // keep nothing (an empty list would mean "all")
    return TRT_OK;
}

int trt_vcf_seek(trt_vcf* v, int64_t coffset, int32_t uoffset) {
    // BGZF virtual file offset (tabix / CSI): member at byte coffset, uoffset bytes into its inflated data
    if (!v || coffset < 0 || uoffset < 0 || uoffset > 65536) return TRT_EINVAL;
    if (v->kind != trt_vcf::BGZF) return fail(v, TRT_ESTATE, "trt_vcf_seek: not a BGZF file");
    if (fseeko(v->fh, (off_t)coffset, SEEK_SET) != 0) return fail(v, TRT_EINVAL, "trt_vcf_seek: %s", strerror(errno));
    v->c_begin = v->c_end = 0;
    v->file_eof = false;
    v->t_pos = v->t_end = 0;
    v->eof = false;
    while (v->t_end < (size_t)uoffset && !v->eof) {
        int rc = fill(v);
        if (rc != TRT_OK) return rc;
    }
    if (v->t_end < (size_t)uoffset) return fail(v, TRT_EINVAL, "trt_vcf_seek: offset beyond the end of the file");
    v->t_pos = (size_t)uoffset;
    return TRT_OK;
}

static int vcf_read_block_impl(trt_vcf* v, int64_t max_loci, int64_t max_bytes, trt_vcf_block** out, int64_t* n_loci) {
    if (!v || !out || !n_loci || max_loci <= 0) return TRT_EINVAL;
    *out = nullptr;
    *n_loci = 0;
    if (max_bytes <= 0) max_bytes = std::numeric_limits<int64_t>::max();
    std::vector<int64_t> off;
    // [t_pos, t_end) of the reader's buffer is unread text.  Nothing is moved per call: the consumed prefix is only
    // compacted away when it is at least as large as what is left (amortised O(1) per byte), and a block that is
    // small next to the unread tail gets a copy of ITS bytes instead of taking the buffer (the tail stays put) —
    // files of many short records (few samples) used to pay a copy of the whole 16+ MiB fill per block.
    size_t scan = v->t_pos;
    size_t first = v->t_pos;
    // drop blank lines in front (the text reader skips them)
    for (;;) {
        while (off.size() < (size_t)max_loci && scan < v->t_end) {
            if (off.empty()) {
                while (scan < v->t_end && v->text[scan] == '\n') ++scan;
                first = scan;
                if (scan == v->t_end) break;
            }
            const void* nl = memchr(v->text.data() + scan, '\n', v->t_end - scan);
            if (!nl) break;
            size_t e = (const char*)nl - v->text.data() + 1;
            off.push_back((int64_t)(scan - first));
            scan = e;
            while (scan < v->t_end && v->text[scan] == '\n') ++scan;   // blank lines between records
            if ((int64_t)(scan - first) >= max_bytes) break;
        }
        if (off.size() >= (size_t)max_loci || (!off.empty() && (int64_t)(scan - first) >= max_bytes)) break;
        if (v->eof) {
            if (scan < v->t_end) {   // unterminated last record
                off.push_back((int64_t)(scan - first));
                scan = v->t_end;
            }
            break;
        }
        {
            // make room for the refill: drop the consumed prefix once it outweighs the unread text
            const size_t keep_from = off.empty() ? scan : first;
            if (keep_from > 0 && keep_from >= v->t_end - keep_from) {
                memmove(v->text.data(), v->text.data() + keep_from, v->t_end - keep_from);
                v->t_end -= keep_from;
                scan -= keep_from;
                first -= std::min(first, keep_from);
                v->t_pos = 0;
            }
        }
        int rc = fill(v);
        if (rc != TRT_OK) return rc;
    }
    if (off.empty()) {
        v->t_pos = v->t_end = 0;
        return TRT_OK;
    }
    trt_vcf_block* b = new trt_vcf_block();
    size_t nbytes = scan - first;
    size_t tail = v->t_end - scan;
    if (nbytes + 1 < tail) {
        // small block, long tail: the block copies its own records
        b->text.resize(nbytes + 1);
        memcpy(b->text.data(), v->text.data() + first, nbytes);
        v->t_pos = scan;
    } else {
        // the block takes the reader's buffer; the (shorter) unread tail is copied into a fresh one
        RawBuf rest(std::max(tail + kCompressedChunk, size_t(1) << 20));
        memcpy(rest.data(), v->text.data() + scan, tail);
        b->text.swap(v->text);
        v->text.swap(rest);
        v->t_pos = 0;
        v->t_end = tail;
        if (first > 0) memmove(b->text.data(), b->text.data() + first, nbytes);
    }
    bool unterminated = b->text[nbytes - 1] != '\n';
    b->text.resize(nbytes + (unterminated ? 1 : 0));
    if (unterminated) b->text[nbytes] = '\n';
    off.push_back((int64_t)b->text.size());
    b->line_off.swap(off);
    b->keep = v->keep;
    b->n_file_samples = v->n_file_samples;
    b->n_threads = v->n_threads;
    int64_t n = (int64_t)b->line_off.size() - 1;
    b->fixed_len.assign(n, -1);
    b->samp_off.assign(n, -1);
    const char* t = b->text.data();
    parallel_for(n, n > 64 ? b->n_threads : 1, [&](int64_t i) {
        const char* ls = t + b->line_off[i];
        const char* le = t + b->line_off[i + 1];
        // the record ends at the first '\n' (blank lines may follow it)
        le = (const char*)memchr(ls, '\n', (size_t)(le - ls));
        const char* p = ls;
        int tabs = 0;
        while (tabs < 9) {
            const void* q = memchr(p, '\t', (size_t)(le - p));
            if (!q) break;
            ++tabs;
            p = (const char*)q + 1;
        }
        if (tabs == 9) {
            b->fixed_len[i] = (p - 1) - ls;
            b->samp_off[i] = p - ls;
        } else if (tabs >= 7) {
            int64_t len = le - ls;
            if (len > 0 && le[-1] == '\r') --len;
            b->fixed_len[i] = len;     // 8 or 9 columns, no samples
        }
    });
    *out = b;
    *n_loci = n;
    return TRT_OK;
}

void trt_vcf_block_free(trt_vcf_block* b) { delete b; }

int trt_vcf_block_text(const trt_vcf_block* b, const char** text, const int64_t** line_off,
                       const int64_t** fixed_len) {
    if (!b || !text || !line_off || !fixed_len) return TRT_EINVAL;
    *text = b->text.data();
    *line_off = b->line_off.data();
    *fixed_len = b->fixed_len.data();
    return TRT_OK;
}

// gt2_out != NULL: the packed transfer form (trt_block_set_gt_packed) instead of gt_out — two bytes per call (allele
// 0..252, 254 = ploidy pad, 255 = no-call) and one phase bit per call.  The record is parsed into a per-thread int16
// row (cache-resident) and packed from there; a record with an allele index above 252 or more than two haplotypes
// gets rec_status 3 and the caller parses the block in the plain form.
static int vcf_block_parse_impl(const trt_vcf_block* b, int ploidy, int16_t* gt_out, int n_keys, const char* const* keys,
                                const int32_t* key_is_float, void* const* key_out, uint8_t* present, int32_t* rec_ploidy,
                                uint8_t* rec_status, uint8_t* gt2_out = nullptr, uint8_t* phase_out = nullptr,
                                bool nibble = false) {
    if (!b || n_keys < 0 || n_keys > 32 || !rec_ploidy || !rec_status || (n_keys && (!keys || !key_out || !present)))
        return TRT_EINVAL;
    if (gt_out && ploidy < 1) return TRT_EINVAL;
    if (gt2_out && (gt_out || ploidy != 2)) return TRT_EINVAL;
    std::vector<KeySpec> ks((size_t)n_keys);
    for (int k = 0; k < n_keys; ++k) {
        ks[k].name = keys[k];
        ks[k].is_float = key_is_float[k];
        ks[k].out = key_out[k];
        if (!key_out[k]) return TRT_EINVAL;
    }
    const int64_t n = (int64_t)b->line_off.size() - 1;
    const bool all = b->keep.empty();
    const int64_t S = all ? b->n_file_samples : ((b->keep.size() == 1 && b->keep[0] < 0) ? 0 : (int64_t)b->keep.size());
    const char* t = b->text.data();
    parallel_for(n, b->n_threads, [&](int64_t i) {
        uint8_t dummy[1];
        uint8_t* pres = n_keys ? present + (size_t)i * n_keys : dummy;
        if (b->fixed_len[i] < 0) {          // malformed: the Python side raises when it reaches the record
            rec_ploidy[i] = 0;
            rec_status[i] = 1;
            for (int k = 0; k < n_keys; ++k) pres[k] = 0;
            return;
        }
        const char* ls = t + b->line_off[i];
        const char* le = (const char*)memchr(ls, '\n', (size_t)(t + b->line_off[i + 1] - ls));
        if (le > ls && le[-1] == '\r') --le;
        if (!gt2_out) {
            LineResult r = parse_line(ls, le, b->samp_off[i], b->keep, b->n_file_samples, ploidy, gt_out, ks, i, S, pres);
            rec_ploidy[i] = r.ploidy;
            rec_status[i] = (uint8_t)r.status;
            return;
        }
        static thread_local std::vector<int16_t> row;
        row.resize((size_t)S * 3 + 8);
        // parse_line addresses record i of a [n][S][3] array: hand it the row shifted back by i records
        LineResult r = parse_line(ls, le, b->samp_off[i], b->keep, b->n_file_samples, 2, row.data() - (size_t)i * S * 3, ks, i, S, pres);
        rec_ploidy[i] = r.ploidy;
        rec_status[i] = (uint8_t)r.status;
        if (r.status != 0) return;
        if (r.ploidy > 2) { rec_status[i] = 3; return; }
        const size_t pbytes = (size_t)(S + 7) / 8;
        uint8_t* ph = phase_out ? phase_out + (size_t)i * pbytes : nullptr;
        if (ph) memset(ph, 0, pbytes);
        const int16_t* g = row.data();
        bool fits = true;
        if (nibble) {                       // one byte per call: low nibble first haplotype; -1 -> 15, -2 -> 14
            uint8_t* o4 = gt2_out + (size_t)i * S;
            for (int64_t s = 0; s < S; ++s, g += 3) {
                const int a0 = g[0], a1 = g[1];
                fits = fits && a0 <= 13 && a1 <= 13 && a0 >= -2 && a1 >= -2;
                o4[s] = (uint8_t)((a0 & 15) | ((a1 & 15) << 4));
                if (ph && g[2]) ph[s >> 3] |= (uint8_t)(1u << (s & 7));
            }
            if (!fits) rec_status[i] = 3;
            return;
        }
        uint8_t* o = gt2_out + (size_t)i * S * 2;
        for (int64_t s = 0; s < S; ++s, g += 3) {
            const int a0 = g[0], a1 = g[1];
            fits = fits && a0 <= 252 && a1 <= 252 && a0 >= -2 && a1 >= -2;
            o[2 * s] = (uint8_t)(a0 >= 0 ? a0 : 256 + a0);          // -1 -> 255, -2 -> 254
            o[2 * s + 1] = (uint8_t)(a1 >= 0 ? a1 : 256 + a1);
            if (ph && g[2]) ph[s >> 3] |= (uint8_t)(1u << (s & 7));
        }
        if (!fits) rec_status[i] = 3;
    });
    return TRT_OK;
}

// Raw text of one FORMAT field (by position in the record's FORMAT column) of every kept sample of one record, as
// fixed-width NUL-padded byte strings [S][width]; a sample column that ends before the field reads '.' (what the
// text reader returns).  *max_len receives the longest token; nothing is written unless out != NULL and
// width >= *max_len.  Returns TRT_OK, or TRT_ERECORD when the record's columns do not match the header.
int trt_vcf_block_field(const trt_vcf_block* b, int64_t rec, int field_index, int32_t width, char* out,
                        int32_t* max_len) {
    if (!b || !max_len || rec < 0 || rec >= (int64_t)b->line_off.size() - 1 || field_index < 0) return TRT_EINVAL;
    if (b->samp_off[rec] < 0) return TRT_ERECORD;
    const char* t = b->text.data();
    const char* ls = t + b->line_off[rec];
    const char* le = (const char*)memchr(ls, '\n', (size_t)(t + b->line_off[rec + 1] - ls));
    if (le > ls && le[-1] == '\r') --le;
    const bool all = b->keep.empty();
    const int64_t S = all ? b->n_file_samples : ((b->keep.size() == 1 && b->keep[0] < 0) ? 0 : (int64_t)b->keep.size());
    for (int pass = 0; pass < 2; ++pass) {
        const bool fill = pass == 1;
        if (fill && (!out || width < *max_len)) return TRT_OK;
        if (fill) memset(out, 0, (size_t)S * (size_t)width);
        int32_t longest = 1;
        size_t kp = 0;
        int64_t col = 0, outi = 0;
        const char* p = ls + b->samp_off[rec];
        while (p <= le) {
            const void* tab = memchr(p, '\t', (size_t)(le - p));
            const char* ce = tab ? (const char*)tab : le;
            bool kept = all ? true : (kp < b->keep.size() && b->keep[kp] == col);
            if (kept) {
                if (outi >= S) return TRT_ERECORD;
                const char* a = p;
                int f = 0;
                const char* c = a;
                for (;;) {
                    c = a;
                    while (c < ce && *c != ':') ++c;
                    if (f == field_index || c >= ce) break;
                    ++f;
                    a = c + 1;
                }
                const char* tok = a;
                size_t n = (size_t)(c - a);
                if (f != field_index) { tok = "."; n = 1; }
                if ((int32_t)n > longest) longest = (int32_t)n;
                if (fill) memcpy(out + (size_t)outi * (size_t)width, tok, n);
                ++outi;
                ++kp;
            }
            ++col;
            if (!tab) break;
            p = ce + 1;
        }
        if (col != b->n_file_samples || outi != S) return TRT_ERECORD;
        if (!fill) *max_len = longest;
    }
    return TRT_OK;
}

// ---- record serialisation (the dumpSTR writer's inner loop; trtools/dumpSTR/dumpSTR.py:1338 write_record) ------
// Sample columns of one record: per sample the fields joined by ':', samples joined by '\t'.  Field kinds:
//   0 fixed-width byte strings [S] (numpy 'S', NUL padded)   1 GT int16 [S][ncol] (cyvcf2 layout, ncol = P+1)
//   2 int32 [S][ncol] (INT32_MIN -> '.', INT32_MIN+1 = vector end, skipped)   3 float32 / 4 float64 [S][ncol]
//   ('%g', NaN -> '.').  A vector prints its entries joined by ',', '.' if none is left.
// Returns the bytes written, or -(bytes needed) when cap is too small (nothing useful is written then).
int64_t trt_vcf_join_samples(int64_t n_samples, int n_fields, const int32_t* kind, const void* const* data,
                             const int32_t* ncol, char* out, int64_t cap) {
    if (n_samples < 0 || n_fields < 1 || !kind || !data || !ncol || (!out && cap > 0)) return 0;
    int64_t w = 0;
    char num[64];
    auto put = [&](const char* p, size_t n) {
        if (w + (int64_t)n <= cap) memcpy(out + w, p, n);
        w += (int64_t)n;
    };
    auto putc_ = [&](char c) {
        if (w < cap) out[w] = c;
        ++w;
    };
    for (int64_t s = 0; s < n_samples; ++s) {
        if (s) putc_('\t');
        for (int f = 0; f < n_fields; ++f) {
            if (f) putc_(':');
            const int nc = ncol[f];
            switch (kind[f]) {
                case 0: {
                    const char* p = (const char*)data[f] + (size_t)s * nc;
                    put(p, strnlen(p, (size_t)nc));
                    break;
                }
                case 1: {
                    const int16_t* g = (const int16_t*)data[f] + (size_t)s * nc;
                    const int P = nc - 1;
                    const char sep = g[P] ? '|' : '/';
                    int printed = 0;
                    for (int j = 0; j < P; ++j) {
                        if (g[j] == -2) continue;
                        if (printed) putc_(sep);
                        if (g[j] == -1) putc_('.');
                        else put(num, (size_t)snprintf(num, sizeof num, "%d", (int)g[j]));
                        ++printed;
                    }
                    if (!printed) putc_('.');
                    break;
                }
                case 2: {
                    const int32_t* v = (const int32_t*)data[f] + (size_t)s * nc;
                    int printed = 0;
                    for (int j = 0; j < nc; ++j) {
                        if (v[j] == INT32_MIN + 1) continue;
                        if (printed) putc_(',');
                        if (v[j] == INT32_MIN) putc_('.');
                        else put(num, (size_t)snprintf(num, sizeof num, "%d", (int)v[j]));
                        ++printed;
                    }
                    if (!printed) putc_('.');
                    break;
                }
                case 3:
                case 4: {
                    int printed = 0;
                    for (int j = 0; j < nc; ++j) {
                        double x;
                        bool vector_end;                  // BCF's float vector-end NaN (payload 2): a ragged row's padding
                        if (kind[f] == 3) {
                            uint32_t u;
                            memcpy(&u, (const float*)data[f] + (size_t)s * nc + j, 4);
                            vector_end = (u & 0x7F800000u) == 0x7F800000u && (u & 0x003FFFFFu) == 2u;
                            float fl;
                            memcpy(&fl, &u, 4);
                            x = (double)fl;
                        } else {
                            uint64_t u;
                            memcpy(&u, (const double*)data[f] + (size_t)s * nc + j, 8);
                            vector_end = (u & 0x7FF0000000000000ull) == 0x7FF0000000000000ull &&
                                         (u & 0x0007FFFFFFFFFFFFull) == (2ull << 29);
                            memcpy(&x, &u, 8);
                        }
                        if (vector_end) continue;
                        if (printed) putc_(',');
                        if (std::isnan(x)) putc_('.');
                        else put(num, (size_t)snprintf(num, sizeof num, "%g", x));
                        ++printed;
                    }
                    if (!printed) putc_('.');
                    break;
                }
                default:
                    return 0;
            }
        }
    }
    return w <= cap ? w : -w;
}

// no exception crosses the ABI (allocation failures of multi-gigabyte blocks included)
int trt_vcf_open(const char* path, int n_threads, trt_vcf** out) {
    try {
        return vcf_open_impl(path, n_threads, out);
    } catch (const std::exception& e) {
        return fail(nullptr, TRT_ENOMEM, "trt_vcf_open: %s", e.what());
    }
}

int trt_vcf_read_block(trt_vcf* v, int64_t max_loci, int64_t max_bytes, trt_vcf_block** out, int64_t* n_loci) {
    try {
        return vcf_read_block_impl(v, max_loci, max_bytes, out, n_loci);
    } catch (const std::exception& e) {
        return fail(v, TRT_ENOMEM, "trt_vcf_read_block: %s", e.what());
    }
}

int trt_vcf_block_parse(const trt_vcf_block* b, int ploidy, int16_t* gt_out, int n_keys, const char* const* keys,
                        const int32_t* key_is_float, void* const* key_out, uint8_t* present, int32_t* rec_ploidy,
                        uint8_t* rec_status) {
    try {
        return vcf_block_parse_impl(b, ploidy, gt_out, n_keys, keys, key_is_float, key_out, present, rec_ploidy,
                                    rec_status);
    } catch (const std::exception&) {
        return TRT_ENOMEM;
    }
}

int trt_vcf_block_parse_packed(const trt_vcf_block* b, uint8_t* gt2_out, uint8_t* phase_out, int n_keys,
                               const char* const* keys, const int32_t* key_is_float, void* const* key_out, uint8_t* present,
                               int32_t* rec_ploidy, uint8_t* rec_status) {
    if (!gt2_out) return TRT_EINVAL;
    try {
        return vcf_block_parse_impl(b, 2, nullptr, n_keys, keys, key_is_float, key_out, present, rec_ploidy, rec_status,
                                    gt2_out, phase_out);
    } catch (const std::exception&) {
        return TRT_ENOMEM;
    }
}

int trt_vcf_block_parse_nibble(const trt_vcf_block* b, uint8_t* g4_out, uint8_t* phase_out, int n_keys,
                               const char* const* keys, const int32_t* key_is_float, void* const* key_out, uint8_t* present,
                               int32_t* rec_ploidy, uint8_t* rec_status) {
    if (!g4_out) return TRT_EINVAL;
    try {
        return vcf_block_parse_impl(b, 2, nullptr, n_keys, keys, key_is_float, key_out, present, rec_ploidy, rec_status,
                                    g4_out, phase_out, true);
    } catch (const std::exception&) {
        return TRT_ENOMEM;
    }
}

}  // extern "C"
