// K1 — harmonize: one warp per locus works on the locus' allele table (strings + INFO scalars)
// and emits, per allele, the trimmed window, the repeat-unit length, the length/sequence
// equivalence classes and their sort orders; per locus, the inferred motif, the homopolymer run
// of the full REF and flag bits.  Replaces the Python string work of
//   _HarmonizeHipSTRRecord  trtools/utils/tr_harmonizer.py:336-408   (trim with python-slice
//       semantics :377-394, full_alleles :360-369, motif :397)
//   _Harmonize{GangSTR,AdVNTR,PopSTR,EH}Record :303-333, :411-436, :473-512, :515-550
//   TRRecord.__init__ :693-773 (lengths :740,757-759)
//   utils.InferRepeatSequence utils.py:465-508, GetCanonicalOneStrand :396-427,
//   FabricateAllele :566-602, GetHomopolymerRun :340-360
//   TRRecord.UniqueStringGenotypeMapping :1049-1082 / UniqueLengthGenotypeMapping :1247-1273
// The work is O(alleles x bases) per locus — tiny next to the sample axis — so the kernel is
// latency-bound by design; it exists to keep Python out of the per-locus loop.
#include <math.h>

#include <algorithm>

#include <stdlib.h>

#include <algorithm>

#include "trt_internal.cuh"

namespace {

struct HarmParams {
    const char* seqs;
    const int64_t* allele_off;
    const int32_t* locus_off;
    const int32_t* pos;
    const int32_t* start;
    const int32_t* end;
    const int32_t* period;
    const double* given_len;
    const char* motif_in;  // may be null; motif bytes per locus at motif_off (needed for fabricated alleles)
    const int64_t* motif_off;
    int64_t L;
    int vcftype;
    // outputs
    double* allele_len;
    int32_t* trim_off;
    int32_t* trim_len;
    int32_t* len_class;
    int32_t* seq_class;
    int32_t* len_order;
    int32_t* seq_order;
    int32_t* hrun;
    int32_t* flags;
    char* motif;
};

__device__ __forceinline__ char up(char c) { return (c >= 'a' && c <= 'z') ? (char)(c - 32) : c; }

// python slice index normalisation for a sequence of length n
__device__ __forceinline__ int py_idx(int i, int n) {
    if (i < 0) {
        i += n;
        if (i < 0) i = 0;
    } else if (i > n) {
        i = n;
    }
    return i;
}

// An allele as a string: a window of the VCF bytes (upper-cased on read) or motif^k fabricated.
struct AlleleStr {
    const char* base;  // window start (real) or motif bytes (fabricated)
    int len;           // number of characters
    int mlen;          // 0 = real window; else motif length (fabricated)
    __device__ __forceinline__ char at(int i) const { return mlen ? up(base[i % mlen]) : up(base[i]); }
};

// 8 upper-cased characters starting at p (any alignment): two aligned 64-bit loads and a funnel shift; bytes past
// the string are garbage and must be masked by the caller (the sequence buffer has 16 bytes of slack)
__device__ __forceinline__ unsigned long long load8_up(const char* p) {
    const unsigned long long* base = (const unsigned long long*)((uintptr_t)p & ~(uintptr_t)7);
    const unsigned sh = (unsigned)((uintptr_t)p & 7) * 8u;
    const unsigned long long lo = base[0], hi = base[1];
    unsigned long long x = sh ? ((lo >> sh) | (hi << (64u - sh))) : lo;
    // SWAR str.upper(): subtract 0x20 from the bytes in 'a'..'z'
    const unsigned long long t = x & 0x7f7f7f7f7f7f7f7full;
    const unsigned long long ge_a = t + 0x1f1f1f1f1f1f1f1full;      // bit 7 set: (byte & 0x7f) >= 'a'
    const unsigned long long gt_z = t + 0x0505050505050505ull;      // bit 7 set: (byte & 0x7f) >  'z'
    const unsigned long long lower = ge_a & ~gt_z & ~x & 0x8080808080808080ull;
    return x - (lower >> 2);
}

// lexicographic compare of two real (non-fabricated) allele windows, 8 characters per step
__device__ int str_cmp_real(const char* pa, int la, const char* pb, int lb) {
    const int n = la < lb ? la : lb;
    for (int i = 0; i < n; i += 8) {
        unsigned long long x = load8_up(pa + i), y = load8_up(pb + i);
        const int rem = n - i;
        if (rem < 8) {
            const unsigned long long m = (1ull << (8 * rem)) - 1ull;
            x &= m;
            y &= m;
        }
        if (x != y) {
            const int byte = (__ffsll((long long)(x ^ y)) - 1) >> 3;     // first differing character (little endian)
            const unsigned cx = (unsigned)(x >> (8 * byte)) & 0xffu, cy = (unsigned)(y >> (8 * byte)) & 0xffu;
            return cx < cy ? -1 : 1;
        }
    }
    return la == lb ? 0 : (la < lb ? -1 : 1);
}

__device__ int str_cmp(const AlleleStr& a, const AlleleStr& b) {
    if (a.mlen == 0 && b.mlen == 0) return str_cmp_real(a.base, a.len, b.base, b.len);
    int n = a.len < b.len ? a.len : b.len;
    for (int i = 0; i < n; i++) {
        unsigned char x = (unsigned char)a.at(i), y = (unsigned char)b.at(i);
        if (x != y) return x < y ? -1 : 1;
    }
    return a.len == b.len ? 0 : (a.len < b.len ? -1 : 1);
}

__device__ __forceinline__ int nuc_code(char c) {  // utils.py:17 nucToNumber
    return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1;
}

// compare rotation r1 vs r2 of the p-mer at s (rotation r = kmer[p-r:] + kmer[:p-r]) under A<C<G<T
__device__ int rot_cmp(const char* s, int p, int r1, int r2) {
    for (int j = 0; j < p; j++) {
        int a = nuc_code(up(s[(j + p - r1) % p])), b = nuc_code(up(s[(j + p - r2) % p]));
        if (a != b) return a < b ? -1 : 1;
    }
    return 0;
}

// GS lanes cooperate on one locus (GS = 32: warp per locus; GS = 8: four loci per warp — most loci have only a
// handful of alleles, so a full warp per locus leaves 80% of the lanes idle).  `lane` is the lane within the group
// and every shuffle / vote / barrier below is restricted to the group's lanes.
constexpr int kMaxKeys = 128;     // k-mer keys kept per locus in shared memory (longer repeats recompute them)

template <int GS>
__global__ void __launch_bounds__(128) harmonize_kernel(HarmParams P) {
    __shared__ unsigned long long skeys[128 / GS][kMaxKeys];
    const int lane = threadIdx.x % GS;
    const unsigned gmask = (GS == 32) ? 0xffffffffu : (((1u << GS) - 1u) << ((threadIdx.x & 31) / GS * GS));
    const int64_t l = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / GS;
    if (l >= P.L) return;
    const int a0 = P.locus_off[l];
    const int A = P.locus_off[l + 1] - a0;
    const int period = P.period[l];
    const bool flanked = (P.vcftype == TRT_VCF_HIPSTR || P.vcftype == TRT_VCF_LONGTR);
    const int64_t ref_off = P.allele_off[a0];
    const int ref_n = (int)(P.allele_off[a0 + 1] - ref_off);
    int start_offset = 0, neg_end = 0;
    if (flanked) {
        start_offset = P.start[l] - P.pos[l];
        neg_end = (P.end[l] - P.pos[l]) + 1 - ref_n;  // tr_harmonizer.py:357-359
    }
    int flags = 0;
    if (flanked && !(start_offset == 0 && neg_end == 0)) flags |= TRT_HF_HAS_FULL;
    if (period <= 0) flags |= TRT_HF_BAD_PERIOD;
    const char* motif_in = P.motif_in ? P.motif_in + P.motif_off[l] : nullptr;

    // ---- 1. trim window + repeat-unit length per allele -------------------------------------
    for (int a = lane; a < A; a += GS) {
        const int64_t off = P.allele_off[a0 + a];
        const int n = (int)(P.allele_off[a0 + a + 1] - off);
        const double g = P.given_len[a0 + a];
        int s = 0, len = n;
        double rl;
        if (!isnan(g)) {
            // length-only allele (EH <STRn>, popSTR <n>): FabricateAllele utils.py:596-602
            rl = g;
            long long fab = (long long)floor(g) * (long long)(period > 0 ? period : 1);
            if (period > 0)
                while ((double)(fab + 1) / (double)period < g) fab++;
            s = 0;
            len = (int)fab;
        } else {
            if (flanked) {
                s = py_idx(start_offset, n);
                int e = (neg_end == 0) ? n : py_idx(neg_end, n);
                len = e > s ? e - s : 0;
            }
            rl = period > 0 ? (double)len / (double)period : nan("");
        }
        P.trim_off[a0 + a] = s;
        P.trim_len[a0 + a] = len;
        P.allele_len[a0 + a] = rl;
    }
    __syncwarp(gmask);

    auto make_str = [&](int a) -> AlleleStr {
        AlleleStr x;
        const double g = P.given_len[a0 + a];
        x.len = P.trim_len[a0 + a];
        if (!isnan(g)) {
            x.base = motif_in;
            x.mlen = period > 0 ? period : 1;
            if (!motif_in) x.len = 0;
        } else {
            x.base = P.seqs + P.allele_off[a0 + a] + P.trim_off[a0 + a];
            x.mlen = 0;
        }
        return x;
    };

    // ---- 2. equivalence classes and sort ranks (rank by counting; A is small) ------------------
    bool len_dups = false, seq_dups = false;
    for (int a = lane; a < A; a += GS) {
        const double la = P.allele_len[a0 + a];
        const AlleleStr sa = make_str(a);
        int lclass = a, sclass = a, lrank = 0, srank = 0;
        for (int b = 0; b < A; b++) {
            if (b == a) continue;
            const double lb = P.allele_len[a0 + b];
            if (lb == la) {
                if (b < lclass) lclass = b;
                if (b < a) lrank++;
            } else if (lb < la) {
                lrank++;
            }
            int c = str_cmp(make_str(b), sa);
            if (c == 0) {
                if (b < sclass) sclass = b;
                if (b < a) srank++;
            } else if (c < 0) {
                srank++;
            }
        }
        P.len_class[a0 + a] = lclass;
        P.seq_class[a0 + a] = sclass;
        P.len_order[a0 + lrank] = a;
        P.seq_order[a0 + srank] = a;
        len_dups |= (lclass != a);
        seq_dups |= (sclass != a);
    }
    if (__any_sync(gmask, len_dups)) flags |= TRT_HF_LEN_DUPS;
    if (__any_sync(gmask, seq_dups)) flags |= TRT_HF_SEQ_DUPS;

    // ---- 3. homopolymer run of the full REF (utils.py:340-360) ---------------------------------
    {
        // full REF when flanks exist (filters.py:205-208), else the (possibly fabricated) ref allele
        AlleleStr r;
        if (flanked) {
            r.base = P.seqs + ref_off;
            r.len = ref_n;
            r.mlen = 0;
        } else {
            r = make_str(0);
        }
        int best = 0;
        for (int i = lane; i < r.len; i += GS) {
            if (i == 0 || r.at(i) != r.at(i - 1)) {
                int j = i + 1;
                const char c = r.at(i);
                while (j < r.len && r.at(j) == c) j++;
                best = max(best, j - i);
            }
        }
#pragma unroll
        for (int o = GS / 2; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(gmask, best, o));
        if (lane == 0) P.hrun[l] = best;
    }

    // ---- 4. motif ---------------------------------------------------------------------------------
    if (period > 0) {
        char* mout = P.motif + P.motif_off[l];
        if (!flanked) {
            // RU / Motif supplied by the caller: pass through upper-cased
            for (int j = lane; j < period; j += GS) mout[j] = motif_in ? up(motif_in[j]) : 'N';
        } else {
            // InferRepeatSequence(ref_allele[start_offset:], PERIOD)  tr_harmonizer.py:397 — the
            // ALREADY trimmed REF is sliced by start_offset a second time (reference quirk).
            const int tn = P.trim_len[a0];
            const int s2 = py_idx(start_offset, tn);
            const char* seq = P.seqs + ref_off + P.trim_off[a0] + s2;
            const int n = tn - s2;
            if (period > n) {
                for (int j = lane; j < period; j += GS) mout[j] = 'N';
                flags |= TRT_HF_MOTIF_N;
            } else {
                const int K = n / period;
                // running count c_i of k-mer i among k-mers 0..i; the winner is the k-mer that first
                // reaches the final maximum count (see oracle/trh.py::infer_repeat_sequence)
                int best_c = 0, best_i = 0x7fffffff;
                const unsigned long long kmask = period >= 8 ? ~0ull : ((1ull << (8 * period)) - 1ull);
                // k-mers of <= 8 bases as upper-cased 64-bit keys (one unaligned 8-byte read each), computed once per
                // k-mer into the locus' shared-memory slot when they fit: the pair loop below then compares keys only
                const bool keyed = period <= 8 && K <= kMaxKeys;
                unsigned long long* keys = skeys[threadIdx.x / GS];
                if (keyed) {
                    for (int i = lane; i < K; i += GS) keys[i] = load8_up(seq + (size_t)i * period) & kmask;
                    __syncwarp(gmask);
                }
                for (int i = lane; i < K; i += GS) {
                    int c = 0;
                    if (keyed) {
                        const unsigned long long ki = keys[i];
                        for (int j = 0; j <= i; j++) c += (keys[j] == ki);
                    } else if (period <= 8) {
                        const unsigned long long ki = load8_up(seq + (size_t)i * period) & kmask;
                        for (int j = 0; j <= i; j++) c += ((load8_up(seq + (size_t)j * period) & kmask) == ki);
                    } else
                    for (int j = 0; j <= i; j++) {
                        bool eq = true;
                        for (int t = 0; t < period; t++)
                            if (up(seq[j * period + t]) != up(seq[i * period + t])) {
                                eq = false;
                                break;
                            }
                        c += eq;
                    }
                    if (c > best_c) {  // i ascending within a lane: first index reaching each count
                        best_c = c;
                        best_i = i;
                    }
                }
#pragma unroll
                for (int o = GS / 2; o > 0; o >>= 1) {
                    int oc = __shfl_xor_sync(gmask, best_c, o);
                    int oi = __shfl_xor_sync(gmask, best_i, o);
                    if (oc > best_c || (oc == best_c && oi < best_i)) {
                        best_c = oc;
                        best_i = oi;
                    }
                }
                const char* kmer = seq + (size_t)best_i * period;
                bool bad = false;
                for (int j = lane; j < period; j += GS) bad |= (nuc_code(up(kmer[j])) < 0);
                if (__any_sync(gmask, bad)) {
                    flags |= TRT_HF_MOTIF_NONACGT;  // GetCanonicalOneStrand would raise KeyError
                    for (int j = lane; j < period; j += GS) mout[j] = up(kmer[j]);
                } else {
                    int br = 0x7fffffff;  // best rotation seen by this lane
                    for (int r = lane; r < period; r += GS)
                        if (br == 0x7fffffff || rot_cmp(kmer, period, r, br) < 0) br = r;
#pragma unroll
                    for (int o = GS / 2; o > 0; o >>= 1) {
                        int orr = __shfl_xor_sync(gmask, br, o);
                        if (orr != 0x7fffffff && (br == 0x7fffffff || rot_cmp(kmer, period, orr, br) < 0 ||
                                                  (rot_cmp(kmer, period, orr, br) == 0 && orr < br)))
                            br = orr;
                    }
                    for (int j = lane; j < period; j += GS) mout[j] = up(kmer[(j + period - br) % period]);
                }
            }
        }
    }
    if (lane == 0) P.flags[l] = flags;
}

// packed length-genotype tensor: int16 [L][S][P] = rank of the haplotype's length class
__global__ void pack_kernel(const int16_t* __restrict__ gt, size_t pitch, int64_t L, int64_t S, int Pl,
                            const int32_t* __restrict__ locus_off, const int32_t* __restrict__ len_rank_of_allele,
                            int16_t* __restrict__ out) {
    for (int64_t l = blockIdx.y; l < L; l += gridDim.y) {
        const int a0 = locus_off[l];
        const int A = locus_off[l + 1] - a0;
        const int16_t* row = (const int16_t*)((const char*)gt + (size_t)l * pitch);
        int16_t* orow = out + (size_t)l * S * Pl;
        for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < S; s += (int64_t)gridDim.x * blockDim.x) {
            for (int h = 0; h < Pl; h++) {
                int a = row[s * (Pl + 1) + h];
                int16_t v = (int16_t)a;
                if (a >= 0) v = a < A ? (int16_t)len_rank_of_allele[a0 + a] : (int16_t)-1;
                orow[s * Pl + h] = v;
            }
        }
    }
}

// Vectorised pack (diploid, S % 4 == 0): a CTA moves 2048 calls at a time.  Loads are fully coalesced 16-byte
// vectors into shared memory; a thread then owns 8 consecutive calls (3 x LDS.128 at a 48-byte stride: conflict-free),
// maps the two alleles of each call through the locus' length-rank table (shared memory) and stages 32 bytes of
// output, which leave as coalesced 16-byte stores.  10 algorithmic bytes per call (6 read, 4 written).
constexpr int kPackThreads = 256;
constexpr int kPackCalls = kPackThreads * 8;          // 2048 calls = 12288 B in, 8192 B out per iteration
constexpr int kPackMaxRanks = 512;

__global__ void __launch_bounds__(kPackThreads) pack_vec_kernel(const int16_t* __restrict__ gt, size_t pitch, int64_t L, int64_t S,
                                                                const int32_t* __restrict__ locus_off,
                                                                const int32_t* __restrict__ len_rank_of_allele,
                                                                int16_t* __restrict__ out, int loci_per_block) {
    __shared__ uint4 sin[3 * kPackThreads];
    __shared__ uint4 sout[2 * kPackThreads];
    __shared__ int16_t srank[kPackMaxRanks];
    const int tid = threadIdx.x;
    const int64_t c0 = (int64_t)blockIdx.x * kPackCalls;          // first call (sample) of this CTA's slab
    const int64_t ncalls = min((int64_t)kPackCalls, S - c0);
    const int64_t in_vecs = (ncalls * 6 + 15) / 16;               // 16-byte vectors of the slab (row pitch covers the tail)
    const int64_t out_vecs = ncalls * 4 / 16;                     // S % 4 == 0 on this path
    const int64_t l_begin = (int64_t)blockIdx.y * loci_per_block, l_end = min(L, l_begin + loci_per_block);
    for (int64_t l = l_begin; l < l_end; l++) {
        const int a0 = locus_off[l];
        const int A = locus_off[l + 1] - a0;
        const bool ranks_in_smem = A <= kPackMaxRanks;
        __syncthreads();                                           // previous iteration's sout / srank readers are done
        if (ranks_in_smem)
            for (int a = tid; a < A; a += kPackThreads) srank[a] = (int16_t)len_rank_of_allele[a0 + a];
        const uint4* src = (const uint4*)((const char*)gt + (size_t)l * pitch + (size_t)c0 * 6);
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const int i = r * kPackThreads + tid;
            if (i < in_vecs) sin[i] = src[i];
        }
        __syncthreads();
        if ((int64_t)tid * 8 < ncalls) {
            const uint4 v0 = sin[3 * tid], v1 = sin[3 * tid + 1], v2 = sin[3 * tid + 2];
            const uint32_t w[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
            uint32_t o[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int k0 = 3 * j, k1 = k0 + 1;
                const int a = (k0 & 1) ? ((int)w[k0 >> 1] >> 16) : (int)(short)(w[k0 >> 1] & 0xffffu);
                const int b = (k1 & 1) ? ((int)w[k1 >> 1] >> 16) : (int)(short)(w[k1 >> 1] & 0xffffu);
                int ra = a, rb = b;                                // -1 / -2 sentinels pass through
                if (a >= 0) ra = a < A ? (ranks_in_smem ? (int)srank[a] : len_rank_of_allele[a0 + a]) : -1;
                if (b >= 0) rb = b < A ? (ranks_in_smem ? (int)srank[b] : len_rank_of_allele[a0 + b]) : -1;
                o[j] = (uint32_t)(uint16_t)(int16_t)ra | ((uint32_t)(uint16_t)(int16_t)rb << 16);
            }
            sout[2 * tid] = make_uint4(o[0], o[1], o[2], o[3]);
            sout[2 * tid + 1] = make_uint4(o[4], o[5], o[6], o[7]);
        }
        __syncthreads();
        uint4* dst = (uint4*)(out + ((size_t)l * S + c0) * 2);
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const int i = r * kPackThreads + tid;
            if (i < out_vecs) dst[i] = sout[i];
        }
    }
}

// rank of each allele's length class among the locus' distinct lengths (ascending)
__global__ void len_rank_kernel(const int32_t* __restrict__ locus_off, const int32_t* __restrict__ len_order,
                                const int32_t* __restrict__ len_class, int64_t L, int32_t* __restrict__ rank_out) {
    int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    const int a0 = locus_off[l], A = locus_off[l + 1] - a0;
    int rank = -1, prev_class = -1;
    for (int i = 0; i < A; i++) {
        int a = len_order[a0 + i];
        int c = len_class[a0 + a];
        if (c != prev_class) {
            rank++;
            prev_class = c;
        }
        rank_out[a0 + a] = rank;
    }
}

}  // namespace

extern "C" {

int trt_harmonize(trt_ctx* ctx) {
    if (!ctx || !ctx->block_open || !ctx->have_alleles)
        return trt_set_error(ctx, TRT_ESTATE, "trt_harmonize: call trt_block_begin and trt_block_set_alleles first");
    TRT_CUDA(cudaSetDevice(ctx->device));
    const int64_t L = ctx->L, nA = ctx->nA;
    // (motif offsets = exclusive prefix sum of max(period, 0): uploaded with the allele tables, trt_block_set_alleles)
    TRT_TRY(trt_ensure(ctx, ctx->motif, (size_t)ctx->motif_bytes + 16));
    TRT_TRY(trt_ensure(ctx, ctx->allele_len, (size_t)nA * 8));
    TRT_TRY(trt_ensure(ctx, ctx->trim_off, (size_t)nA * 4));
    TRT_TRY(trt_ensure(ctx, ctx->trim_len, (size_t)nA * 4));
    TRT_TRY(trt_ensure(ctx, ctx->len_class, (size_t)nA * 4));
    TRT_TRY(trt_ensure(ctx, ctx->seq_class, (size_t)nA * 4));
    TRT_TRY(trt_ensure(ctx, ctx->len_order, (size_t)nA * 4));
    TRT_TRY(trt_ensure(ctx, ctx->seq_order, (size_t)nA * 4));
    TRT_TRY(trt_ensure(ctx, ctx->hrun, (size_t)L * 4));
    TRT_TRY(trt_ensure(ctx, ctx->hflags, (size_t)L * 4));
    HarmParams P;
    P.seqs = (const char*)ctx->seqs.p;
    P.allele_off = (const int64_t*)ctx->allele_off.p;
    P.locus_off = (const int32_t*)ctx->locus_off.p;
    P.pos = (const int32_t*)ctx->pos.p;
    P.start = (const int32_t*)ctx->start.p;
    P.end = (const int32_t*)ctx->end.p;
    P.period = (const int32_t*)ctx->period.p;
    P.given_len = (const double*)ctx->given_len.p;
    P.motif_in = ctx->have_motif_in ? (const char*)ctx->motif_in.p : nullptr;
    P.motif_off = (const int64_t*)ctx->motif_off.p;
    P.L = L;
    P.vcftype = ctx->vcftype;
    P.allele_len = (double*)ctx->allele_len.p;
    P.trim_off = (int32_t*)ctx->trim_off.p;
    P.trim_len = (int32_t*)ctx->trim_len.p;
    P.len_class = (int32_t*)ctx->len_class.p;
    P.seq_class = (int32_t*)ctx->seq_class.p;
    P.len_order = (int32_t*)ctx->len_order.p;
    P.seq_order = (int32_t*)ctx->seq_order.p;
    P.hrun = (int32_t*)ctx->hrun.p;
    P.flags = (int32_t*)ctx->hflags.p;
    P.motif = (char*)ctx->motif.p;
    if (L > 0) {
        trt_timer_begin(ctx);
        // 8 lanes per locus unless the block has loci with many alleles (lanes stride over alleles / k-mers)
        if (ctx->maxA <= 16 && !getenv("TRT_HARMONIZE_WARP")) {
            const int64_t per = 128 / 8;
            harmonize_kernel<8><<<(unsigned)((L + per - 1) / per), 128, 0, ctx->stream>>>(P);
        } else {
            const int64_t per = 128 / 32;
            harmonize_kernel<32><<<(unsigned)((L + per - 1) / per), 128, 0, ctx->stream>>>(P);
        }
        TRT_KERNEL_CHECK();
        trt_timer_end_async(ctx);
    }
    // no host synchronisation: every consumer of the tables is a later operation on the context's stream
    ctx->harmonized = true;
    ctx->have_packed = false;
    return TRT_OK;
}

int trt_get_harmonized(trt_ctx* ctx, trt_harmonize_out* out) {
    if (!ctx || !ctx->harmonized) return trt_set_error(ctx, TRT_ESTATE, "trt_get_harmonized: call trt_harmonize first");
    if (!out) return trt_set_error(ctx, TRT_EINVAL, "trt_get_harmonized: out is NULL");
    const int64_t L = ctx->L, nA = ctx->nA;
#define D2H(dst, buf, bytes) \
    if ((dst) && (bytes)) TRT_CUDA(cudaMemcpyAsync((dst), (buf).p, (bytes), cudaMemcpyDeviceToHost, ctx->stream))
    D2H(out->allele_len, ctx->allele_len, (size_t)nA * 8);
    D2H(out->trim_off, ctx->trim_off, (size_t)nA * 4);
    D2H(out->trim_len, ctx->trim_len, (size_t)nA * 4);
    D2H(out->len_class, ctx->len_class, (size_t)nA * 4);
    D2H(out->seq_class, ctx->seq_class, (size_t)nA * 4);
    D2H(out->len_order, ctx->len_order, (size_t)nA * 4);
    D2H(out->seq_order, ctx->seq_order, (size_t)nA * 4);
    D2H(out->hrun, ctx->hrun, (size_t)L * 4);
    D2H(out->flags, ctx->hflags, (size_t)L * 4);
    D2H(out->motif, ctx->motif, (size_t)ctx->motif_bytes);
    D2H(out->motif_off, ctx->motif_off, (size_t)(L + 1) * 8);
#undef D2H
    TRT_CUDA(cudaStreamSynchronize(ctx->stream));
    return TRT_OK;
}

int trt_pack_length_genotypes(trt_ctx* ctx) {
    if (!ctx || !ctx->harmonized || !ctx->have_gt)
        return trt_set_error(ctx, TRT_ESTATE, "trt_pack_length_genotypes: needs GT and trt_harmonize");
    const int64_t L = ctx->L, S = ctx->S;
    TRT_TRY(trt_ensure(ctx, ctx->stat_i32, (size_t)ctx->nA * 4 + 16));
    TRT_TRY(trt_ensure(ctx, ctx->packed, (size_t)L * S * ctx->P * 2 + 16));
    if (L > 0 && S > 0) {
        trt_timer_begin(ctx);
        len_rank_kernel<<<(unsigned)((L + 127) / 128), 128, 0, ctx->stream>>>(
            (const int32_t*)ctx->locus_off.p, (const int32_t*)ctx->len_order.p, (const int32_t*)ctx->len_class.p, L,
            (int32_t*)ctx->stat_i32.p);
        TRT_KERNEL_CHECK();
        if (ctx->P == 2 && S % 4 == 0 && (ctx->gt_active_pitch % 16) == 0 && ((uintptr_t)ctx->d_gt_active % 16) == 0 &&
            !getenv("TRT_PACK_SCALAR")) {
            const int64_t slabs = (S + kPackCalls - 1) / kPackCalls;
            // enough CTAs to fill the machine several times; each walks a chunk of loci over its sample slab
            int64_t per = std::max<int64_t>(1, (L * slabs) / ((int64_t)ctx->sm_count * 64));
            per = std::min<int64_t>(per, 64);
            dim3 grid((unsigned)slabs, (unsigned)((L + per - 1) / per));
            pack_vec_kernel<<<grid, kPackThreads, 0, ctx->stream>>>(ctx->d_gt_active, ctx->gt_active_pitch, L, S,
                                                                     (const int32_t*)ctx->locus_off.p,
                                                                     (const int32_t*)ctx->stat_i32.p, (int16_t*)ctx->packed.p, (int)per);
        } else {
            dim3 grid((unsigned)std::min<int64_t>((S + 255) / 256, 64), (unsigned)std::min<int64_t>(L, 32768));
            pack_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->d_gt_active, ctx->gt_active_pitch, L, S, ctx->P,
                                                       (const int32_t*)ctx->locus_off.p, (const int32_t*)ctx->stat_i32.p,
                                                       (int16_t*)ctx->packed.p);
        }
        TRT_KERNEL_CHECK();
        trt_timer_end(ctx);
    }
    ctx->have_packed = true;
    return TRT_OK;
}

int trt_get_packed_gt(trt_ctx* ctx, int16_t* out_host) {
    if (!ctx || !ctx->have_packed) return trt_set_error(ctx, TRT_ESTATE, "trt_get_packed_gt: call trt_pack_length_genotypes first");
    size_t bytes = (size_t)ctx->L * ctx->S * ctx->P * 2;
    if (bytes) TRT_CUDA(cudaMemcpyAsync(out_host, ctx->packed.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    TRT_CUDA(cudaStreamSynchronize(ctx->stream));
    return TRT_OK;
}

}  // extern "C"
