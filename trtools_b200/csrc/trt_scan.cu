// K2a — the GT scan: everything per-locus statistics need from the sample axis, in ONE pass over the
// native cyvcf2 GT rows (int16 [S][3] per locus, 6 algorithmic bytes per call), no intermediate tensor.
//
// Output per (group, locus):  ac[a] = allele counts keyed by allele INDEX (TRRecord.GetAlleleCounts,
// trtools/utils/tr_harmonizer.py:1420-1499: -1/-2 dropped, partial calls count) and the counters
// TRT_LC_* (fully-called samples = sum of TRRecord.GetGenotypeCounts :1326-1418, non-strict called
// samples :864-897, fully-called samples carrying a ploidy pad, and homozygotes under the index /
// length / sequence equivalences, i.e. "first two sorted haplotypes equal" of
// utils.GetHardyWeinbergBinomialTest trtools/utils/utils.py:328-333).
//
// Three tiers, all HBM-bandwidth bound by design (no data reuse; DRAM traffic = algorithmic bytes):
//   scan_pairs_kernel<NG, CELL>  (diploid, A <= 14 alleles): persistent CTAs, one locus at a time (several short loci per
//       CTA pass when S is small).  576 threads: a producer warp streams the row through an up-to-8-stage shared-memory
//       ring of 4096-call chunks with 1-D bulk TMA copies (cp.async.bulk + mbarrier tx bytes); 16 consumer warps read
//       48 B = 8 calls per thread (conflict-free LDS.128), turn each call into a bin of the unordered genotype-pair
//       table with packed 16-bit arithmetic (VIADDMNMX.U16x2 clamp + IDP.2A index) and do ONE thread-private
//       shared-memory increment per call — 8-bit cells [(A+2)(A+3)/2][512] folded every <= 31 chunks (16-bit cells as the
//       fallback); an epilogue warp turns the previous locus' reduced pair table into allele counts and the TRT_LC_*
//       counters while the consumers already stream the next locus (named barriers, no CTA-wide barrier in the loop).
//       Every statistic is derived from the pair table, so there is no per-call flag logic at all.
//       NG > 0: up to three sample groups (statSTR --samples f1,f2,..) are counted in the SAME pass — one membership
//       byte per sample (bit per group), a table per group.
//   scan_wide_kernel   (diploid, A <= 96): same TMA ring, thread-private per-haplotype 16-bit counters.
//   scan_generic_kernel: warp per locus; any ploidy, any allele count, tiny sample counts.
#include <stdlib.h>

#include <algorithm>

#include "trt_internal.cuh"
#include "trt_scan.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------
// generic: one warp per locus
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) scan_generic_kernel(ScanParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t l = warp; l < p.L; l += nwarps) {
        const int a0 = p.locus_off[l];
        const int A = p.locus_off[l + 1] - a0;
        if (p.fast_enabled && scan_tier(A) != TIER_GENERIC) continue;
        const int16_t* row = (const int16_t*)((const char*)p.gt + (size_t)l * p.pitch);
        const int P = p.P;
        long long n_full = 0, n_non = 0, n_pad = 0, h_idx = 0, h_len = 0, h_seq = 0, n_bad = 0;
        for (int64_t s = lane; s < p.S; s += 32) {
            if (p.mask && !p.mask[s]) continue;
            const int16_t* g = row + s * (P + 1);
            bool any_m1 = false, any_called = false, any_pad = false;
            // two smallest keys under each relation (pads sort first: key -2)
            int i1 = INT_MAX, i2 = INT_MAX, l1 = INT_MAX, l2 = INT_MAX, q1 = INT_MAX, q2 = INT_MAX;
            for (int h = 0; h < P; h++) {
                const int a = g[h];
                int ki, kl, kq;
                if (a == -1) {
                    any_m1 = true;
                    continue;
                } else if (a == -2) {
                    any_pad = true;
                    ki = kl = kq = -2;
                } else if (a >= 0 && a < A) {
                    any_called = true;
                    atomicAdd(&p.ac[a0 + a], 1);
                    ki = a;
                    kl = p.len_rank[a0 + a];
                    kq = p.seq_rank[a0 + a];
                } else {
                    n_bad++;
                    any_m1 = true;
                    continue;
                }
                if (ki < i1) { i2 = i1; i1 = ki; } else if (ki < i2) i2 = ki;
                if (kl < l1) { l2 = l1; l1 = kl; } else if (kl < l2) l2 = kl;
                if (kq < q1) { q2 = q1; q1 = kq; } else if (kq < q2) q2 = kq;
            }
            if (any_called) n_non++;
            if (any_m1 && any_called && p.ac_part) {
                for (int h = 0; h < P; h++) {
                    const int a = g[h];
                    if (a >= 0 && a < A) atomicAdd(&p.ac_part[a0 + a], 1);
                }
            }
            if (!any_m1) {
                n_full++;
                if (any_pad) n_pad++;
                if (P >= 2) {
                    h_idx += (i1 == i2);
                    h_len += (l1 == l2);
                    h_seq += (q1 == q2);
                }
            }
        }
        n_full = warp_sum_ll(n_full); n_non = warp_sum_ll(n_non); n_pad = warp_sum_ll(n_pad);
        h_idx = warp_sum_ll(h_idx); h_len = warp_sum_ll(h_len); h_seq = warp_sum_ll(h_seq);
        n_bad = warp_sum_ll(n_bad);
        if (lane == 0) {
            long long* o = p.lc + l * TRT_LC_N;
            o[TRT_LC_NFULL] = n_full; o[TRT_LC_NNONSTRICT] = n_non; o[TRT_LC_NPAD] = n_pad;
            o[TRT_LC_HOM_IDX] = h_idx; o[TRT_LC_HOM_LEN] = h_len; o[TRT_LC_HOM_SEQ] = h_seq;
            o[6] = n_bad; o[7] = 0;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// pair-table tiers
// ---------------------------------------------------------------------------------------------------
constexpr int kPT = 512;                       // consumer threads (16 warps: 4 per scheduler hide the RMW latency)
constexpr int kPWarps = kPT / 32;
constexpr int kPThreads = kPT + 64;            // + producer warp + epilogue warp
constexpr int kPChunkBytes = kPT * 48;         // 24576 B = 4096 calls; thread t owns bytes [48t, 48t+48) = 8 calls
constexpr int kPChunkCalls = kPT * 8;
constexpr int kPMaxStages = 8;
constexpr int kPMinStages = 3;
constexpr int kPMaxDigits = kPairsMaxAlleles + 3;
constexpr int kPMaxBins = kPMaxDigits * (kPMaxDigits + 1) / 2;   // unordered digit pairs of the largest tier
constexpr int kPMaxGroups = 8;                 // fold groups: warp pairs (16-bit cells) or warp quads (8-bit cells)
constexpr int kPMaxChunks8 = 31;               // 8-bit cells hold 8 calls x 31 chunks = 248 <= 255 per (thread, bin)

// Thread-private count cells, one table row per bin.  A row is 512 cells; within one warp the 32 lanes always touch 32
// different banks whatever bins they address, because the warps that share a 32-bit word own different BYTES of it:
//   16-bit cells (row = 1 KB): halfword  64*(t/64) + 2*(t%32) + ((t/32)&1)      — a warp PAIR shares words
//    8-bit cells (row = 512 B): byte     128*(w/4) + 4*(t%32) + (w%4), w = t/32 — a warp QUAD shares words
// 8-bit cells halve the table (more ring stages fit: the ring depth is what bounds the stream) and are used whenever a
// locus has at most 31 chunks (S <= 126 976).  The warps sharing words form a fold group: they reduce and zero their
// own 128-byte slice of every row behind a group-local named barrier.
template <typename CELL> struct CellTraits;
template <> struct CellTraits<uint16_t> {
    static constexpr int kGroupWarps = 2;
    __device__ static __forceinline__ int cell_index(int warp, int lane) { return 64 * (warp >> 1) + 2 * lane + (warp & 1); }
    __device__ static __forceinline__ unsigned sum4(const uint4& x) {
        return (x.x & 0xffffu) + (x.x >> 16) + (x.y & 0xffffu) + (x.y >> 16) + (x.z & 0xffffu) + (x.z >> 16) + (x.w & 0xffffu) + (x.w >> 16);
    }
};
template <> struct CellTraits<uint8_t> {
    static constexpr int kGroupWarps = 4;
    __device__ static __forceinline__ int cell_index(int warp, int lane) { return 128 * (warp >> 2) + 4 * lane + (warp & 3); }
    __device__ static __forceinline__ unsigned sum4(const uint4& x) {
        return (unsigned)__dp4a(x.x, 0x01010101u, __dp4a(x.y, 0x01010101u, __dp4a(x.z, 0x01010101u, __dp4a(x.w, 0x01010101u, 0u))));
    }
};

constexpr int kPMaxSampleGroups = 3;           // sample groups (statSTR --samples) counted in ONE pass over the GT rows
template <int NGH>
struct __align__(16) PairHeaderT {
    uint64_t full[kPMaxStages];
    uint64_t empty[kPMaxStages];
    uint64_t part_free[2];                                   // the epilogue warp has consumed partial[parity]
    unsigned int partial[2][NGH][kPMaxGroups][kPMaxBins];    // per-fold-group sums of each sample group, double buffered by locus parity
    unsigned int T[kPMaxBins];                               // the epilogue warp's CTA-wide pair table
};

// bin of an UNORDERED digit pair (genotypes are unphased for every statistic): tri(hi, lo), lo <= hi
__device__ __forceinline__ unsigned tri(unsigned hi, unsigned lo) { return ((hi * (hi + 1u)) >> 1) + lo; }

// two thread-private increments with both loads in flight; equal bins both store x + 2
template <typename CELL>
__device__ __forceinline__ void bump2(CELL* my, unsigned i0, unsigned i1) {
    CELL* p0 = my + (size_t)i0 * kPT;
    CELL* p1 = my + (size_t)i1 * kPT;
    const unsigned x0 = *p0, x1 = *p1;
    const unsigned e = (i0 == i1) ? 2u : 1u;
    *p0 = (CELL)(x0 + e);
    *p1 = (CELL)(x1 + e);
}

// metadata of the i-th locus of this launch's tier list (or L past the end)
__device__ __forceinline__ int64_t list_locus(const ScanParams& p, int i, int& A, int& a0) {
    if (i >= p.n_list) return p.L;
    const int64_t l = p.list[i];
    a0 = p.locus_off[l];
    A = p.locus_off[l + 1] - a0;
    return l;
}

__device__ __forceinline__ void named_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// Warp roles: 16 consumer warps, one producer warp (one elected lane feeds the ring), one epilogue warp (turns the
// fold groups' partial sums of a finished locus into every per-locus output while the consumers are already streaming
// the next locus: no consumer ever leaves the stream, so the ring never loses depth to a lagging warp).
// A digit of a haplotype: pad(-2) -> 0, no-call(-1) -> 1, allele a -> a+2, anything else -> D-1 ("bad").  Both digits of
// a call come from ONE add-and-min on the packed int16 pair (VIADDMNMX.U16x2) and the bin from ONE dot product
// (IDP.2A: d0*D + d1) while D^2 <= kSquareRows rows; wider loci use UNORDERED pairs (D(D+1)/2 rows).
// NG = 0: every sample counts.  NG = 1..3: that many SAMPLE GROUPS (statSTR --samples, associaTR's design membership)
// are counted in this one pass: p.gbits holds one byte per sample whose bit (p.group0 + g) says "in group g", every
// group has its own table, and a call bumps the tables of the groups its sample belongs to — the GT rows are read once
// however many groups there are (the reference re-derives its counts per group, statSTR.py:520-542).
template <int NG, typename CELL>
__global__ void __launch_bounds__(kPThreads, 1) scan_pairs_kernel(ScanParams p, int tier, int max_rows, int stages) {
    constexpr int NGH = NG > 0 ? NG : 1;
    typedef PairHeaderT<NGH> PairHeader;
    using CT = CellTraits<CELL>;
    constexpr int kGW = CT::kGroupWarps;                  // warps per fold group
    constexpr int kGT = kGW * 32;                         // threads per fold group
    constexpr int kGroups = kPWarps / kGW;
    constexpr int kRowWords = kPT * (int)sizeof(CELL) / 4;    // 32-bit words per table row
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* ring = smem;
    constexpr size_t stage_bytes = kPChunkBytes;
    PairHeader* hdr = (PairHeader*)(smem + (size_t)stages * stage_bytes);
    CELL* table = (CELL*)(smem + (size_t)stages * stage_bytes + sizeof(PairHeader));   // [max_rows + 1][512]

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const size_t row_bytes = (size_t)p.S * 6;
    const size_t copy_bytes = (row_bytes + 15) & ~size_t(15);
    const int nchunks = (int)((copy_bytes + kPChunkBytes - 1) / kPChunkBytes);
    const int nfull = (int)(p.S / kPChunkCalls);   // chunks whose calls are all real samples
    const unsigned trash = (unsigned)max_rows;     // extra row: calls beyond S in the last chunk land here

    if (tid == 0) {
        for (int s = 0; s < stages; s++) {
            mbar_init(&hdr->full[s], 1);
            mbar_init(&hdr->empty[s], kPWarps);
        }
        mbar_init(&hdr->part_free[0], 1);
        mbar_init(&hdr->part_free[1], 1);
        mbar_fence_init();
    }
    {
        const int words = NGH * (max_rows + 1) * kRowWords;
        for (int i = tid; i < words; i += kPThreads) ((uint32_t*)table)[i] = 0u;
    }
    __syncthreads();

    if (warp == kPWarps) {
        // ===== producer warp: one elected lane feeds the ring, running ahead across loci =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int i = blockIdx.x; i < p.n_list; i += gridDim.x) {
                const int64_t l = p.list[i];
                const char* src = (const char*)p.gt + (size_t)l * p.pitch;
                for (int st = 0; st < nchunks; st++) {
                    mbar_wait(&hdr->empty[stage], phase ^ 1u);
                    const size_t off = (size_t)st * stage_bytes;
                    const uint32_t bytes = (uint32_t)min(stage_bytes, copy_bytes - off);
                    mbar_arrive_expect_tx(&hdr->full[stage], bytes);
                    tma_load_1d(ring + (size_t)stage * stage_bytes, src + off, bytes, &hdr->full[stage]);
                    if (++stage == stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
        return;
    }

    if (warp == kPWarps + 1) {
        // ===== epilogue warp: CTA-wide table of UNORDERED pairs, then every per-locus output =====
        int parity = 0;
        int A, a0;
        for (int li = blockIdx.x;; li += gridDim.x) {
            const int64_t l = list_locus(p, li, A, a0);
            if (l >= p.L) break;
            const unsigned D = (unsigned)A + 3u, Dm1 = D - 1u;
            const bool sq = pairs_square(A);
            // class of digit `lane` (pad digit: -2; no-call / bad: -1), exchanged by shuffles below
            int cl = -1, cq = -1;
            if (lane == 0) cl = cq = -2;
            else if (lane >= 2 && lane < A + 2) {
                cl = p.len_class[a0 + lane - 2];
                cq = p.seq_class[a0 + lane - 2];
            }
            named_sync(1 + parity, kPT + 32);            // the fold groups' partials of this locus are complete
            const int nb = (int)(D * (D + 1u) / 2u);
            unsigned int* T = hdr->T;
            for (int sg = 0; sg < NGH; sg++) {
                long long n_full = 0, n_non = 0, n_pad = 0, h_idx = 0, h_len = 0, h_seq = 0, n_bad = 0;
                for (int b0 = 0; b0 < nb; b0 += 32) {
                    const int b = min(b0 + lane, nb - 1);
                    unsigned hi = (unsigned)((sqrtf(8.0f * (float)b + 1.0f) - 1.0f) * 0.5f);
                    while (tri(hi + 1u, 0u) <= (unsigned)b) hi++;
                    while (tri(hi, 0u) > (unsigned)b) hi--;
                    const unsigned lo = (unsigned)b - tri(hi, 0u);
                    unsigned t = 0;
#pragma unroll
                    for (int g = 0; g < kGroups; g++) {
                        const unsigned int* pp = hdr->partial[parity][sg][g];
                        if (sq) t += pp[hi * D + lo] + ((lo != hi) ? pp[lo * D + hi] : 0u);
                        else t += pp[b];
                    }
                    if (b0 + lane < nb) T[b] = t;
                    const int cl_lo = __shfl_sync(0xffffffffu, cl, lo), cl_hi = __shfl_sync(0xffffffffu, cl, hi);
                    const int cq_lo = __shfl_sync(0xffffffffu, cq, lo), cq_hi = __shfl_sync(0xffffffffu, cq, hi);
                    const long long n = (b0 + lane < nb) ? (long long)t : 0;
                    const bool bad = (hi == Dm1);                        // lo <= hi
                    const bool m1 = (lo == 1u) | (hi == 1u) | bad;
                    const bool vlo = (lo >= 2u) & (lo < Dm1), vhi = (hi >= 2u) & (hi < Dm1);
                    if (bad) n_bad += n;
                    if (vlo | vhi) n_non += n;
                    if (!m1) {
                        n_full += n;
                        if (lo == 0u) n_pad += n;
                        if (lo == hi) h_idx += n;
                        if (cl_lo == cl_hi) h_len += n;
                        if (cq_lo == cq_hi) h_seq += n;
                    }
                }
                __syncwarp();
                if (sg == NGH - 1 && lane == 0) mbar_arrive(&hdr->part_free[parity]);   // partial[parity] may be refilled (locus + 2)
                n_full = warp_sum_ll(n_full); n_non = warp_sum_ll(n_non); n_pad = warp_sum_ll(n_pad);
                h_idx = warp_sum_ll(h_idx); h_len = warp_sum_ll(h_len); h_seq = warp_sum_ll(h_seq);
                n_bad = warp_sum_ll(n_bad);
                const size_t og = (size_t)(p.group0 + sg);
                if (lane == 0) {
                    long long* o = p.lc + og * p.lc_stride + l * TRT_LC_N;
                    o[TRT_LC_NFULL] = n_full; o[TRT_LC_NNONSTRICT] = n_non; o[TRT_LC_NPAD] = n_pad;
                    o[TRT_LC_HOM_IDX] = h_idx; o[TRT_LC_HOM_LEN] = h_len; o[TRT_LC_HOM_SEQ] = h_seq;
                    o[6] = n_bad; o[7] = 0;
                }
                for (int a = lane; a < A; a += 32) {
                    const unsigned d = (unsigned)a + 2u;
                    unsigned cnt = 2u * T[tri(d, d)];
                    for (unsigned e = 0; e < d; e++) cnt += T[tri(d, e)];
                    for (unsigned e = d + 1u; e < D; e++) cnt += T[tri(e, d)];
                    p.ac[og * p.ac_stride + a0 + a] = (int)cnt;
                    if (p.ac_part) p.ac_part[og * p.ac_stride + a0 + a] = (int)(T[tri(d, 1u)] + T[tri(Dm1, d)]);   // partner '.' or invalid
                }
                __syncwarp();
            }
            __syncwarp();
            parity ^= 1;
        }
        return;
    }

    // ===== consumers =====
    const int group = warp / kGW, gwarp = warp % kGW;
    CELL* my = table + CT::cell_index(warp, lane);
    const size_t gstride = (size_t)(max_rows + 1) * kPT;      // cells per sample group's table
    int stage = 0;
    uint32_t phase = 0;
    int parity = 0;
    unsigned n_locus = 0;                 // loci this CTA has finished (uses of partial[] = n_locus >> 1 per parity)
    int A, a0, A_next = 0, a0_next = 0;
    int li = blockIdx.x;
    int64_t l = list_locus(p, li, A, a0);
    const int s_rel0 = tid * 8;           // first sample of this thread within a chunk
    while (l < p.L) {
        // metadata of the following locus is fetched now so its latency hides under this locus' stream
        li += gridDim.x;
        const int64_t l_next = list_locus(p, li, A_next, a0_next);
        const unsigned D = (unsigned)A + 3u, Dm1 = D - 1u;
        const bool sq = pairs_square(A);
        const unsigned dm2 = Dm1 | (Dm1 << 16);
        const int dot = (int)(D | (1u << 8));           // IDP.2A: lo16 * byte0 + hi16 * byte1 = d0 * D + d1

        for (int c = 0; c < nchunks; c++) {
            mbar_wait(&hdr->full[stage], phase);
            if (p.stream_only) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&hdr->empty[stage]);
                if (++stage == stages) { stage = 0; phase ^= 1u; }
                continue;
            }
            const uint4* sp = (const uint4*)(ring + (size_t)stage * stage_bytes + (size_t)tid * 48);
            const uint4 v0 = sp[0], v1 = sp[1], v2 = sp[2];
            const int stage_used = stage;
            if (++stage == stages) { stage = 0; phase ^= 1u; }
            // packed (a, b) int16 pairs of the 8 calls: even calls sit in one word, odd calls straddle two
            const uint32_t pr[8] = {v0.x, __byte_perm(v0.y, v0.z, 0x5432), v0.w, __byte_perm(v1.x, v1.y, 0x5432),
                                    v1.z, __byte_perm(v1.w, v2.x, 0x5432), v2.y, __byte_perm(v2.z, v2.w, 0x5432)};
            unsigned idx[8], dg[8];
#pragma unroll
            for (int j = 0; j < 8; j++) dg[j] = __vminu2(__vadd2(pr[j], 0x00020002u), dm2);     // both digits: one VIADDMNMX.U16x2
            if (sq) {
#pragma unroll
                for (int j = 0; j < 8; j++) idx[j] = (unsigned)__dp2a_lo((int)dg[j], dot, 0);
            } else {
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const unsigned d0 = dg[j] & 0xffffu, d1 = dg[j] >> 16;
                    idx[j] = tri(max(d0, d1), min(d0, d1));
                }
            }
            {
                // Release the ring slot only after this warp's loads have RETURNED: the predicate below consumes every
                // loaded word (it is always true, bins are < 2^16, but the compiler cannot prove it), so the arrive cannot
                // issue while a generic-proxy read of the slot is still in flight — the refill is an async-proxy (TMA)
                // write, which is not ordered behind such a read.
                const unsigned any = idx[0] | idx[1] | idx[2] | idx[3] | idx[4] | idx[5] | idx[6] | idx[7];
                const bool returned = __all_sync(0xffffffffu, any != 0xffffffffu);
                if (lane == 0 && returned) mbar_arrive(&hdr->empty[stage_used]);
            }
            if (c >= nfull) {
                const int left = (int)min((int64_t)kPChunkCalls, p.S - (int64_t)c * kPChunkCalls) - s_rel0;   // live calls of this thread
#pragma unroll
                for (int j = 0; j < 8; j++) idx[j] = (j < left) ? idx[j] : trash;
            }
            if (NG == 0) {
#pragma unroll
                for (int j = 0; j < 8; j += 2) bump2<CELL>(my, idx[j], idx[j + 1]);
            } else {
                // membership bytes of this thread's 8 samples (zero beyond S), one 64-bit load from the L1/L2-resident array
                const uint64_t mb = *reinterpret_cast<const uint64_t*>(p.gbits + (size_t)c * kPChunkCalls + s_rel0) >> p.group0;
#pragma unroll
                for (int g = 0; g < NG; g++) {
                    CELL* mine = my + (size_t)g * gstride;
                    unsigned ig[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) ig[j] = ((mb >> (8 * j + g)) & 1ull) ? idx[j] : trash;
#pragma unroll
                    for (int j = 0; j < 8; j += 2) bump2<CELL>(mine, ig[j], ig[j + 1]);
                }
            }
        }

        // ---- each fold GROUP reduces and zeroes its own 128-byte slice of every row (group-local barrier only) ------
        const int nrows = sq ? (int)(D * D) : (int)(D * (D + 1u) / 2u);
        named_sync(3 + group, kGT);
        // partial[parity] was last used two loci ago: wait until the epilogue warp has consumed it
        if (n_locus >= 2u) mbar_wait(&hdr->part_free[parity], ((n_locus >> 1) - 1u) & 1u);
        for (int sg = 0; sg < NGH; sg++) {
            unsigned int* part = hdr->partial[parity][sg][group];
            uint32_t* tab = (uint32_t*)(table + (size_t)sg * gstride);
            for (int b = gwarp * 32 + lane; b < nrows; b += kGT) {
                // this group's cells of row b = 128 B = 8 x 16 B; lanes own different rows (same bank offset), so each
                // rotates its 16-byte slot: the 8 lanes of a quarter-warp hit 8 different bank groups (conflict-free)
                uint4* rowp = (uint4*)(tab + (size_t)b * kRowWords + group * 32);
                unsigned sum = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int kk = (k + lane) & 7;
                    const uint4 x = rowp[kk];
                    rowp[kk] = make_uint4(0u, 0u, 0u, 0u);
                    sum += CT::sum4(x);
                }
                part[b] = sum;
            }
        }
        // the words just folded/zeroed hold every group warp's cells: none may start the next locus earlier
        named_sync(3 + group, kGT);
        named_arrive(1 + parity, kPT + 32);              // hand the partials to the epilogue warp and move on
        parity ^= 1;
        n_locus++;
        l = l_next;
        A = A_next;
        a0 = a0_next;
    }
}

// ---------------------------------------------------------------------------------------------------
// wide-allele tier: thread-private per-haplotype 16-bit counters
// ---------------------------------------------------------------------------------------------------
constexpr int kWT = 512;
constexpr int kWWarps = kWT / 32;
constexpr int kWThreads = kWT + 32;
constexpr int kWChunkBytes = kWT * 48;
constexpr int kWChunkCalls = kWT * 8;
constexpr int kWStages = 4;

struct __align__(16) WideHeader {
    uint64_t full[kWStages];
    uint64_t empty[kWStages];
    long long warp_part[kWWarps][8];
    uint32_t cls[kWideMaxAlleles];   // (len_class << 16) | seq_class
};

__device__ __forceinline__ void wbar() { asm volatile("bar.sync 1, %0;" ::"n"(kWT) : "memory"); }

template <bool MASKED>
__global__ void __launch_bounds__(kWThreads, 1) scan_wide_kernel(ScanParams p, int max_alleles_smem) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* ring = smem;
    WideHeader* hdr = (WideHeader*)(smem + kWStages * kWChunkBytes);
    uint16_t* cnt = (uint16_t*)(smem + kWStages * kWChunkBytes + sizeof(WideHeader));   // [A][512]

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const size_t row_bytes = (size_t)p.S * 6;
    const size_t copy_bytes = (row_bytes + 15) & ~size_t(15);
    const int nchunks = (int)((copy_bytes + kWChunkBytes - 1) / kWChunkBytes);

    if (tid == 0) {
        for (int s = 0; s < kWStages; s++) {
            mbar_init(&hdr->full[s], 1);
            mbar_init(&hdr->empty[s], kWWarps);
        }
        mbar_fence_init();
    }
    for (int i = tid; i < max_alleles_smem * kWT / 2; i += kWThreads) ((uint32_t*)cnt)[i] = 0u;
    __syncthreads();

    if (warp == kWWarps) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int i = blockIdx.x; i < p.n_list; i += gridDim.x) {
                const int64_t l = p.list[i];
                const char* src = (const char*)p.gt + (size_t)l * p.pitch;
                for (int c = 0; c < nchunks; c++, it++) {
                    const int stage = it % kWStages;
                    const uint32_t phase = (it / kWStages) & 1u;
                    mbar_wait(&hdr->empty[stage], phase ^ 1u);
                    const size_t off = (size_t)c * kWChunkBytes;
                    const uint32_t bytes = (uint32_t)min((size_t)kWChunkBytes, copy_bytes - off);
                    mbar_arrive_expect_tx(&hdr->full[stage], bytes);
                    tma_load_1d(ring + (size_t)stage * kWChunkBytes, src + off, bytes, &hdr->full[stage]);
                }
            }
        }
        return;
    }

    uint16_t* my = cnt + tid;
    uint32_t it = 0;
    for (int i = blockIdx.x; i < p.n_list; i += gridDim.x) {
        const int64_t l = p.list[i];
        const int a0 = p.locus_off[l];
        const int A = p.locus_off[l + 1] - a0;
        for (int a = tid; a < A; a += kWT)
            hdr->cls[a] = ((uint32_t)p.len_class[a0 + a] << 16) | (uint32_t)p.seq_class[a0 + a];
        wbar();
        int n_full = 0, n_non = 0, n_pad = 0, h_idx = 0, h_len = 0, h_seq = 0, n_bad = 0;
        for (int c = 0; c < nchunks; c++, it++) {
            const int stage = it % kWStages;
            const uint32_t phase = (it / kWStages) & 1u;
            mbar_wait(&hdr->full[stage], phase);
            const uint4* src = (const uint4*)(ring + (size_t)stage * kWChunkBytes + (size_t)tid * 48);
            const uint4 v0 = src[0], v1 = src[1], v2 = src[2];
            const uint32_t w[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
            const int64_t s_base = (int64_t)c * kWChunkCalls + (int64_t)tid * 8;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int k0 = 3 * j, k1 = 3 * j + 1;
                const int a = (k0 & 1) ? ((int)w[k0 >> 1] >> 16) : (int)(short)(w[k0 >> 1] & 0xffffu);
                const int b = (k1 & 1) ? ((int)w[k1 >> 1] >> 16) : (int)(short)(w[k1 >> 1] & 0xffffu);
                bool live = (s_base + j < p.S);
                if (MASKED) live = live && p.mask[live ? s_base + j : 0] != 0;
                if (!live) continue;
                const bool va = (unsigned)a < (unsigned)A, vb = (unsigned)b < (unsigned)A;
                if (va) my[a * kWT] += 1;
                if (vb) my[b * kWT] += 1;
                const bool bad = (a < -2) | (b < -2) | (a >= A) | (b >= A);
                const bool m1 = (a == -1) | (b == -1) | bad;
                n_bad += bad;
                n_non += (va | vb);
                if (m1 && p.ac_part) {
                    if (va) atomicAdd(&p.ac_part[a0 + a], 1);
                    if (vb) atomicAdd(&p.ac_part[a0 + b], 1);
                }
                if (!m1) {
                    n_full++;
                    n_pad += ((a == -2) | (b == -2));
                    h_idx += (a == b);
                    const uint32_t ca = va ? hdr->cls[a] : 0xfffe0000u, cb = vb ? hdr->cls[b] : 0xfffd0001u;
                    h_len += ((ca >> 16) == (cb >> 16)) | (a == b);
                    h_seq += ((ca & 0xffffu) == (cb & 0xffffu)) | (a == b);
                }
            }
            // release the slot after the chunk has been consumed (the counter updates above depend on every loaded
            // word), never while a read of the slot may still be in flight: see scan_pairs_kernel
            {
                const unsigned any = w[0] | w[1] | w[2] | w[3] | w[4] | w[5] | w[6] | w[7] | w[8] | w[9] | w[10] | w[11];
                const bool returned = __any_sync(0xffffffffu, any != 0x7fff7fffu);   // true unless ALL 32 lanes hold the constant
                if (lane == 0 && returned) mbar_arrive(&hdr->empty[stage]);
            }
        }
        n_full = warp_sum(n_full); n_non = warp_sum(n_non); n_pad = warp_sum(n_pad);
        h_idx = warp_sum(h_idx); h_len = warp_sum(h_len); h_seq = warp_sum(h_seq); n_bad = warp_sum(n_bad);
        if (lane == 0) {
            long long* wp = hdr->warp_part[warp];
            wp[0] = n_full; wp[1] = n_non; wp[2] = n_pad; wp[3] = h_idx; wp[4] = h_len; wp[5] = h_seq; wp[6] = n_bad;
        }
        wbar();
        for (int a = warp; a < A; a += kWWarps) {
            uint32_t* rowp = (uint32_t*)(cnt + (size_t)a * kWT);
            int sum = 0;
#pragma unroll
            for (int k = 0; k < kWT / 64; k++) {
                const uint32_t x = rowp[lane + 32 * k];
                rowp[lane + 32 * k] = 0u;
                sum += (int)(x & 0xffffu) + (int)(x >> 16);
            }
            sum = warp_sum(sum);
            if (lane == 0) p.ac[a0 + a] = sum;
        }
        if (tid < 7) {
            long long t = 0;
            for (int w2 = 0; w2 < kWWarps; w2++) t += hdr->warp_part[w2][tid];
            p.lc[l * TRT_LC_N + tid] = t;
        }
        if (tid == 7) p.lc[l * TRT_LC_N + 7] = 0;
        wbar();
    }
}

template <typename K>
int set_smem(trt_ctx* ctx, K kernel, size_t smem) {
    TRT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return TRT_OK;
}

__global__ void group_bits_kernel(const uint8_t* __restrict__ masks, int64_t S, int64_t S_pad, int g0, int ng, uint8_t* __restrict__ bits) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S_pad) return;
    unsigned b = 0;
    if (s < S)
        for (int g = 0; g < ng; g++) b |= (masks[(size_t)(g0 + g) * S + s] != 0 ? 1u : 0u) << g;
    bits[s] = (uint8_t)b;
}

}  // namespace

// The GT scan of the block for G sample groups.  Pair-table tiers: up to kPMaxSampleGroups groups per pass over the GT
// rows (as many as leave the TMA ring three stages of shared memory); wide / generic tiers: one launch per group.
int trt_run_scan(trt_ctx* ctx, const uint8_t* d_masks, int G) {
    const int64_t L = ctx->L, S = ctx->S, nA = ctx->nA;
    if (G < 1 || (G > 1 && !d_masks)) return trt_set_error(ctx, TRT_EINVAL, "scan: %d groups need their masks", G);
    ScanParams sp;
    sp.gt = ctx->d_gt_active;
    sp.pitch = ctx->gt_active_pitch;
    sp.L = L;
    sp.S = S;
    sp.P = ctx->P;
    sp.locus_off = (const int32_t*)ctx->locus_off.p;
    sp.len_class = (const int32_t*)ctx->len_class.p;
    sp.seq_class = (const int32_t*)ctx->seq_class.p;
    sp.len_rank = (const int32_t*)ctx->stat_i32.p;
    sp.seq_rank = (const int32_t*)ctx->stat_i32.p + nA;
    sp.hflags = (const int32_t*)ctx->hflags.p;
    sp.mask = nullptr;
    sp.gbits = nullptr;
    sp.group0 = 0;
    sp.ac_stride = (size_t)nA;
    sp.lc_stride = (size_t)L * TRT_LC_N;
    sp.ac = (int32_t*)ctx->ac.p;
    sp.ac_part = ctx->want_ac_part ? (int32_t*)ctx->ac_part.p : nullptr;
    sp.lc = (long long*)ctx->lc.p;
    const bool fast = (ctx->P == 2 && S >= kMinFastSamples);
    sp.fast_enabled = fast ? 1 : 0;
    sp.list = nullptr;
    sp.n_list = 0;
    sp.stream_only = getenv("TRT_SCAN_STREAM_ONLY") ? 1 : 0;   // HBM-read ceiling of this access pattern (calibration)
    // which tiers occur in this block, and the locus list of each: computed once per block (the allele table fixes them)
    int* n_tier = ctx->scan_n_tier;
    int* max_in_tier = ctx->scan_max_in_tier;
    int* rows_in_tier = ctx->scan_rows_in_tier;
    if (!ctx->scan_lists_valid || ctx->scan_lists_fast != (fast ? 1 : 0)) {
        for (int t = 0; t < TIER_COUNT; t++) n_tier[t] = max_in_tier[t] = rows_in_tier[t] = 0;
        for (int64_t l = 0; l < L; l++) {
            const int A = ctx->h_locus_off[l + 1] - ctx->h_locus_off[l];
            const int t = fast ? scan_tier(A) : TIER_GENERIC;
            n_tier[t]++;
            max_in_tier[t] = std::max(max_in_tier[t], A);
            rows_in_tier[t] = std::max(rows_in_tier[t], pairs_rows(A));
        }
        std::vector<int32_t> lists((size_t)L);
        int off[TIER_COUNT + 1];
        off[0] = 0;
        for (int t = 0; t < TIER_COUNT; t++) off[t + 1] = off[t] + n_tier[t];
        int fill[TIER_COUNT];
        for (int t = 0; t < TIER_COUNT; t++) fill[t] = off[t];
        for (int64_t l = 0; l < L; l++) {
            const int A = ctx->h_locus_off[l + 1] - ctx->h_locus_off[l];
            lists[(size_t)fill[fast ? scan_tier(A) : TIER_GENERIC]++] = (int32_t)l;
        }
        TRT_TRY(trt_ensure(ctx, ctx->scan_lists, (size_t)L * 4 + 16));
        if (L) TRT_CUDA(cudaMemcpyAsync(ctx->scan_lists.p, lists.data(), (size_t)L * 4, cudaMemcpyHostToDevice, ctx->stream));
        TRT_CUDA(cudaStreamSynchronize(ctx->stream));   // `lists` is a local
        for (int t = 0; t <= TIER_COUNT; t++) ctx->scan_tier_off[t] = off[t];
        ctx->scan_lists_valid = true;
        ctx->scan_lists_fast = fast ? 1 : 0;
    }
    const int grid_persist = (int)std::min<int64_t>(std::max<int64_t>(L, 1), ctx->sm_count);
    const size_t smem_limit = (size_t)ctx->max_smem_optin;
    const size_t row_bytes_gt = (((size_t)S * 6 + 15) & ~size_t(15));
    const int nchunks_p = (int)((row_bytes_gt + kPChunkBytes - 1) / kPChunkBytes);
    const bool cells8 = nchunks_p <= kPMaxChunks8 && !getenv("TRT_SCAN_CELLS16");
    const bool masked = d_masks != nullptr;
    const int64_t S_pad = (int64_t)nchunks_p * kPChunkCalls + 8;
    if (masked && (n_tier[TIER_PAIRS_A] || n_tier[TIER_PAIRS_B])) TRT_TRY(trt_ensure(ctx, ctx->scan_gbits, (size_t)S_pad + 16));
    for (int t = TIER_PAIRS_A; t <= TIER_PAIRS_B; t++) {
        if (!n_tier[t]) continue;
        sp.list = (const int32_t*)ctx->scan_lists.p + ctx->scan_tier_off[t];
        sp.n_list = n_tier[t];
        const int rows = rows_in_tier[t];
        const size_t table1 = (size_t)(rows + 1) * kPT * (cells8 ? 1 : 2);
        auto stages_for = [&](int ngh) {
            const size_t hdr = ngh == 1 ? sizeof(PairHeaderT<1>) : (ngh == 2 ? sizeof(PairHeaderT<2>) : sizeof(PairHeaderT<3>));
            const long long left = (long long)smem_limit - (long long)hdr - (long long)table1 * ngh - 256;
            return std::make_pair((int)std::min<long long>(kPMaxStages, left / kPChunkBytes), hdr);
        };
        // sample groups per pass: as many (<= 3) as leave the ring its minimum depth
        int per_pass = 1;
        if (masked)
            for (int n = std::min(G, kPMaxSampleGroups); n >= 1; n--)
                if (stages_for(n).first >= kPMinStages) { per_pass = n; break; }
        const int grid = std::min(grid_persist, n_tier[t]);
        for (int g0 = 0; g0 < G; g0 += per_pass) {
            const int ng = masked ? std::min(per_pass, G - g0) : 0;
            const int ngh = std::max(ng, 1);
            auto sh = stages_for(ngh);
            int stages = std::max(kPMinStages, sh.first);
            if (const char* e = getenv("TRT_SCAN_STAGES")) stages = std::max(2, std::min(stages, atoi(e)));
            const size_t smem = (size_t)stages * kPChunkBytes + sh.second + table1 * ngh;
            if (smem > smem_limit) return trt_set_error(ctx, TRT_ENOMEM, "scan: %zu B of shared memory needed, %zu available", smem, smem_limit);
            if (masked) {
                // bit g of byte s = "sample s is in group g0 + g" for this pass' groups
                group_bits_kernel<<<(unsigned)((S_pad + 255) / 256), 256, 0, ctx->stream>>>(d_masks, S, S_pad, g0, ng, (uint8_t*)ctx->scan_gbits.p);
                TRT_KERNEL_CHECK();
                sp.gbits = (const uint8_t*)ctx->scan_gbits.p;
            }
            sp.group0 = 0;                                   // bits of this pass start at 0 ...
            sp.ac = (int32_t*)ctx->ac.p + (size_t)g0 * nA;   // ... and its outputs at group g0
            sp.ac_part = ctx->want_ac_part ? (int32_t*)ctx->ac_part.p + (size_t)g0 * nA : nullptr;
            sp.lc = (long long*)ctx->lc.p + (size_t)g0 * L * TRT_LC_N;
#define LAUNCH_PAIRS(N, C)                                                                        \
    do {                                                                                          \
        TRT_TRY(set_smem(ctx, scan_pairs_kernel<N, C>, smem));                                    \
        scan_pairs_kernel<N, C><<<grid, kPThreads, smem, ctx->stream>>>(sp, t, rows, stages);     \
    } while (0)
#define LAUNCH_NG(N)                                                               \
    do {                                                                           \
        if (cells8) LAUNCH_PAIRS(N, uint8_t); else LAUNCH_PAIRS(N, uint16_t);      \
    } while (0)
            switch (ng) {
                case 0: LAUNCH_NG(0); break;
                case 1: LAUNCH_NG(1); break;
                case 2: LAUNCH_NG(2); break;
                default: LAUNCH_NG(3); break;
            }
#undef LAUNCH_NG
#undef LAUNCH_PAIRS
            TRT_KERNEL_CHECK();
        }
    }
    // wide / generic tiers: one launch per group under its byte mask
    for (int g = 0; g < G; g++) {
        sp.mask = masked ? d_masks + (size_t)g * S : nullptr;
        sp.ac = (int32_t*)ctx->ac.p + (size_t)g * nA;
        sp.ac_part = ctx->want_ac_part ? (int32_t*)ctx->ac_part.p + (size_t)g * nA : nullptr;
        sp.lc = (long long*)ctx->lc.p + (size_t)g * L * TRT_LC_N;
        if (n_tier[TIER_WIDE]) {
            sp.list = (const int32_t*)ctx->scan_lists.p + ctx->scan_tier_off[TIER_WIDE];
            sp.n_list = n_tier[TIER_WIDE];
            const int amax = max_in_tier[TIER_WIDE];
            const size_t smem = (size_t)kWStages * kWChunkBytes + sizeof(WideHeader) + (size_t)amax * kWT * 2;
            if (masked) {
                TRT_TRY(set_smem(ctx, scan_wide_kernel<true>, smem));
                scan_wide_kernel<true><<<std::min(grid_persist, n_tier[TIER_WIDE]), kWThreads, smem, ctx->stream>>>(sp, amax);
            } else {
                TRT_TRY(set_smem(ctx, scan_wide_kernel<false>, smem));
                scan_wide_kernel<false><<<std::min(grid_persist, n_tier[TIER_WIDE]), kWThreads, smem, ctx->stream>>>(sp, amax);
            }
            TRT_KERNEL_CHECK();
        }
        if (n_tier[TIER_GENERIC]) {
            const int warps_per_block = 8;
            const int64_t blocks = std::min<int64_t>((L + warps_per_block - 1) / warps_per_block, (int64_t)ctx->sm_count * 8);
            scan_generic_kernel<<<(unsigned)std::max<int64_t>(blocks, 1), warps_per_block * 32, 0, ctx->stream>>>(sp);
            TRT_KERNEL_CHECK();
        }
    }
    return TRT_OK;
}
