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
//   scan_pairs_kernel  (diploid, A <= 14 alleles): persistent CTAs, one locus at a time.  A producer
//       warp streams the row through a 3-stage shared-memory ring with 1-D bulk TMA copies
//       (cp.async.bulk + mbarrier tx bytes).  8 consumer warps read 2 x 48 B per thread (conflict-free
//       LDS.128) and do ONE thread-private shared-memory increment per CALL into a genotype-pair table
//       [(A+3)^2][256 threads]; every statistic is derived from the reduced pair table, so there is no
//       per-call flag logic at all (~13 instructions per call; the roofline allows ~33).
//   scan_wide_kernel   (diploid, A <= 96): same TMA ring, thread-private per-haplotype 16-bit counters.
//   scan_generic_kernel: warp per locus; any ploidy, any allele count, tiny sample counts.
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
// pair-table tier
// ---------------------------------------------------------------------------------------------------
constexpr int kPT = 256;                       // consumer threads
constexpr int kPWarps = kPT / 32;
constexpr int kPThreads = kPT + 32;            // + producer warp
constexpr int kPChunkBytes = 24576;            // 4096 calls; thread t owns bytes [48t,48t+48) and [12288+48t, ...)
constexpr int kPChunkCalls = 4096;
constexpr int kPStages = 3;
constexpr int kPMaxDigits = kPairsMaxAllelesU16 + 3;   // 17

struct __align__(16) PairHeader {
    uint64_t full[kPStages];
    uint64_t empty[kPStages];
    unsigned int T[2][kPMaxDigits * kPMaxDigits];   // CTA-reduced pair table, double buffered by locus parity
    int cls_len[2][kPMaxDigits];                    // class of each digit (pad digit: -2; bad/nocall: -1)
    int cls_seq[2][kPMaxDigits];
};

__device__ __forceinline__ void pbar() { asm volatile("bar.sync 1, %0;" ::"n"(kPT) : "memory"); }

// digit of a haplotype: pad(-2) -> 0, no-call(-1) -> 1, allele a -> a+2, anything else -> D-1 ("bad")
__device__ __forceinline__ unsigned digit(int a, unsigned Dm1) { return min((unsigned)(a + 2), Dm1); }

template <typename CT>
__device__ __forceinline__ void bump2(CT* my, unsigned i0, unsigned i1) {
    // two thread-private increments with both loads in flight; equal bins are handled by writing x+2 twice
    CT* p0 = my + (size_t)i0 * kPT;
    CT* p1 = my + (size_t)i1 * kPT;
    const unsigned x0 = *p0, x1 = *p1;
    const unsigned e = (i0 == i1) ? 2u : 1u;
    *p0 = (CT)(x0 + e);
    *p1 = (CT)(x1 + e);
}

template <typename CT, bool MASKED>
__global__ void __launch_bounds__(kPThreads, 1) scan_pairs_kernel(ScanParams p, int tier, int max_digits) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* ring = smem;
    PairHeader* hdr = (PairHeader*)(smem + kPStages * kPChunkBytes);
    CT* table = (CT*)(smem + kPStages * kPChunkBytes + sizeof(PairHeader));   // [D*D][256]

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const size_t row_bytes = (size_t)p.S * 6;
    const size_t copy_bytes = (row_bytes + 15) & ~size_t(15);
    const int nchunks = (int)((copy_bytes + kPChunkBytes - 1) / kPChunkBytes);
    const int nfull = (int)(p.S / kPChunkCalls);   // chunks whose 4096 calls are all real samples

    if (tid == 0) {
        for (int s = 0; s < kPStages; s++) {
            mbar_init(&hdr->full[s], 1);
            mbar_init(&hdr->empty[s], kPWarps);
        }
        mbar_fence_init();
    }
    {
        const int words = max_digits * max_digits * kPT * (int)sizeof(CT) / 4;
        for (int i = tid; i < words; i += kPThreads) ((uint32_t*)table)[i] = 0u;
    }
    __syncthreads();

    if (warp == kPWarps) {
        // ===== producer warp =====
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t l = blockIdx.x; l < p.L; l += gridDim.x) {
                const int A = p.locus_off[l + 1] - p.locus_off[l];
                if (scan_tier(A) != tier) continue;
                const char* src = (const char*)p.gt + (size_t)l * p.pitch;
                for (int c = 0; c < nchunks; c++, it++) {
                    const int stage = it % kPStages;
                    const uint32_t phase = (it / kPStages) & 1u;
                    mbar_wait(&hdr->empty[stage], phase ^ 1u);
                    const size_t off = (size_t)c * kPChunkBytes;
                    const uint32_t bytes = (uint32_t)min((size_t)kPChunkBytes, copy_bytes - off);
                    mbar_arrive_expect_tx(&hdr->full[stage], bytes);
                    tma_load_1d(ring + (size_t)stage * kPChunkBytes, src + off, bytes, &hdr->full[stage]);
                }
            }
        }
        return;
    }

    // ===== consumers =====
    CT* my = table + tid;
    uint32_t it = 0;
    int parity = 0;
    for (int64_t l = blockIdx.x; l < p.L; l += gridDim.x) {
        const int a0 = p.locus_off[l];
        const int A = p.locus_off[l + 1] - a0;
        if (scan_tier(A) != tier) continue;
        const unsigned D = (unsigned)A + 3u, Dm1 = D - 1u;
        if (tid < (int)D) {
            int cl = -1, cq = -1;
            if (tid == 0) cl = cq = -2;
            else if (tid >= 2 && tid < A + 2) {
                cl = p.len_class[a0 + tid - 2];
                cq = p.seq_class[a0 + tid - 2];
            }
            hdr->cls_len[parity][tid] = cl;
            hdr->cls_seq[parity][tid] = cq;
        }

        for (int c = 0; c < nchunks; c++, it++) {
            const int stage = it % kPStages;
            const uint32_t phase = (it / kPStages) & 1u;
            mbar_wait(&hdr->full[stage], phase);
            const unsigned char* base = ring + (size_t)stage * kPChunkBytes;
            const uint4* s0p = (const uint4*)(base + (size_t)tid * 48);
            const uint4* s1p = (const uint4*)(base + 12288 + (size_t)tid * 48);
            const uint4 v0 = s0p[0], v1 = s0p[1], v2 = s0p[2], v3 = s1p[0], v4 = s1p[1], v5 = s1p[2];
            __syncwarp();
            if (lane == 0) mbar_arrive(&hdr->empty[stage]);
            const uint32_t w[24] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w,
                                    v3.x, v3.y, v3.z, v3.w, v4.x, v4.y, v4.z, v4.w, v5.x, v5.y, v5.z, v5.w};
            unsigned idx[16];
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const int piece = j >> 3, jj = j & 7;
                const int k0 = piece * 24 + 3 * jj, k1 = k0 + 1;      // half-word indices
                const int a = (k0 & 1) ? ((int)w[k0 >> 1] >> 16) : (int)(short)(w[k0 >> 1] & 0xffffu);
                const int b = (k1 & 1) ? ((int)w[k1 >> 1] >> 16) : (int)(short)(w[k1 >> 1] & 0xffffu);
                idx[j] = digit(a, Dm1) * D + digit(b, Dm1);
            }
            if (!MASKED && c < nfull) {
#pragma unroll
                for (int j = 0; j < 16; j += 2) bump2<CT>(my, idx[j], idx[j + 1]);
            } else {
                const int64_t sb0 = (int64_t)c * kPChunkCalls + (int64_t)tid * 8;
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const int64_t s = sb0 + (j >> 3) * 2048 + (j & 7);
                    bool live = s < p.S;
                    if (MASKED) live = live && p.mask[live ? s : 0] != 0;
                    if (live) my[(size_t)idx[j] * kPT] += 1;
                }
            }
        }

        // ---- reduce the thread-private tables of this locus to the CTA table ----------------------
        pbar();
        const int nb = (int)(D * D);
        unsigned int* T = hdr->T[parity];
        for (int b = warp; b < nb; b += kPWarps) {
            uint32_t* rowp = (uint32_t*)(table + (size_t)b * kPT);
            unsigned sum = 0;
            if (sizeof(CT) == 4) {
#pragma unroll
                for (int k = 0; k < kPT / 32; k++) {
                    sum += rowp[lane + 32 * k];
                    rowp[lane + 32 * k] = 0u;
                }
            } else {
#pragma unroll
                for (int k = 0; k < kPT / 64; k++) {
                    const uint32_t x = rowp[lane + 32 * k];
                    rowp[lane + 32 * k] = 0u;
                    sum += (x & 0xffffu) + (x >> 16);
                }
            }
            sum = (unsigned)warp_sum((int)sum);
            if (lane == 0) T[b] = sum;
        }
        pbar();
        // ---- derive the per-locus outputs from the pair table (the other warps run ahead) ----------
        if (warp == 0) {
            const int* cl = hdr->cls_len[parity];
            const int* cq = hdr->cls_seq[parity];
            long long n_full = 0, n_non = 0, n_pad = 0, h_idx = 0, h_len = 0, h_seq = 0, n_bad = 0;
            for (int b = lane; b < nb; b += 32) {
                const unsigned d0 = (unsigned)b / D, d1 = (unsigned)b % D;
                const long long n = T[b];
                if (n == 0) continue;
                const bool bad = (d0 == Dm1) | (d1 == Dm1);
                const bool m1 = (d0 == 1u) | (d1 == 1u) | bad;
                const bool v0 = (d0 >= 2u) & (d0 < Dm1), v1 = (d1 >= 2u) & (d1 < Dm1);
                if (bad) n_bad += n;
                if (v0 | v1) n_non += n;
                if (!m1) {
                    n_full += n;
                    if ((d0 == 0u) | (d1 == 0u)) n_pad += n;
                    if (d0 == d1) h_idx += n;
                    if (cl[d0] == cl[d1]) h_len += n;
                    if (cq[d0] == cq[d1]) h_seq += n;
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
            for (int a = lane; a < A; a += 32) {
                const unsigned d = (unsigned)a + 2u;
                unsigned cnt = 0;
                for (unsigned e = 0; e < D; e++) cnt += T[d * D + e] + T[e * D + d];
                p.ac[a0 + a] = (int)cnt;
            }
        }
        parity ^= 1;
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
            for (int64_t l = blockIdx.x; l < p.L; l += gridDim.x) {
                const int A = p.locus_off[l + 1] - p.locus_off[l];
                if (scan_tier(A) != TIER_WIDE) continue;
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
    for (int64_t l = blockIdx.x; l < p.L; l += gridDim.x) {
        const int a0 = p.locus_off[l];
        const int A = p.locus_off[l + 1] - a0;
        if (scan_tier(A) != TIER_WIDE) continue;
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
            __syncwarp();
            if (lane == 0) mbar_arrive(&hdr->empty[stage]);
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
                if (!m1) {
                    n_full++;
                    n_pad += ((a == -2) | (b == -2));
                    h_idx += (a == b);
                    const uint32_t ca = va ? hdr->cls[a] : 0xfffe0000u, cb = vb ? hdr->cls[b] : 0xfffd0001u;
                    h_len += ((ca >> 16) == (cb >> 16)) | (a == b);
                    h_seq += ((ca & 0xffffu) == (cb & 0xffffu)) | (a == b);
                }
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

}  // namespace

// run the scan (all tiers) for one group mask; results into ctx->ac / ctx->lc at group g
int trt_run_scan(trt_ctx* ctx, const uint8_t* d_mask, int g, int G) {
    (void)G;
    const int64_t L = ctx->L, S = ctx->S, nA = ctx->nA;
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
    sp.mask = d_mask;
    sp.ac = (int32_t*)ctx->ac.p + (size_t)g * nA;
    sp.lc = (long long*)ctx->lc.p + (size_t)g * L * TRT_LC_N;
    const bool fast = (ctx->P == 2 && S >= kMinFastSamples);
    sp.fast_enabled = fast ? 1 : 0;
    // which tiers occur in this block
    int n_tier[4] = {0, 0, 0, 0};
    int max_in_tier[4] = {0, 0, 0, 0};
    for (int64_t l = 0; l < L; l++) {
        const int A = ctx->h_locus_off[l + 1] - ctx->h_locus_off[l];
        const int t = fast ? scan_tier(A) : TIER_GENERIC;
        n_tier[t]++;
        max_in_tier[t] = std::max(max_in_tier[t], A);
    }
    const int grid_persist = (int)std::min<int64_t>(L, ctx->sm_count);
    if (n_tier[TIER_PAIRS32]) {
        const int D = max_in_tier[TIER_PAIRS32] + 3;
        const size_t smem = (size_t)kPStages * kPChunkBytes + sizeof(PairHeader) + (size_t)D * D * kPT * 4;
        if (d_mask) {
            TRT_TRY(set_smem(ctx, scan_pairs_kernel<uint32_t, true>, smem));
            scan_pairs_kernel<uint32_t, true><<<grid_persist, kPThreads, smem, ctx->stream>>>(sp, TIER_PAIRS32, D);
        } else {
            TRT_TRY(set_smem(ctx, scan_pairs_kernel<uint32_t, false>, smem));
            scan_pairs_kernel<uint32_t, false><<<grid_persist, kPThreads, smem, ctx->stream>>>(sp, TIER_PAIRS32, D);
        }
        TRT_KERNEL_CHECK();
    }
    if (n_tier[TIER_PAIRS16]) {
        const int D = max_in_tier[TIER_PAIRS16] + 3;
        const size_t smem = (size_t)kPStages * kPChunkBytes + sizeof(PairHeader) + (size_t)D * D * kPT * 2;
        if (d_mask) {
            TRT_TRY(set_smem(ctx, scan_pairs_kernel<uint16_t, true>, smem));
            scan_pairs_kernel<uint16_t, true><<<grid_persist, kPThreads, smem, ctx->stream>>>(sp, TIER_PAIRS16, D);
        } else {
            TRT_TRY(set_smem(ctx, scan_pairs_kernel<uint16_t, false>, smem));
            scan_pairs_kernel<uint16_t, false><<<grid_persist, kPThreads, smem, ctx->stream>>>(sp, TIER_PAIRS16, D);
        }
        TRT_KERNEL_CHECK();
    }
    if (n_tier[TIER_WIDE]) {
        const int amax = max_in_tier[TIER_WIDE];
        const size_t smem = (size_t)kWStages * kWChunkBytes + sizeof(WideHeader) + (size_t)amax * kWT * 2;
        if (d_mask) {
            TRT_TRY(set_smem(ctx, scan_wide_kernel<true>, smem));
            scan_wide_kernel<true><<<grid_persist, kWThreads, smem, ctx->stream>>>(sp, amax);
        } else {
            TRT_TRY(set_smem(ctx, scan_wide_kernel<false>, smem));
            scan_wide_kernel<false><<<grid_persist, kWThreads, smem, ctx->stream>>>(sp, amax);
        }
        TRT_KERNEL_CHECK();
    }
    if (n_tier[TIER_GENERIC]) {
        const int warps_per_block = 8;
        const int64_t blocks = std::min<int64_t>((L + warps_per_block - 1) / warps_per_block, (int64_t)ctx->sm_count * 8);
        scan_generic_kernel<<<(unsigned)std::max<int64_t>(blocks, 1), warps_per_block * 32, 0, ctx->stream>>>(sp);
        TRT_KERNEL_CHECK();
    }
    return TRT_OK;
}
