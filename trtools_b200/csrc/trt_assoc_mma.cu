// K4 tensor path — the associaTR cross moments as an exact integer contraction on the tensor cores.
//
// Per locus the regression (reference associaTR.py:246-291) needs sum g z_k over the called design samples for the
// design's K real columns z = (y, c_1 .. c_{K-1}): a skinny GEMM  G[loci x samples] . Z[samples x K].  FP64 FMAs cap it
// at 37 % of the FP64 pipe (shared-memory operand traffic, trt_assoc_tile.cu); here both operands become 8-bit integers
// and the products are accumulated EXACTLY in int32 by mma.sync.m16n8k32 (u8 x s8):
//   * G: a length genotype is (n_a + n_b) * unit with n = (allele bp length - shortest allele's) / gcd, a small
//     non-negative integer per haplotype (0..127, so a call fits a byte).  The shift by the shortest allele is a constant
//     added to g, which the regression on (g, intercept, ...) does not see.  Loci that do not fit (spread > 127 units, a
//     length that is not a whole number of bp, a ploidy pad among the called samples) are flagged for the FP64 tile path.
//   * Z: every column is cut into J = 7 balanced base-256 digits of a 54-bit fixed-point value (scaled by the column's
//     power-of-two exponent), one int8 column per digit, plus one 0/1 column "sample is in the design".  The int32
//     results are recombined in FP64: sum_j 256^j M_j is the exact dot product of the fixed-point column, so the only
//     rounding is the 2^-54 (of the column maximum) quantisation of z — below FP64's own 2^-53.
// A warp owns 16 loci; a lane builds its A fragment straight from the native cyvcf2 GT rows in shared memory (the
// MMA's k index is mapped to samples so that a lane's 8 bytes are 8 CONSECUTIVE samples: three conflict-free LDS.128
// per row) with a 16-entry byte table looked up by PRMT; B fragments are one LDS.64 per 8 digit columns.  A producer
// warp streams [192 loci] x [32 samples] GT boxes (2-D TMA) and the 32 samples' digit rows (1-D bulk copy) through a
// ring.  Uncalled design samples are emitted as one 32-bit mask per (locus, 32 samples) for the down-date kernel.
//
// Algorithmic traffic: 6 B/call of GT once; the digit matrix (S x 88 B) is L2-resident.
#include <cuda.h>
#include <math.h>

#include <algorithm>

#include "trt_assoc_tile.cuh"

namespace {

#ifndef TRT_MMA_WARPS
#define TRT_MMA_WARPS 12
#endif
constexpr int kMW = TRT_MMA_WARPS;              // consumer warps
constexpr int kMLoci = kAssocMmaTileLoci;       // loci per tile (16 per warp)
static_assert(kMLoci == kMW * 16, "tile = consumer warps x 16 loci");
constexpr int kMThreads = kMW * 32 + 32;        // + producer warp
constexpr int kMRow = 32 * 6;                   // bytes of one locus' 32 calls
constexpr int kMGtBytes = kMLoci * kMRow;       // one stage of GT
constexpr int kMJ = 7;                          // digits per design column
constexpr int kMMaxStages = 6;
constexpr int kMMaxNT = 15;                     // 8-column groups of the digit matrix (K <= 16: 16 * 7 + 1 = 113 columns)
constexpr int kMSegKsteps = 2016;               // 32-sample steps per segment: 254 * 128 * 32 * 2016 < 2^31

struct MmaParams {
    int64_t L, S;
    int n_tiles, nk, nseg, ks_per_seg, nwin, stages;
    int ncolp;                      // digit columns padded to a multiple of 8
    int ind_col;                    // the 0/1 design-membership column
    const uint8_t* tab;             // [n_tiles * kMLoci][16] haplotype value by (allele & 15); 0x80 = not a called allele
    const int8_t* xd;               // [nk][ncolp][32]
    int32_t* part;                  // [nseg][n_tiles * kMLoci][ncolp + 4]: digit sums, then n, sum a^2 (lo, hi)
    uint32_t* masks;                // [n_tiles][nwin][kMLoci][32]: bit i of word (k & 31) of window k >> 5 = sample 32 k + i is an
                                    // uncalled design sample
};

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int x, int y, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}

// PRMT with selector 0xBA98: the sign bit of every byte of x replicated over that byte (the msb of a selector nibble asks
// for sign replication; __byte_perm documents three selector bits only, so the instruction is spelled out)
__device__ __forceinline__ uint32_t sign_bytes(uint32_t x) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %1, 0xBA98;" : "=r"(r) : "r"(x));
    return r;
}

// 16-entry byte table (t0..t3, entry d in byte d & 3 of t[d >> 2]) looked up for the low nibbles of the 4 bytes of x
__device__ __forceinline__ uint32_t lut16(uint32_t x, uint32_t t0, uint32_t t1, uint32_t t2, uint32_t t3) {
    const uint32_t nib = x & 0x07070707u;
    const uint32_t t = nib | (nib >> 4);                  // byte 0: d0 | d1 << 4, byte 2: d2 | d3 << 4 (low 3 bits each)
    const uint32_t sel = __byte_perm(t, 0u, 0x4420);      // PRMT selector nibbles d0, d1, d2, d3
    const uint32_t lo = __byte_perm(t0, t1, sel);
    const uint32_t hi = __byte_perm(t2, t3, sel);
    const uint32_t m = sign_bytes(x << 4);                // bit 3 of every byte, replicated over the byte
    return (hi & m) | (lo & ~m);
}

// 8 consecutive calls of one locus (48 bytes of cyvcf2 GT: int16 a0, a1, phase per call) -> two A-fragment registers
// (4 call values each), the 8 "uncalled design sample" bits, the called count and sum of squares
__device__ __forceinline__ void decode8(const uint4 v0, const uint4 v1, const uint4 v2, const uint4 tb, uint32_t ind_lo,
                                        uint32_t ind_hi, uint32_t& a_lo, uint32_t& a_hi, uint32_t& bits, uint32_t& n,
                                        uint32_t& ssq) {
    // (a1 << 16 | a0) of the 8 calls
    const uint32_t p0 = v0.x, p1 = __byte_perm(v0.y, v0.z, 0x5432), p2 = v0.w, p3 = __byte_perm(v1.x, v1.y, 0x5432);
    const uint32_t p4 = v1.z, p5 = __byte_perm(v1.w, v2.x, 0x5432), p6 = v2.y, p7 = __byte_perm(v2.z, v2.w, 0x5432);
    // low bytes: (a0 c, a1 c, a0 c+1, a1 c+1)
    const uint32_t r01 = lut16(__byte_perm(p0, p1, 0x6420), tb.x, tb.y, tb.z, tb.w);
    const uint32_t r23 = lut16(__byte_perm(p2, p3, 0x6420), tb.x, tb.y, tb.z, tb.w);
    const uint32_t r45 = lut16(__byte_perm(p4, p5, 0x6420), tb.x, tb.y, tb.z, tb.w);
    const uint32_t r67 = lut16(__byte_perm(p6, p7, 0x6420), tb.x, tb.y, tb.z, tb.w);
    {
        const uint32_t u = __byte_perm(r01, r23, 0x6420), v = __byte_perm(r01, r23, 0x7531);   // first / second haplotypes
        const uint32_t inv = sign_bytes(u | v);                                                 // 0xff: not (strictly) called
        const uint32_t keep = ~inv & ind_lo;
        a_lo = (u & keep) + (v & keep);                    // bytes <= 254: no carry between them
        const uint32_t unc = inv & ind_lo;
        bits = ((unc & 0x08040201u) * 0x01010101u) >> 24;
        n += __popc(keep & 0x01010101u);
        ssq = __dp4a(a_lo, a_lo, ssq);
    }
    {
        const uint32_t u = __byte_perm(r45, r67, 0x6420), v = __byte_perm(r45, r67, 0x7531);
        const uint32_t inv = sign_bytes(u | v);
        const uint32_t keep = ~inv & ind_hi;
        a_hi = (u & keep) + (v & keep);
        const uint32_t unc = inv & ind_hi;
        bits |= (((unc & 0x08040201u) * 0x01010101u) >> 24) << 4;
        n += __popc(keep & 0x01010101u);
        ssq = __dp4a(a_hi, a_hi, ssq);
    }
}

__device__ __forceinline__ void mma_u8s8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int NT>
__global__ void __launch_bounds__(kMThreads, 1) assoc_mma_kernel(const __grid_constant__ CUtensorMap tmap, MmaParams p) {
    constexpr int kXBytes = NT * 8 * 32;                  // digit rows of one 32-sample step
    extern __shared__ __align__(128) unsigned char smem[];
    const int stages = p.stages;
    unsigned char* gt_ring = smem;                                             // [stages][kMGtBytes]
    unsigned char* x_ring = smem + (size_t)stages * kMGtBytes;                 // [stages][kXBytes]
    uint64_t* full = (uint64_t*)(x_ring + (size_t)stages * kXBytes);
    uint64_t* empty = full + kMMaxStages;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < stages; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kMW);
        }
        mbar_fence_init();
    }
    __syncthreads();
    const int n_units = p.n_tiles * p.nseg;

    if (warp == kMW) {
        // ===== producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
                const int tile = u / p.nseg, seg = u % p.nseg;
                const int k0 = seg * p.ks_per_seg, k1 = min(p.nk, k0 + p.ks_per_seg);
                for (int ks = k0; ks < k1; ks++) {
                    mbar_wait(&empty[stage], phase ^ 1u);
                    mbar_arrive_expect_tx(&full[stage], (uint32_t)(kMGtBytes + kXBytes));
                    tma_load_2d(gt_ring + (size_t)stage * kMGtBytes, &tmap, ks * 96, tile * kMLoci, &full[stage]);
                    tma_load_1d(x_ring + (size_t)stage * kXBytes, p.xd + (size_t)ks * kXBytes, kXBytes, &full[stage]);
                    if (++stage == stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
        return;
    }

    // ===== consumers: warp <-> 16 loci; lane (g, t) <-> loci g, g + 8 of the warp, samples 8 t .. 8 t + 7 of the step =====
    const int g = lane >> 2, t4 = lane & 3;
    int stage = 0;
    uint32_t phase = 0;
    const int pw = p.ncolp + 4;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        const int tile = u / p.nseg, seg = u % p.nseg;
        const int k0 = seg * p.ks_per_seg, k1 = min(p.nk, k0 + p.ks_per_seg);
        const int row0 = warp * 16 + g, row1 = row0 + 8;                    // rows of the tile
        const size_t lrow0 = (size_t)tile * kMLoci + row0, lrow1 = lrow0 + 8;
        const uint4 tb0 = *(const uint4*)(p.tab + lrow0 * 16), tb1 = *(const uint4*)(p.tab + lrow1 * 16);
        int acc[NT][4];
#pragma unroll
        for (int j = 0; j < NT; j++) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0;
        uint32_t n0 = 0, n1 = 0, q0 = 0, q1 = 0;       // called design samples, sum of squares (this lane's samples)
        unsigned long long qq0 = 0, qq1 = 0;
        uint32_t keep0 = 0, keep1 = 0;                  // mask words of the step (ks & 3) == t4
        for (int ks = k0; ks < k1; ks++) {
            mbar_wait(&full[stage], phase);
            const unsigned char* gts = gt_ring + (size_t)stage * kMGtBytes;
            const unsigned char* xs = x_ring + (size_t)stage * kXBytes;
            const uint4* r0 = (const uint4*)(gts + row0 * kMRow + t4 * 48);
            const uint4* r1 = (const uint4*)(gts + row1 * kMRow + t4 * 48);
            const uint4 a0v = r0[0], a1v = r0[1], a2v = r0[2];
            const uint4 b0v = r1[0], b1v = r1[1], b2v = r1[2];
            const uint2 iv = *(const uint2*)(xs + p.ind_col * 32 + t4 * 8);
            const uint32_t ind_lo = iv.x * 0xffu, ind_hi = iv.y * 0xffu;   // 0/1 bytes -> 0x00/0xff
            uint32_t fa0, fa1, fa2, fa3, bits0, bits1;
            decode8(a0v, a1v, a2v, tb0, ind_lo, ind_hi, fa0, fa2, bits0, n0, q0);
            decode8(b0v, b1v, b2v, tb1, ind_lo, ind_hi, fa1, fa3, bits1, n1, q1);
            unsigned dep = a2v.w ^ b2v.w ^ iv.y;        // words of the last-issued loads (see the arrive below)
#pragma unroll
            for (int j = 0; j < NT; j++) {
                const uint2 bv = *(const uint2*)(xs + j * 256 + lane * 8);
                mma_u8s8(acc[j], fa0, fa1, fa2, fa3, bv.x, bv.y);
                dep ^= bv.y;
            }
            {
                // release the stage only after every shared-memory load of the step has RETURNED: the refill is an
                // async-proxy write that is not ordered behind generic-proxy reads still in flight (same rule as the
                // FP64 tile kernel and the GT scan)
                const bool returned = __any_sync(0xffffffffu, dep != 0x9e3779b9u);
                if (lane == 0 && returned) mbar_arrive(&empty[stage]);
            }
            if (++stage == stages) { stage = 0; phase ^= 1u; }
            // mask words: the four lanes of a row OR their bytes; lane t keeps the word of step (ks & 3) == t so that
            // four steps leave the row's lanes holding 16 consecutive bytes
            uint32_t w0 = bits0 << (8 * t4), w1 = bits1 << (8 * t4);
            w0 |= __shfl_xor_sync(0xffffffffu, w0, 1);
            w1 |= __shfl_xor_sync(0xffffffffu, w1, 1);
            w0 |= __shfl_xor_sync(0xffffffffu, w0, 2);
            w1 |= __shfl_xor_sync(0xffffffffu, w1, 2);
            if ((ks & 3) == t4) { keep0 = w0; keep1 = w1; }
            if ((ks & 3) == 3 || ks == k1 - 1) {
                const int kw = (ks & ~3) + t4;
                if (kw <= ks) {
                    uint32_t* mrow = p.masks + ((size_t)tile * p.nwin + (kw >> 5)) * kMLoci * 32 + (kw & 31);
                    mrow[(size_t)row0 * 32] = keep0;
                    mrow[(size_t)row1 * 32] = keep1;
                }
            }
            if ((ks & 255) == 255) {     // the 32-bit sums of squares: <= 8 * 254^2 per step
                qq0 += q0; qq1 += q1;
                q0 = q1 = 0;
            }
        }
        qq0 += q0; qq1 += q1;
        // ---- results of the unit ----
        int32_t* o0 = p.part + ((size_t)seg * p.n_tiles * kMLoci + lrow0) * pw;
        int32_t* o1 = p.part + ((size_t)seg * p.n_tiles * kMLoci + lrow1) * pw;
#pragma unroll
        for (int j = 0; j < NT; j++) {
            *(int2*)(o0 + j * 8 + 2 * t4) = make_int2(acc[j][0], acc[j][1]);
            *(int2*)(o1 + j * 8 + 2 * t4) = make_int2(acc[j][2], acc[j][3]);
        }
#pragma unroll
        for (int o = 1; o < 4; o <<= 1) {
            n0 += __shfl_xor_sync(0xffffffffu, n0, o);
            n1 += __shfl_xor_sync(0xffffffffu, n1, o);
            qq0 += __shfl_xor_sync(0xffffffffu, qq0, o);
            qq1 += __shfl_xor_sync(0xffffffffu, qq1, o);
        }
        if (t4 == 0) {
            o0[p.ncolp] = (int32_t)n0;
            o0[p.ncolp + 1] = (int32_t)(uint32_t)(qq0 & 0xffffffffull);
            o0[p.ncolp + 2] = (int32_t)(uint32_t)(qq0 >> 32);
            o1[p.ncolp] = (int32_t)n1;
            o1[p.ncolp + 1] = (int32_t)(uint32_t)(qq1 & 0xffffffffull);
            o1[p.ncolp + 2] = (int32_t)(uint32_t)(qq1 >> 32);
        }
    }
}

// ---- per locus: does it fit the integer form, and its haplotype table ------------------------------------------------
__global__ void assoc_mma_prep_kernel(int64_t L, int64_t L_pad, const int32_t* __restrict__ locus_off, const double* __restrict__ allele_len,
                                      const int32_t* __restrict__ period, const long long* __restrict__ lc, uint8_t* __restrict__ tab,
                                      double* __restrict__ scale, uint8_t* __restrict__ flags, int* __restrict__ n_fp64) {
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L_pad) return;
    uint8_t t[16];
    for (int i = 0; i < 16; i++) t[i] = 0x80;
    uint8_t flag = 0;
    double sc = 0.0;
    if (l < L) {
        const int a0 = locus_off[l], A = locus_off[l + 1] - a0;
        if (A <= kAssocFastMaxAlleles) {
            flag = 2;
            const int per = period[l];
            bool ok = (A >= 1) && (per >= 1) && (lc[l * TRT_LC_N + TRT_LC_NPAD] == 0);
            long long m[kAssocFastMaxAlleles];
            long long gcd = 0, mn = 0, mx = 0;
            if (ok) {
                const double ref = allele_len[a0];
                for (int a = 0; a < A; a++) {
                    const double d = (allele_len[a0 + a] - ref) * (double)per;
                    const double r = rint(d);
                    if (!(fabs(d - r) <= 1e-9 * (1.0 + fabs(r))) || fabs(r) > 1e9) { ok = false; break; }
                    m[a] = (long long)r;
                    long long x = m[a] < 0 ? -m[a] : m[a], y = gcd;
                    while (x) { const long long z = y % x; y = x; x = z; }
                    gcd = y;
                    mn = (a == 0 || m[a] < mn) ? m[a] : mn;
                    mx = (a == 0 || m[a] > mx) ? m[a] : mx;
                }
            }
            if (ok) {
                if (gcd == 0) gcd = 1;
                if ((mx - mn) / gcd > 127) ok = false;
            }
            if (ok) {
                for (int a = 0; a < A; a++) t[a] = (uint8_t)((m[a] - mn) / gcd);
                sc = (double)gcd / (double)per;
                flag = 1;
            } else {
                atomicAdd(n_fp64, 1);
            }
        }
    }
    // entries 14 (pad, -2 & 15) and 15 (no call, -1 & 15) stay 0x80; an ineligible locus has every entry 0x80: all calls drop out
    uint4 v;
    v.x = t[0] | t[1] << 8 | t[2] << 16 | (uint32_t)t[3] << 24;
    v.y = t[4] | t[5] << 8 | t[6] << 16 | (uint32_t)t[7] << 24;
    v.z = t[8] | t[9] << 8 | t[10] << 16 | (uint32_t)t[11] << 24;
    v.w = t[12] | t[13] << 8 | t[14] << 16 | (uint32_t)t[15] << 24;
    *(uint4*)(tab + l * 16) = v;
    if (l < L) {
        scale[l] = sc;
        flags[l] = flag;
    }
}

// ---- design columns -> digit matrix ----------------------------------------------------------------------------------
// colscale[k] = 2^(E_k - 8 J + 2), colscale[K + k] = its inverse, E_k = exponent of the column's largest magnitude
__global__ void __launch_bounds__(256) assoc_colscale_kernel(const double* __restrict__ covars, const double* __restrict__ outcome, int64_t n,
                                                             int K, double* __restrict__ colscale) {
    __shared__ double part[256];
    const int k = blockIdx.x;      // z column: 0 = outcome, 1 .. K-1 = covars column k
    double m = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 256) {
        const double v = fabs(k == 0 ? outcome[i] : covars[i * K + k]);
        if (v > m) m = v;           // NaN never enters a design (the host drops such rows)
    }
    part[threadIdx.x] = m;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) part[threadIdx.x] = fmax(part[threadIdx.x], part[threadIdx.x + w]);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        int e = 0;
        if (part[0] > 0.0 && isfinite(part[0])) frexp(part[0], &e);     // max = f * 2^e, f in [0.5, 1)
        colscale[k] = ldexp(1.0, e - 8 * kMJ + 2);
        colscale[K + k] = ldexp(1.0, 8 * kMJ - 2 - e);
    }
}

__global__ void assoc_digits_kernel(const double* __restrict__ covars, const double* __restrict__ outcome,
                                    const int32_t* __restrict__ row_of_sample, int64_t S, int64_t S_pad, int K, int ncolp,
                                    const double* __restrict__ colscale, int8_t* __restrict__ xd) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S_pad * (K + 1)) return;
    const int64_t s = i % S_pad;             // consecutive threads: consecutive samples of one column
    const int k = (int)(i / S_pad);
    const int r = (s < S) ? row_of_sample[s] : -1;
    int8_t* base = xd + (size_t)(s >> 5) * ncolp * 32 + (s & 31);
    if (k == K) {
        base[(size_t)(K * kMJ) * 32] = (r >= 0) ? 1 : 0;
        return;
    }
    long long q = 0;
    if (r >= 0) q = llrint((k == 0 ? outcome[r] : covars[(int64_t)r * K + k]) * colscale[K + k]);
#pragma unroll
    for (int j = 0; j < kMJ; j++) {
        const int d = (int)((q + 128) & 255) - 128;       // balanced digit in [-128, 127]
        q = (q - d) >> 8;
        base[(size_t)(k * kMJ + j) * 32] = (int8_t)d;
    }
}

// ---- int32 digit sums -> the solve kernel's FP64 moments -----------------------------------------------------------------
__global__ void assoc_mma_finish_kernel(MmaParams p, int K, const uint8_t* __restrict__ flags, const double* __restrict__ scale,
                                        const double* __restrict__ colscale, double* __restrict__ mom) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.L * (K + 1)) return;
    const int64_t l = i / (K + 1);
    const int k = (int)(i % (K + 1));
    if (flags[l] != 1) return;
    const int pw = p.ncolp + 4, nacc = K + 3;
    const size_t seg_stride = (size_t)p.n_tiles * kMLoci * pw;
    const int32_t* row = p.part + (size_t)l * pw;
    const double sc = scale[l];
    if (k < K) {
        double v = 0.0;
        for (int j = kMJ - 1; j >= 0; j--) {
            long long m = 0;
            for (int sg = 0; sg < p.nseg; sg++) m += row[sg * seg_stride + k * kMJ + j];
            v = v * 256.0 + (double)m;
        }
        mom[l * nacc + 3 + k] = v * colscale[k] * sc;
    } else {
        long long n = 0, sn = 0;
        unsigned long long ss = 0;
        for (int sg = 0; sg < p.nseg; sg++) {
            const int32_t* r = row + sg * seg_stride;
            n += r[p.ncolp];
            sn += r[p.ind_col];
            ss += (unsigned long long)(uint32_t)r[p.ncolp + 1] | ((unsigned long long)(uint32_t)r[p.ncolp + 2] << 32);
        }
        mom[l * nacc + 0] = (double)n;
        mom[l * nacc + 1] = (double)sn * sc;
        mom[l * nacc + 2] = (double)ss * sc * sc;
    }
}

template <int NT>
int launch_mma(trt_ctx* ctx, const CUtensorMap& tmap, MmaParams& mp, int grid) {
    constexpr size_t xbytes = (size_t)NT * 8 * 32;
    const size_t fixed = 2 * kMMaxStages * 8 + 128;
    int stages = (int)(((size_t)ctx->max_smem_optin - fixed) / (kMGtBytes + xbytes));
    stages = std::min(stages, kMMaxStages);
    if (stages < 2) return trt_set_error(ctx, TRT_ENOMEM, "assoc mma kernel: not enough shared memory");
    mp.stages = stages;
    const size_t smem = (size_t)stages * (kMGtBytes + xbytes) + fixed;
    TRT_CUDA(cudaFuncSetAttribute(assoc_mma_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    assoc_mma_kernel<NT><<<grid, kMThreads, smem, ctx->stream>>>(tmap, mp);
    TRT_KERNEL_CHECK();
    return TRT_OK;
}

}  // namespace

bool trt_assoc_mma_supported(const trt_ctx* ctx) {
    return ctx->K * kMJ + 1 <= kMMaxNT * 8 && ctx->S >= kAssocFastMinSamples && !getenv("TRT_ASSOC_NO_MMA");
}

// column scales of the current design (trt_assoc_set_design): needs ctx->covars / outcome / n_design / K
int trt_assoc_mma_design(trt_ctx* ctx) {
    const int K = ctx->K;
    TRT_TRY(trt_ensure(ctx, ctx->assoc_colscale, (size_t)2 * K * 8 + 16));
    assoc_colscale_kernel<<<K, 256, 0, ctx->stream>>>((const double*)ctx->covars.p, (const double*)ctx->outcome.p, ctx->n_design, K,
                                                      (double*)ctx->assoc_colscale.p);
    TRT_KERNEL_CHECK();
    return TRT_OK;
}

// Moments (mom) and down-dates (dd) of every locus that fits the integer form; ctx->assoc_flags [L] afterwards holds
// 1 for those, 2 for loci with <= kAssocFastMaxAlleles alleles that need the FP64 tile path (*n_fp64 of them), 0 for wider
// loci (generic kernels).  Requires the scan's per-locus counters (ctx->lc, group 0) of the design samples.
int trt_assoc_mma(trt_ctx* ctx, const int32_t* d_row_of_sample, double* mom, double* dd, int* n_fp64) {
    const int64_t L = ctx->L, S = ctx->S;
    const int K = ctx->K;
    *n_fp64 = 0;
    if (L == 0) return TRT_OK;
    EncodeTiledFn enc = trt_tmap_encode_fn();
    if (!enc) return trt_set_error(ctx, TRT_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const int ncol = K * kMJ + 1, nt = (ncol + 7) / 8, ncolp = nt * 8;
    const int n_tiles = (int)((L + kMLoci - 1) / kMLoci);
    const int64_t L_pad = (int64_t)n_tiles * kMLoci;
    const int nk = (int)((S + 31) / 32);
    const int64_t S_pad = (int64_t)nk * 32;
    const int sms = ctx->sm_count;
    // work units = tile x sample segment: segments of whole 32-step windows, at most kMSegKsteps steps (int32 range),
    // as many as make the persistent grid's last wave >= 95 % full
    int nseg = (nk + kMSegKsteps - 1) / kMSegKsteps;
    for (; nseg < 16; nseg++) {
        const int64_t units = (int64_t)n_tiles * nseg;
        const int64_t waves = (units + sms - 1) / sms;
        if ((double)units / (double)(waves * sms) >= 0.95 || nk / (nseg + 1) < 64) break;
    }
    int kps = (nk + nseg - 1) / nseg;
    kps = (kps + 31) & ~31;
    nseg = (nk + kps - 1) / kps;
    const int nwin = (nk + 31) / 32;
    const int pw = ncolp + 4;

    TRT_TRY(trt_ensure(ctx, ctx->assoc_flags, (size_t)L + 16));
    TRT_TRY(trt_ensure(ctx, ctx->assoc_mma_tab, (size_t)L_pad * 16 + (size_t)L * 8 + 64));      // tables | scales | counter
    TRT_TRY(trt_ensure(ctx, ctx->assoc_xd, (size_t)nk * ncolp * 32 + 64));
    TRT_TRY(trt_ensure(ctx, ctx->assoc_mma_part, (size_t)nseg * L_pad * pw * 4 + 64));
    TRT_TRY(trt_ensure(ctx, ctx->assoc_mma_masks, (size_t)n_tiles * nwin * kMLoci * 32 * 4 + 64));
    uint8_t* tab = (uint8_t*)ctx->assoc_mma_tab.p;
    double* scale = (double*)(tab + (size_t)L_pad * 16);
    int* d_nfp64 = (int*)(scale + L);
    TRT_CUDA(cudaMemsetAsync(d_nfp64, 0, 4, ctx->stream));
    assoc_mma_prep_kernel<<<(unsigned)((L_pad + 127) / 128), 128, 0, ctx->stream>>>(
        L, L_pad, (const int32_t*)ctx->locus_off.p, (const double*)ctx->allele_len.p, (const int32_t*)ctx->period.p,
        (const long long*)ctx->lc.p, tab, scale, (uint8_t*)ctx->assoc_flags.p, d_nfp64);
    TRT_KERNEL_CHECK();
    TRT_CUDA(cudaMemsetAsync(ctx->assoc_xd.p, 0, (size_t)nk * ncolp * 32, ctx->stream));
    {
        const int64_t n = S_pad * (K + 1);
        assoc_digits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
            (const double*)ctx->covars.p, (const double*)ctx->outcome.p, d_row_of_sample, S, S_pad, K, ncolp,
            (const double*)ctx->assoc_colscale.p, (int8_t*)ctx->assoc_xd.p);
        TRT_KERNEL_CHECK();
    }
    CUtensorMap tmap;
    {
        const cuuint64_t gdim[2] = {(cuuint64_t)S * 3, (cuuint64_t)L};
        const cuuint64_t gstr[1] = {(cuuint64_t)ctx->gt_active_pitch};
        const cuuint32_t box[2] = {96u, (cuuint32_t)kMLoci};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, (void*)ctx->d_gt_active, gdim, gstr, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS)
            return trt_set_error(ctx, TRT_ECUDA, "cuTensorMapEncodeTiled failed (%d): GT base %p pitch %zu", (int)r,
                                 (const void*)ctx->d_gt_active, ctx->gt_active_pitch);
    }
    MmaParams mp;
    mp.L = L; mp.S = S;
    mp.n_tiles = n_tiles; mp.nk = nk; mp.nseg = nseg; mp.ks_per_seg = kps; mp.nwin = nwin; mp.stages = 0;
    mp.ncolp = ncolp; mp.ind_col = K * kMJ;
    mp.tab = tab;
    mp.xd = (const int8_t*)ctx->assoc_xd.p;
    mp.part = (int32_t*)ctx->assoc_mma_part.p;
    mp.masks = (uint32_t*)ctx->assoc_mma_masks.p;
    const int grid = (int)std::min<int64_t>((int64_t)n_tiles * nseg, sms);
    switch (nt) {
#define CASE(NN) case NN: TRT_TRY(launch_mma<NN>(ctx, tmap, mp, grid)); break;
        CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14) CASE(15)
#undef CASE
        default: return trt_set_error(ctx, TRT_EINVAL, "assoc mma path: %d digit column groups", nt);
    }
    {
        const int64_t n = L * (K + 1);
        assoc_mma_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
            mp, K, (const uint8_t*)ctx->assoc_flags.p, scale, (const double*)ctx->assoc_colscale.p, mom);
        TRT_KERNEL_CHECK();
    }
    // down-dates from the masks (the FP64 z-rows in sample order are shared with the tile path)
    TRT_TRY(trt_assoc_downdate(ctx, d_row_of_sample, mp.masks, n_tiles, nk, nwin, kMLoci, 32, (const uint8_t*)ctx->assoc_flags.p, 1, dd));
    TRT_CUDA(cudaMemcpyAsync(n_fp64, d_nfp64, 4, cudaMemcpyDeviceToHost, ctx->stream));
    TRT_CUDA(cudaStreamSynchronize(ctx->stream));
    return TRT_OK;
}
