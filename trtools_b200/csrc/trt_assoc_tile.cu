// K4 FP64 tile path — associaTR moments with one THREAD per locus.  Since round 2 this is the FALLBACK of the tensor path
// (trt_assoc_mma.cu) for loci outside its integer form (flags[l] == 2); this file also holds the down-date kernel and the
// z-row table both paths share.
//
// The per-locus regression needs, over the called design samples of the locus (associaTR.py:246-291 in the
// reference tree), n, sum g, sum g^2, g.y and g.c_k for every covariate column: a skinny FP64 contraction
// G[loci x samples] . Z[samples x K].  Mapping that keeps every operand cheap:
//   * a CTA owns a tile of 256 x NL consecutive loci x a segment of the sample axis; thread t <-> loci t, t + 256, ..
//     of the tile (NL = 2: the z-row fetched from shared memory is reused by both loci, which is what bounds the
//     kernel — measured 19 shared-memory wavefronts per call with NL = 1, 14 of them the z-row);
//   * a producer warp streams the tile through a shared-memory ring: 2-D TMA tensor copies (256 rows each) bring
//     24 samples of native cyvcf2 GT per locus (144 B per row: an odd multiple of 16 B, so the 32 lanes of a
//     warp read their own rows with conflict-free LDS.128), one 1-D bulk copy brings the 24 z-rows
//     (y, c_1..c_{K-1}, in-design flag) of the same samples;
//   * every lane of a warp is at the SAME sample at the same time, so the z-row is a broadcast LDS.128 stream
//     and the K+2 DFMAs per call have all operands in registers;
//   * allele length lookups go to a thread-private column of a [16][256 NL] FP64 table (2-wavefront LDS.64,
//     never a bank conflict whatever alleles the lanes carry);
//   * the accumulators live in registers for the whole segment, so there is no cross-thread reduction at all.
// Uncalled design samples (the rows whose outer products must be removed from C'C, C'y, y'y for this locus)
// are emitted as one 24-bit mask per (locus, chunk) (32-bit words from the tensor path); assoc_downdate_mask_kernel turns
// the masks into the exact down-dates without reading GT again (FP64 mma.sync Gram updates).
//
// Algorithmic traffic: 6 B/call of GT once (+ 0.2 B/call of masks written and read back).
#include <cuda.h>
#include <math.h>

#include <algorithm>

#include "trt_assoc_tile.cuh"

namespace {

constexpr int kNL = kAssocLociPerThread;        // loci per consumer thread
constexpr int kTCons = 256;                     // consumer threads
constexpr int kTLoci = kAssocTileLoci;          // 256 x NL loci per tile
constexpr int kTWarps = kTCons / 32;            // 8 consumer warps
constexpr int kTThreads = kTCons + 32;          // + producer warp
constexpr int kTChunk = kAssocChunk;            // samples per stage
constexpr int kTRowBytes = kTChunk * 6;         // 144 B = 9 x 16 B (NL = 2) / 240 B = 15 x 16 B (NL = 1)
constexpr int kTGtBytes = kTLoci * kTRowBytes;  // bytes of GT per stage
static_assert((kTRowBytes / 16) % 2 == 1 && kTRowBytes % 48 == 0, "rows: odd multiple of 16 B, whole 8-call groups");
static_assert(kTLoci == kTCons * kNL, "tile = consumer threads x loci per thread");
constexpr int kTMaxD = kAssocFastMaxAlleles + 2;   // digits: pad, no-call, alleles
constexpr int kTMaxStages = 4;

struct TileParams {
    int64_t L, S;
    const int32_t* locus_off;
    const double* allele_len;
    int n_tiles;                    // ceil(L / tile loci)
    int nchunks;                    // ceil(S / chunk)
    int nseg, chunks_per_seg;       // sample-axis segments (work unit = tile x segment)
    const double* zt;               // [nchunks * 40][ZW] z-rows in VCF sample order (zeros outside the design)
    double* mom_part;               // [nseg][L][K + 3]
    uint32_t* masks;                // [n_tiles][nwin][tile loci][32]: word (c & 31) of window c >> 5 of a locus holds chunk c's mask
                                    // (bit i = sample i of the chunk is an uncalled design sample): the tile kernel writes one
                                    // word per (locus, chunk), the down-date kernel reads a locus' window as ONE 128-byte line
    int nwin;                       // ceil(nchunks / 32)
    int stages;
    const uint8_t* flags;           // per locus path flag (null: every locus with <= kAssocFastMaxAlleles alleles is ours)
    int want;                       // ... the flag value of the loci to process
    int mask_tile_loci, mask_bits;  // geometry of `masks` for the down-date kernel: loci per tile, samples per word
};

__device__ __forceinline__ bool tile_owns(const TileParams& p, int64_t l) {
    if (p.locus_off[l + 1] - p.locus_off[l] > kAssocFastMaxAlleles) return false;       // generic-path locus
    return !p.flags || p.flags[l] == p.want;
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int x, int y, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}

// K = design columns (genotype + intercept + covariates); z-row = [y, c_1 .. c_{K-1}, ind, (pad)] = ZW doubles
template <int K>
__global__ void __launch_bounds__(kTThreads, 1) assoc_tile_kernel(const __grid_constant__ CUtensorMap tmap, TileParams p) {
    constexpr int ZW = (K + 2) & ~1;            // even: rows stay 16-byte aligned
    constexpr int NZ2 = ZW / 2;
    constexpr int kZBytes = kTChunk * ZW * 8;
    extern __shared__ __align__(128) unsigned char smem[];
    const int stages = p.stages;
    unsigned char* gt_ring = smem;                                            // [stages][kTGtBytes]
    unsigned char* z_ring = smem + (size_t)stages * kTGtBytes;                // [stages][kZBytes]
    double* table = (double*)(z_ring + (size_t)stages * kZBytes);             // [kTMaxD][kTLoci]
    uint64_t* full = (uint64_t*)(table + kTMaxD * kTLoci);
    uint64_t* empty = full + kTMaxStages;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < stages; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kTWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();
    const int n_units = p.n_tiles * p.nseg;

    if (warp == kTWarps) {
        // ===== producer: one elected lane feeds the ring across work units =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
                const int tile = u / p.nseg;
                const int seg = u % p.nseg;
                const int c0 = seg * p.chunks_per_seg, c1 = min(p.nchunks, c0 + p.chunks_per_seg);
                for (int c = c0; c < c1; c++) {
                    mbar_wait(&empty[stage], phase ^ 1u);
                    mbar_arrive_expect_tx(&full[stage], (uint32_t)(kTGtBytes + kZBytes));
#pragma unroll
                    for (int h = 0; h < kNL; h++)
                        tma_load_2d(gt_ring + (size_t)stage * kTGtBytes + (size_t)h * kTCons * kTRowBytes, &tmap, c * (kTChunk * 3),
                                    tile * kTLoci + h * kTCons, &full[stage]);
                    tma_load_1d(z_ring + (size_t)stage * kZBytes, p.zt + (size_t)c * kTChunk * ZW, kZBytes, &full[stage]);
                    if (++stage == stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
        return;
    }

    // ===== consumers: thread <-> loci tid, tid + 256, .. of the tile =====
    int stage = 0;
    uint32_t phase = 0;
    double* mytab = table + tid;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        const int tile = u / p.nseg;
        const int seg = u % p.nseg;
        const int c0 = seg * p.chunks_per_seg, c1 = min(p.nchunks, c0 + p.chunks_per_seg);
        bool valid[kNL];
        unsigned D[kNL];
#pragma unroll
        for (int h = 0; h < kNL; h++) {
            // thread-private length table by digit d = allele + 2: pad -> -2 - ref, no-call -> NaN, allele -> len - ref
            // (the per-haplotype shift by the reference length is the conditioning shift of the generic kernel);
            // g = tab[da] + tab[db] is NaN exactly when a haplotype is missing or out of range: "called" = (g == g)
            const int64_t l = (int64_t)tile * kTLoci + h * kTCons + tid;
            double ref = 0.0;
            int a0 = 0, A = 0;
            if (l < p.L) {
                a0 = p.locus_off[l];
                A = p.locus_off[l + 1] - a0;
            }
            valid[h] = (l < p.L) && A <= kAssocFastMaxAlleles &&     // wider loci go through the generic kernels,
                       (!p.flags || p.flags[l] == p.want);            // integer-form loci through the tensor path
            if (!valid[h]) A = 0;
            if (A > 0) ref = p.allele_len[a0];
            D[h] = (unsigned)A + 2u;
            double* t = mytab + h * kTCons;
            t[0] = valid[h] ? -2.0 - ref : nan("");
            t[kTLoci] = nan("");
            for (int a = 0; a < A; a++) t[(a + 2) * kTLoci] = p.allele_len[a0 + a] - ref;
        }
        double accz[kNL][K], sg[kNL], sgg[kNL];
        int n[kNL];
#pragma unroll
        for (int h = 0; h < kNL; h++) {
#pragma unroll
            for (int k = 0; k < K; k++) accz[h][k] = 0.0;
            sg[h] = sgg[h] = 0.0;
            n[h] = 0;
        }
        uint32_t* mrow = p.masks + ((size_t)tile * p.nwin * kTLoci + tid) * 32;
        for (int c = c0; c < c1; c++) {
            mbar_wait(&full[stage], phase);
            const unsigned char* gts = gt_ring + (size_t)stage * kTGtBytes + (size_t)tid * kTRowBytes;
            const double2* zc = (const double2*)(z_ring + (size_t)stage * kZBytes);
            uint32_t mask[kNL];
            unsigned dep = 0u;      // one word of every shared-memory load of the chunk's LAST sample (see the arrive below)
#pragma unroll
            for (int h = 0; h < kNL; h++) mask[h] = 0u;
#pragma unroll
            for (int grp = 0; grp < kTChunk / 8; grp++) {
                uint32_t w[kNL][12];
#pragma unroll
                for (int h = 0; h < kNL; h++) {
                    const uint4* row = (const uint4*)(gts + (size_t)h * kTCons * kTRowBytes);
                    const uint4 v0 = row[3 * grp], v1 = row[3 * grp + 1], v2 = row[3 * grp + 2];
                    w[h][0] = v0.x; w[h][1] = v0.y; w[h][2] = v0.z; w[h][3] = v0.w;
                    w[h][4] = v1.x; w[h][5] = v1.y; w[h][6] = v1.z; w[h][7] = v1.w;
                    w[h][8] = v2.x; w[h][9] = v2.y; w[h][10] = v2.z; w[h][11] = v2.w;
                }
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int sidx = grp * 8 + j;
                    const double2* zr = zc + sidx * NZ2;
                    double z[ZW];
#pragma unroll
                    for (int q = 0; q < NZ2; q++) {
                        const double2 t = zr[q];
                        z[2 * q] = t.x;
                        z[2 * q + 1] = t.y;
                    }
                    const bool ind = z[K] != 0.0;
                    if (grp == kTChunk / 8 - 1 && j == 7) {
#pragma unroll
                        for (int q = 0; q < NZ2; q++) dep ^= (unsigned)__double2loint(z[2 * q]);
                    }
                    const int k0 = 3 * j, k1 = k0 + 1;
#pragma unroll
                    for (int h = 0; h < kNL; h++) {
                        const int a = (k0 & 1) ? ((int)w[h][k0 >> 1] >> 16) : (int)(short)(w[h][k0 >> 1] & 0xffffu);
                        const int b = (k1 & 1) ? ((int)w[h][k1 >> 1] >> 16) : (int)(short)(w[h][k1 >> 1] & 0xffffu);
                        const unsigned da = (unsigned)(a + 2), db = (unsigned)(b + 2);
                        const double* t = mytab + h * kTCons;
                        const double la = t[((da < D[h]) ? da : 1u) * kTLoci];
                        const double lb = t[((db < D[h]) ? db : 1u) * kTLoci];
                        if (grp == kTChunk / 8 - 1 && j == 7) dep ^= (unsigned)__double2loint(la) ^ (unsigned)__double2loint(lb);
                        const double gs = la + lb;
                        const bool called = (gs == gs);             // an invalid locus has D = 2 and a NaN pad entry: never called
                        const bool ok = called & ind;
                        const double g = ok ? gs : 0.0;
                        if (ok) n[h]++;
                        if (ind & !called) mask[h] |= 1u << sidx;
                        sg[h] += g;
                        sgg[h] = fma(g, g, sgg[h]);      // explicit fma: the library builds with -fmad=false
#pragma unroll
                        for (int k = 0; k < K; k++) accz[h][k] = fma(g, z[k], accz[h][k]);
                    }
                }
            }
            {
                // Release the stage only after every shared-memory load of the chunk has RETURNED.  Loads return in order,
                // so a vote on a value built from the last sample's loads is enough; it is true unless all 32 lanes hold
                // the constant.  Without the dependency ptxas hoists the arrive above the trailing DFMAs, and the refill
                // (an async-proxy TMA write) is not ordered behind a generic-proxy read that is still in flight — the
                // same hazard that made the GT scan miscount on cold launches (profiles/README.md).
                const bool returned = __any_sync(0xffffffffu, dep != 0x9e3779b9u);
                if (lane == 0 && returned) mbar_arrive(&empty[stage]);
            }
            if (++stage == stages) { stage = 0; phase ^= 1u; }
#pragma unroll
            for (int h = 0; h < kNL; h++)
                mrow[((size_t)(c >> 5) * kTLoci + h * kTCons) * 32 + (c & 31)] = valid[h] ? mask[h] : 0u;
        }
#pragma unroll
        for (int h = 0; h < kNL; h++) {
            if (!valid[h]) continue;
            // mom layout of the solve kernel: n, sum g', sum g'^2, g'.y, g'.c_1 .. g'.c_{K-1}
            const int64_t l = (int64_t)tile * kTLoci + h * kTCons + tid;
            double* o = p.mom_part + ((size_t)seg * p.L + l) * (K + 3);
            o[0] = (double)n[h];
            o[1] = sg[h];
            o[2] = sgg[h];
#pragma unroll
            for (int k = 0; k < K; k++) o[3 + k] = accz[h][k];
        }
    }
}

// z-rows in VCF sample order: [y, c_1 .. c_{K-1}, ind, (pad)]; zeros for samples outside the design and beyond S
__global__ void assoc_ztable_kernel(const double* __restrict__ covars, const double* __restrict__ outcome,
                                    const int32_t* __restrict__ row_of_sample, int64_t S, int64_t S_pad, int K, int ZW,
                                    double* __restrict__ zt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S_pad * ZW) return;
    const int64_t s = i / ZW;
    const int k = (int)(i % ZW);
    const int r = (s < S) ? row_of_sample[s] : -1;
    double v = 0.0;
    if (r >= 0) {
        if (k == 0) v = outcome[r];
        else if (k < K) v = covars[(int64_t)r * K + k];
        else if (k == K) v = 1.0;
    }
    zt[i] = v;
}

__global__ void assoc_reduce_mom_kernel(const double* __restrict__ part, int nseg, int64_t n, double* __restrict__ mom,
                                        TileParams p, int nacc) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t l = i / nacc;
    if (!tile_owns(p, l)) return;
    double s = 0.0;
    for (int g = 0; g < nseg; g++) s += part[(size_t)g * n + i];   // fixed order: bit-reproducible
    mom[i] = s;
}

// exact down-dates from the uncalled-sample masks: warp per locus.  The K x K matrix sum z z' over the uncalled design
// rows (z = (c_1 .. c_{K-1}, y)) is a Gram matrix, accumulated four samples at a time by FP64 mma.sync.m8n8k4: lane
// (g, t) holds z_t[g] (and z_t[g + 8] when K > 8) of the group's four samples — the A fragment of a column block and,
// the same numbers, the B fragment — loaded straight from the z-row table in L2 (eight lanes read 64 consecutive bytes of
// one row), so there is no staging, and the accumulators of a locus are the C fragments: nothing to reduce at the end.
// NB8 = column blocks of 8: K <= 8 -> 1 MMA per group, K <= 16 -> 3 (blocks 00, 01, 11).
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <int NB8>
__global__ void __launch_bounds__(256) assoc_downdate_mask_kernel(TileParams p, int K, int ZW, double* __restrict__ dd) {
    constexpr int kU = 4;                           // sample groups per round trip to L2 (16 samples)
    constexpr int kRound = 4 * kU;
    __shared__ uint32_t lst_all[8][32 * 32 + kRound];   // pending uncalled samples of the locus (absolute sample indices)
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    uint32_t* lst = lst_all[wib];
    const int ne = K * (K + 1) / 2;
    // column j of the Gram matrix (order c_1 .. c_{K-1}, y) sits at index j + 1 (or 0 for y) of a z-row [y, c_1 .. c_{K-1}, ind];
    // lanes whose column does not exist read column 0 and drop it
    const int src_lo = (g < K - 1) ? g + 1 : (g == K - 1 ? 0 : -1);
    const int src_hi = (g + 8 < K - 1) ? g + 9 : (g + 8 == K - 1 ? 0 : -1);
    const double* zlo = p.zt + (src_lo >= 0 ? src_lo : 0);
    const double* zhi = p.zt + (src_hi >= 0 ? src_hi : 0);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t l = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib; l < p.L; l += nwarps) {
        if (!tile_owns(p, l)) continue;
        const int tloci = p.mask_tile_loci, wbits = p.mask_bits;
        const int64_t tile = l / tloci;
        const int tl = (int)(l % tloci);
        const uint32_t* mrow = p.masks + ((size_t)tile * p.nwin * tloci + tl) * 32 + lane;   // + window * tloci * 32
        double c00[2] = {0.0, 0.0}, c01[2] = {0.0, 0.0}, c11[2] = {0.0, 0.0};
        // one round: 16 listed samples (fewer only when the locus is flushed) -> 4 groups of 4 -> up to 12 MMAs
        auto round = [&](int pos, int n) {
            double lo[kU], hi[kU];
#pragma unroll
            for (int u = 0; u < kU; u++) {
                const int e = 4 * u + t;
                const unsigned row = lst[pos + min(e, n - 1)] * (unsigned)ZW;        // always a listed sample
                const double vlo = zlo[row];
                lo[u] = (e < n && src_lo >= 0) ? vlo : 0.0;
                if (NB8 > 1) {
                    const double vhi = zhi[row];
                    hi[u] = (e < n && src_hi >= 0) ? vhi : 0.0;
                }
            }
#pragma unroll
            for (int u = 0; u < kU; u++) {
                if (4 * u < n) {                    // warp-uniform; false only in the flush round
                    dmma884(c00, lo[u], lo[u]);
                    if (NB8 > 1) {
                        dmma884(c01, lo[u], hi[u]);
                        dmma884(c11, hi[u], hi[u]);
                    }
                }
            }
        };
        int pending = 0;                            // listed samples not yet accumulated (< kRound between windows)
        uint32_t m_next = (lane < p.nchunks) ? mrow[0] : 0u;
        for (int cb = 0; cb < p.nchunks; cb += 32) {
            uint32_t m = m_next;
            {   // the next window's masks are requested now: their latency hides under this window's work
                const int cn = cb + 32 + lane;
                m_next = (cn < p.nchunks) ? mrow[(size_t)((cb >> 5) + 1) * tloci * 32] : 0u;
            }
            const int cnt = __popc(m);
            int off = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, off, o);
                if (lane >= o) off += v;
            }
            const int total = __shfl_sync(0xffffffffu, off, 31);
            if (total == 0) continue;
            off += pending - cnt;
            const uint32_t s0 = (uint32_t)(cb + lane) * (uint32_t)wbits;
            while (m) {
                const int b = __ffs((int)m) - 1;
                m &= m - 1;
                lst[off++] = s0 + (uint32_t)b;
            }
            __syncwarp();
            const int have = pending + total;
            int pos = 0;
            for (; pos + kRound <= have; pos += kRound) round(pos, kRound);
            pending = have - pos;
            // the remainder moves to the front of the list for the next window
            uint32_t carry = 0;
            if (lane < pending) carry = lst[pos + lane];
            __syncwarp();
            if (lane < pending) lst[lane] = carry;
            __syncwarp();
        }
        if (pending > 0) round(0, pending);
        __syncwarp();
        // C fragments: rows g, columns 2 t and 2 t + 1 of each block
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const int ca = 2 * t + i;
            if (g <= ca && ca < K) dd[l * ne + (g * K - g * (g - 1) / 2 + (ca - g))] = c00[i];
            if (NB8 > 1) {
                const int cb2 = 8 + ca, gb = 8 + g;
                if (cb2 < K) dd[l * ne + (g * K - g * (g - 1) / 2 + (cb2 - g))] = c01[i];
                if (gb <= cb2 && cb2 < K) dd[l * ne + (gb * K - gb * (gb - 1) / 2 + (cb2 - gb))] = c11[i];
            }
        }
    }
}

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)f;
    }
    return fn;
}

template <int K>
int launch_tile(trt_ctx* ctx, const CUtensorMap& tmap, TileParams& tp, int grid) {
    constexpr int ZW = (K + 2) & ~1;
    constexpr size_t zbytes = (size_t)kTChunk * ZW * 8;
    const size_t fixed = (size_t)kTMaxD * kTLoci * 8 + 2 * kTMaxStages * 8 + 128;
    int stages = (int)(((size_t)ctx->max_smem_optin - fixed) / (kTGtBytes + zbytes));
    stages = std::min(stages, kTMaxStages);
    if (stages < 2) return trt_set_error(ctx, TRT_ENOMEM, "assoc tile kernel: not enough shared memory");
    tp.stages = stages;
    const size_t smem = (size_t)stages * (kTGtBytes + zbytes) + fixed;
    TRT_CUDA(cudaFuncSetAttribute(assoc_tile_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    assoc_tile_kernel<K><<<grid, kTThreads, smem, ctx->stream>>>(tmap, tp);
    TRT_KERNEL_CHECK();
    return TRT_OK;
}

}  // namespace

EncodeTiledFn trt_tmap_encode_fn() { return encode_fn(); }

int trt_assoc_fast_zw(int K) { return (K + 2) & ~1; }

int trt_assoc_downdate(trt_ctx* ctx, const int32_t* d_row_of_sample, const uint32_t* masks, int n_tiles, int nwords, int nwin,
                       int tile_loci, int bits, const uint8_t* flags, int want, double* dd) {
    const int64_t L = ctx->L, S = ctx->S;
    const int K = ctx->K, ZW = trt_assoc_fast_zw(K);
    if (L == 0) return TRT_OK;
    const int64_t S_pad = (int64_t)nwords * bits;
    TRT_TRY(trt_ensure(ctx, ctx->assoc_zt, (size_t)S_pad * ZW * 8 + 64));
    {
        const int64_t n = S_pad * ZW;
        assoc_ztable_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
            (const double*)ctx->covars.p, (const double*)ctx->outcome.p, d_row_of_sample, S, S_pad, K, ZW, (double*)ctx->assoc_zt.p);
        TRT_KERNEL_CHECK();
    }
    TileParams tp;
    memset(&tp, 0, sizeof(tp));
    tp.L = L; tp.S = S;
    tp.locus_off = (const int32_t*)ctx->locus_off.p;
    tp.n_tiles = n_tiles;
    tp.nchunks = nwords;
    tp.nwin = nwin;
    tp.zt = (const double*)ctx->assoc_zt.p;
    tp.masks = const_cast<uint32_t*>(masks);
    tp.flags = flags;
    tp.want = want;
    tp.mask_tile_loci = tile_loci;
    tp.mask_bits = bits;
    const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>((L + 7) / 8, (int64_t)ctx->sm_count * 8));
    if (K <= 8) assoc_downdate_mask_kernel<1><<<blocks, 256, 0, ctx->stream>>>(tp, K, ZW, dd);
    else assoc_downdate_mask_kernel<2><<<blocks, 256, 0, ctx->stream>>>(tp, K, ZW, dd);
    TRT_KERNEL_CHECK();
    return TRT_OK;
}

// Moments (into mom [L][K+3]) and down-dates (into dd [L][K(K+1)/2]) of every locus with at most
// kAssocFastMaxAlleles alleles; the others are left to the generic kernels.
int trt_assoc_fast(trt_ctx* ctx, const int32_t* d_row_of_sample, double* mom, double* dd, const uint8_t* flags) {
    const int64_t L = ctx->L, S = ctx->S;
    const int K = ctx->K, nacc = K + 3, ZW = trt_assoc_fast_zw(K);
    if (L == 0) return TRT_OK;
    EncodeTiledFn enc = encode_fn();
    if (!enc) return trt_set_error(ctx, TRT_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const int n_tiles = (int)((L + kTLoci - 1) / kTLoci);
    const int nchunks = (int)((S + kTChunk - 1) / kTChunk);
    const int64_t S_pad = (int64_t)nchunks * kTChunk;
    // work units = tile x sample segment, sized so the persistent grid's last wave is >= 95% full
    const int sms = ctx->sm_count;
    int nseg = 1;
    for (; nseg < 16; nseg++) {
        const int64_t units = (int64_t)n_tiles * nseg;
        const int64_t waves = (units + sms - 1) / sms;
        if ((double)units / (double)(waves * sms) >= 0.95 || nchunks / (nseg + 1) < 32) break;
    }
    nseg = std::max(1, std::min(nseg, nchunks));
    const int cps = (nchunks + nseg - 1) / nseg;
    nseg = (nchunks + cps - 1) / cps;

    TRT_TRY(trt_ensure(ctx, ctx->assoc_zt, (size_t)S_pad * ZW * 8 + 64));
    const int nwin = (nchunks + 31) / 32;
    TRT_TRY(trt_ensure(ctx, ctx->assoc_masks, (size_t)n_tiles * nwin * kTLoci * 32 * 4 + 64));
    TRT_TRY(trt_ensure(ctx, ctx->assoc_mom_part, (size_t)nseg * L * nacc * 8 + 64));
    {
        const int64_t n = S_pad * ZW;
        assoc_ztable_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
            (const double*)ctx->covars.p, (const double*)ctx->outcome.p, d_row_of_sample, S, S_pad, K, ZW, (double*)ctx->assoc_zt.p);
        TRT_KERNEL_CHECK();
    }
    CUtensorMap tmap;
    {
        const cuuint64_t gdim[2] = {(cuuint64_t)S * 3, (cuuint64_t)L};
        const cuuint64_t gstr[1] = {(cuuint64_t)ctx->gt_active_pitch};
        const cuuint32_t box[2] = {(cuuint32_t)kTChunk * 3, (cuuint32_t)kTCons};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, (void*)ctx->d_gt_active, gdim, gstr, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS)
            return trt_set_error(ctx, TRT_ECUDA, "cuTensorMapEncodeTiled failed (%d): GT base %p pitch %zu", (int)r,
                                 (const void*)ctx->d_gt_active, ctx->gt_active_pitch);
    }
    TileParams tp;
    tp.L = L; tp.S = S;
    tp.locus_off = (const int32_t*)ctx->locus_off.p;
    tp.allele_len = (const double*)ctx->allele_len.p;
    tp.n_tiles = n_tiles;
    tp.nchunks = nchunks;
    tp.nwin = nwin;
    tp.nseg = nseg;
    tp.chunks_per_seg = cps;
    tp.zt = (const double*)ctx->assoc_zt.p;
    tp.mom_part = (double*)ctx->assoc_mom_part.p;
    tp.masks = (uint32_t*)ctx->assoc_masks.p;
    tp.stages = 0;
    tp.flags = flags;
    tp.want = 2;
    tp.mask_tile_loci = kTLoci;
    tp.mask_bits = kTChunk;
    const int grid = (int)std::min<int64_t>((int64_t)n_tiles * nseg, sms);
    switch (K) {
#define CASE(KK) case KK: TRT_TRY(launch_tile<KK>(ctx, tmap, tp, grid)); break;
        CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14) CASE(15) CASE(16)
#undef CASE
        default: return trt_set_error(ctx, TRT_EINVAL, "assoc fast path supports 2 <= K <= %d", kAssocFastMaxK);
    }
    {
        const int64_t n = L * nacc;
        assoc_reduce_mom_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
            (const double*)ctx->assoc_mom_part.p, nseg, n, mom, tp, nacc);
        TRT_KERNEL_CHECK();
    }
    // the z-rows table built above covers nchunks * kTChunk samples: the down-date call rebuilds the same table
    TRT_TRY(trt_assoc_downdate(ctx, d_row_of_sample, tp.masks, n_tiles, nchunks, nwin, kTLoci, kTChunk, flags, 2, dd));
    return TRT_OK;
}
