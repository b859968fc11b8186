// K3 — dumpSTR: call-level filter operators + ApplyCallFilters bookkeeping, and locus-level filters.
//
// call_filter_kernel: the grid tiles [loci chunk] x [sample slab]; a thread owns 8 CONSECUTIVE samples
// (48 B of GT = 3 x 16 B vector loads, 32 B per int32/float32 FORMAT field = 2 x 16 B) and walks down
// the loci of its chunk, so the per-sample accumulators of dumpSTR's sample log (numcalls, totaldp,
// one counter per filter) live in registers / thread-private shared memory and are flushed with one
// atomic per (sample, counter) per chunk instead of one per call.  It writes the masked genotypes
// (filtered call -> all haplotypes -1, phase 0) that every later statistic reads, and a per-call
// bitmask of the filters that fired.
//
// Reference semantics reproduced (file:line in the reference tree):
//   filters.CallFilterMinValue/MaxValue           trtools/dumpSTR/filters.py:327-409
//   filters.HipSTRCallFlankIndels / CallStutter   :415-484   (int32/int32 -> float64 ratio)
//   filters.GangSTRCallExpansionProb{Hom,Het,Total} :573-674 (float32 compare, called samples only)
//   dumpSTR.ApplyCallFilters                      trtools/dumpSTR/dumpSTR.py:613-774
//   filters.Filter_MinLocusCallrate/HWEP/Het/MaxLocusHet/LocusHrun  filters.py:35-217
//   dumpSTR.ApplyLocusFilters :917-973 and the INFO recompute :1307-1336
#include <math.h>

#include <algorithm>
#include <cmath>

#include "trt_internal.cuh"
#include "trt_scan.cuh"


namespace {

constexpr int kSlab = 8;              // samples per thread
constexpr int kCfThreads = 256;
constexpr int kSlabSamples = kSlab * kCfThreads;   // 2048 samples per CTA
constexpr int kMaxSpecs = TRT_MAX_CALL_FILTERS;

struct CfSpec {
    int kind;
    int field;      // TRT_FMT_*
    int is_float;   // field holds float32
    double thr;
    float thr_f32;  // threshold rounded to float32 (numpy weak-scalar comparison for float32 arrays)
    int variant;    // CFV_*: the comparison the TMA kernel runs for this filter (kind x dtype, resolved on the host)
    long long thr_i64;   // int32 fields: value < thr <=> value < ceil(thr); value > thr <=> value > floor(thr)
    int thr_i32;         // the same threshold when it lies inside the int32 range (else the variant is ALWAYS / NEVER)
    int ratio_slot;      // CFV_RATIO_TABLE: which cut table of CfTmaParams::ratio_cut this filter reads
    int hit_missing;     // CFV_RATIO_TABLE: INT32_MIN / INT32_MIN = 1.0 > thr  (a call whose depth fields are both '.')
};
enum { CFV_NEVER = 0, CFV_MIN_I32, CFV_MAX_I32, CFV_MIN_F32, CFV_MAX_F32, CFV_RATIO_TABLE, CFV_RATIO_EXACT, CFV_HOST, CFV_ALWAYS };
constexpr int kRatioTable = 256;     // depths 0 .. 255 decide field/DP > thr with one table lookup

struct CfParams {
    const int16_t* gt;
    size_t pitch;
    int16_t* gt_out;       // same pitch
    int64_t L, S;
    const void* fmt[TRT_FMT_NFIELDS];
    int fmt_ncol[TRT_FMT_NFIELDS];
    int n_specs;
    CfSpec specs[kMaxSpecs];
    int dp_field;          // field used for totaldp (-1: none)
    int dp_is_float;       // e.g. ExpansionHunter LC: float32 coverage (generic kernel only)
    uint32_t* call_mask;   // [L][S] or null
    double* trig;          // [n_specs][L][S] or null
    long long* filter_counts;  // [n_specs][S]
    long long* numcalls;       // [S]
    long long* dpsum;          // [S]
    unsigned int* dp_poison;   // [S] non-zero: a PASS call had a missing DP
    int* neg_dp_locus;         // min locus index with a PASS call of negative DP
    int loci_per_chunk;
};

__device__ __forceinline__ void load8_i32(const int32_t* base, int64_t s0, int64_t S, int32_t (&v)[8], bool aligned) {
    if (aligned && s0 + 8 <= S) {
        const int4 a = *reinterpret_cast<const int4*>(base + s0);
        const int4 b = *reinterpret_cast<const int4*>(base + s0 + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = (s0 + j < S) ? base[s0 + j] : 0;
    }
}

__global__ void __launch_bounds__(kCfThreads, 2) call_filter_kernel(CfParams p) {
    extern __shared__ unsigned int fcnt[];   // [n_specs][2048] thread-private columns (no conflicts)
    const int tid = threadIdx.x;
    const int64_t s0 = (int64_t)blockIdx.x * kSlabSamples + (int64_t)tid * kSlab;
    const int64_t l_begin = (int64_t)blockIdx.y * p.loci_per_chunk;
    const int64_t l_end = min(p.L, l_begin + p.loci_per_chunk);
    for (int f = 0; f < p.n_specs; f++)
#pragma unroll
        for (int j = 0; j < kSlab; j++) fcnt[f * kSlabSamples + tid * kSlab + j] = 0;
    int ncalls[kSlab];
    long long dps[kSlab];
    unsigned int poison = 0;
#pragma unroll
    for (int j = 0; j < kSlab; j++) { ncalls[j] = 0; dps[j] = 0; }
    const bool live_thread = s0 < p.S;
    // rows of int32/float32 FORMAT arrays are S*4 bytes: 16-byte aligned vector loads need S % 4 == 0
    const bool fmt_aligned = (p.S % 4) == 0;
    const bool gt_vec = (s0 + kSlab <= p.S);

    if (live_thread) {
        for (int64_t l = l_begin; l < l_end; l++) {
            // ---- GT: 8 calls = 24 int16 ------------------------------------------------------------
            const int16_t* row = (const int16_t*)((const char*)p.gt + (size_t)l * p.pitch);
            int16_t h[24];
            if (gt_vec) {
                const uint4* src = reinterpret_cast<const uint4*>(row + s0 * 3);
                uint4 v0 = src[0], v1 = src[1], v2 = src[2];
                const uint32_t w[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
#pragma unroll
                for (int k = 0; k < 24; k++) h[k] = (int16_t)((k & 1) ? (w[k >> 1] >> 16) : (w[k >> 1] & 0xffffu));
            } else {
#pragma unroll
                for (int k = 0; k < 24; k++) h[k] = (s0 + k / 3 < p.S) ? row[s0 * 3 + k] : (int16_t)-1;
            }
            // ---- FORMAT fields used by the filters ---------------------------------------------------
            int32_t dpv[8];
            bool have_dp = false;
            if (p.dp_field >= 0) {
                load8_i32((const int32_t*)p.fmt[p.dp_field] + l * p.S, s0, p.S, dpv, fmt_aligned);
                have_dp = true;
            }
            uint32_t fired[8];
            bool nocall[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                fired[j] = 0;
                nocall[j] = (h[3 * j] == -1) | (h[3 * j + 1] == -1);
            }
            for (int f = 0; f < p.n_specs; f++) {
                const CfSpec sp = p.specs[f];
                int32_t raw[8], den[8];
                if (sp.kind <= TRT_CF_RATIO_GT || sp.kind == TRT_CF_HOST_VALUE)
                    load8_i32((const int32_t*)p.fmt[sp.field] + l * p.S, s0, p.S, raw, fmt_aligned);
                if (sp.kind == TRT_CF_RATIO_GT)
                    load8_i32((const int32_t*)p.fmt[TRT_FMT_DP] + l * p.S, s0, p.S, den, fmt_aligned);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const bool in = s0 + j < p.S;
                    bool hit;
                    double val;
                    if (sp.kind == TRT_CF_MIN || sp.kind == TRT_CF_MAX) {
                        if (sp.is_float) {
                            const float x = __int_as_float(raw[j]);
                            hit = (sp.kind == TRT_CF_MIN) ? (x < sp.thr_f32) : (x > sp.thr_f32);
                            val = (double)x;
                        } else {
                            val = (double)raw[j];
                            hit = (sp.kind == TRT_CF_MIN) ? (val < sp.thr) : (val > sp.thr);
                        }
                    } else if (sp.kind == TRT_CF_RATIO_GT) {
                        val = (double)raw[j] / (double)den[j];   // numpy int32/int32 -> float64
                        hit = val > sp.thr;
                    } else if (sp.kind == TRT_CF_HOST_VALUE) {
                        const float x = __int_as_float(raw[j]);
                        hit = !isnan(x);
                        val = (double)x;
                    } else {   // GangSTR QEXP, float32 [L][S][3], called samples only
                        float x = 0.f;
                        if (in) {
                            const float* q = (const float*)p.fmt[TRT_FMT_QEXP] + (l * p.S + s0 + j) * 3;
                            x = sp.kind == TRT_CF_QEXP_HET ? q[1] : (sp.kind == TRT_CF_QEXP_HOM ? q[2] : __fadd_rn(q[1], q[2]));
                        }
                        hit = !nocall[j] && (x < sp.thr_f32);
                        val = (double)x;
                    }
                    if (in && hit) {
                        fired[j] |= 1u << f;
                        if (!nocall[j]) fcnt[f * kSlabSamples + tid * kSlab + j] += 1;
                    }
                    if (p.trig && in) p.trig[((size_t)f * p.L + l) * p.S + s0 + j] = hit ? val : nan("");
                }
            }
            // ---- bookkeeping + masked genotypes ---------------------------------------------------------
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (s0 + j >= p.S) continue;
                const bool pass = (fired[j] == 0) && !nocall[j];
                if (pass) {
                    ncalls[j]++;
                    if (have_dp) {
                        const int d = dpv[j];
                        if (d == INT_MIN) poison |= 1u << j;
                        else if (d < 0) atomicMin(p.neg_dp_locus, (int)l);
                        else dps[j] += d;
                    }
                }
                if (fired[j] != 0 && !nocall[j]) {   // filtered call: every haplotype -> '.', unphased
                    h[3 * j] = -1;
                    h[3 * j + 1] = -1;
                    h[3 * j + 2] = 0;
                }
                if (p.call_mask) p.call_mask[(size_t)l * p.S + s0 + j] = fired[j] | (nocall[j] ? 0x80000000u : 0u);
            }
            int16_t* orow = (int16_t*)((char*)p.gt_out + (size_t)l * p.pitch);
            if (gt_vec) {
                uint32_t w[12];
#pragma unroll
                for (int k = 0; k < 12; k++) w[k] = (uint32_t)(uint16_t)h[2 * k] | ((uint32_t)(uint16_t)h[2 * k + 1] << 16);
                uint4* dst = reinterpret_cast<uint4*>(orow + s0 * 3);
                dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
                dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
                dst[2] = make_uint4(w[8], w[9], w[10], w[11]);
            } else {
#pragma unroll
                for (int k = 0; k < 24; k++)
                    if (s0 + k / 3 < p.S) orow[s0 * 3 + k] = h[k];
            }
        }
        // ---- flush the per-sample accumulators of this (chunk, slab) -----------------------------------
#pragma unroll
        for (int j = 0; j < kSlab; j++) {
            if (s0 + j >= p.S) continue;
            if (ncalls[j]) atomicAdd((unsigned long long*)&p.numcalls[s0 + j], (unsigned long long)ncalls[j]);
            if (dps[j]) atomicAdd((unsigned long long*)&p.dpsum[s0 + j], (unsigned long long)dps[j]);
            if ((poison >> j) & 1u) atomicOr(&p.dp_poison[s0 + j], 1u);
            for (int f = 0; f < p.n_specs; f++) {
                const unsigned int c = fcnt[f * kSlabSamples + tid * kSlab + j];
                if (c) atomicAdd((unsigned long long*)&p.filter_counts[(size_t)f * p.S + s0 + j], (unsigned long long)c);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// TMA-pipelined variant (diploid, int32/float32 [L][S] fields, at most kTmaMaxFields distinct fields).
// Work item = (2048-sample slab, chunk of loci); persistent CTAs walk their items locus by locus.  A producer
// warp brings the slab of one locus — GT (12 KB) and every FORMAT field the filters read (8 KB each) — into a
// shared-memory ring with 1-D bulk TMA copies; 256 consumer threads own 8 consecutive samples each (per-sample
// counters in registers / thread-private shared-memory columns), patch filtered calls IN PLACE in the ring stage and
// signal a per-stage "done" mbarrier; a store warp then writes the stage's GT back to the masked-genotype tensor
// with a bulk TMA store (contiguous 12 KB instead of 48-byte-strided 16-byte stores) and recycles the stage.  No
// CTA-wide barrier sits in the loop: consumer warps run ahead of each other by up to the ring depth.
// ---------------------------------------------------------------------------------------------------
constexpr int kTmaMaxFields = 4;
constexpr int kTmaMaxStages = 8;
constexpr int kTmaGtBytes = kSlabSamples * 6;       // 12288
constexpr int kTmaFieldBytes = kSlabSamples * 4;    // 8192

struct CfTmaParams {
    CfParams base;
    int n_fields;                       // distinct fields staged per locus
    int field_of_slot[kTmaMaxFields];   // TRT_FMT_* of each staged slot
    int slot_of_spec[kMaxSpecs];        // slot the filter reads (numerator for RATIO_GT)
    int dp_slot;                        // slot of TRT_FMT_DP for ratio filters (-1: none)
    int acc_slot;                       // slot of the depth field summed into totaldp (-1: none)
    int n_ratio;                        // ratio filters with a cut table
    const int32_t* ratio_cut;           // [n_ratio][kRatioTable]: field / depth > thr  <=>  field > cut[depth]  (0 <= depth < 256)
    int stages;
    int loci_per_item;
    int64_t n_slabs, n_items;
};

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_store_1d(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

constexpr int kTS = 4;                               // samples per consumer thread
constexpr int kTmaCons = kSlabSamples / kTS;         // 512 consumer threads: 4 warps per scheduler hide the dependent chains
template <bool WANT_MASK>
__global__ void __launch_bounds__(kTmaCons + 64, 1) call_filter_tma_kernel(CfTmaParams q) {
    extern __shared__ __align__(128) unsigned char smem[];
    const CfParams& p = q.base;
    const int stages = q.stages;
    const size_t stage_bytes = (size_t)kTmaGtBytes + (size_t)q.n_fields * kTmaFieldBytes;
    unsigned char* ring = smem;
    // [n_specs][512] thread-private words: four 8-bit counters (one per call of the thread) of the calls each filter
    // removed in the current item; then the ratio filters' cut tables
    unsigned int* fcnt = (unsigned int*)(smem + (size_t)stages * stage_bytes);
    int* cut_tab = (int*)(fcnt + (size_t)max(p.n_specs, 1) * kTmaCons);          // [n_ratio][kRatioTable]
    uint64_t* full = (uint64_t*)(cut_tab + (size_t)q.n_ratio * kRatioTable + ((q.n_ratio * kRatioTable) & 1));
    uint64_t* empty = full + kTmaMaxStages;
    uint64_t* done = empty + kTmaMaxStages;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < stages; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
            mbar_init(&done[s], kTmaCons / 32);
        }
        mbar_fence_init();
    }
    for (int i = tid; i < q.n_ratio * kRatioTable; i += kTmaCons + 64) cut_tab[i] = q.ratio_cut[i];
    __syncthreads();

    if (warp == kTmaCons / 32 + 1) {
        // ===== store warp: masked GT of every finished stage back to HBM, then the stage is free again =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int pending_stage = -1;      // stage whose GT is still being read by the previous bulk store
            for (int64_t item = blockIdx.x; item < q.n_items; item += gridDim.x) {
                const int64_t slab = item % q.n_slabs, chunk = item / q.n_slabs;
                const int64_t s0 = slab * kSlabSamples;
                const int64_t ns = min((int64_t)kSlabSamples, p.S - s0);
                const uint32_t gt_bytes = (uint32_t)((ns * 6 + 15) & ~int64_t(15));
                const int64_t l0 = chunk * q.loci_per_item, l1 = min(p.L, l0 + q.loci_per_item);
                for (int64_t l = l0; l < l1; l++) {
                    mbar_wait(&done[stage], phase);
                    tma_store_1d((char*)p.gt_out + (size_t)l * p.pitch + (size_t)s0 * 6, ring + (size_t)stage * stage_bytes, gt_bytes);
                    tma_store_commit();
                    if (pending_stage >= 0) {
                        tma_store_wait_read<1>();        // the PREVIOUS store has finished reading its stage
                        mbar_arrive(&empty[pending_stage]);
                    }
                    pending_stage = stage;
                    if (++stage == stages) { stage = 0; phase ^= 1u; }
                }
            }
            if (pending_stage >= 0) {
                tma_store_wait_read<0>();
                mbar_arrive(&empty[pending_stage]);
            }
        }
        return;
    }
    if (warp == kTmaCons / 32) {
        // ===== producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t item = blockIdx.x; item < q.n_items; item += gridDim.x) {
                const int64_t slab = item % q.n_slabs, chunk = item / q.n_slabs;
                const int64_t s0 = slab * kSlabSamples;
                const int64_t ns = min((int64_t)kSlabSamples, p.S - s0);
                const uint32_t gt_bytes = (uint32_t)((ns * 6 + 15) & ~int64_t(15));
                const uint32_t f_bytes = (uint32_t)(ns * 4);                 // S % 4 == 0 on this path
                const int64_t l0 = chunk * q.loci_per_item, l1 = min(p.L, l0 + q.loci_per_item);
                for (int64_t l = l0; l < l1; l++) {
                    mbar_wait(&empty[stage], phase ^ 1u);
                    unsigned char* dst = ring + (size_t)stage * stage_bytes;
                    mbar_arrive_expect_tx(&full[stage], gt_bytes + (uint32_t)q.n_fields * f_bytes);
                    tma_load_1d(dst, (const char*)p.gt + (size_t)l * p.pitch + (size_t)s0 * 6, gt_bytes, &full[stage]);
                    for (int k = 0; k < q.n_fields; k++)
                        tma_load_1d(dst + kTmaGtBytes + (size_t)k * kTmaFieldBytes,
                                    (const char*)p.fmt[q.field_of_slot[k]] + ((size_t)l * p.S + s0) * 4, f_bytes, &full[stage]);
                    if (++stage == stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
        return;
    }

    // ===== consumers: thread <-> kTS consecutive samples of the slab =====
    // Bit j of every mask below = call j of this thread.  Everything per call is branch-free; the per-sample counters of
    // the sample log are 8-bit fields of packed words (an item is at most 255 loci), spread from the bit masks with one
    // multiply:  (m * 0x00204081) & 0x01010101  puts bit j of m into byte j.
    int stage = 0;
    uint32_t phase = 0;
    constexpr uint32_t kAll = (1u << kTS) - 1u;
    for (int64_t item = blockIdx.x; item < q.n_items; item += gridDim.x) {
        const int64_t slab = item % q.n_slabs, chunk = item / q.n_slabs;
        const int64_t s0 = slab * kSlabSamples;
        const int64_t l0 = chunk * q.loci_per_item, l1 = min(p.L, l0 + q.loci_per_item);
        const int64_t sb = s0 + (int64_t)tid * kTS;          // first sample of this thread
        const uint32_t valid_mask = (sb + kTS <= p.S) ? kAll : (sb >= p.S ? 0u : ((1u << (int)(p.S - sb)) - 1u));
        for (int f = 0; f < p.n_specs; f++) fcnt[f * kTmaCons + tid] = 0u;
        uint32_t ncalls4 = 0u;               // packed: calls of each sample that passed every filter
        uint32_t dps[kTS];                   // depth of the passing calls (an item cannot overflow 32 bits: see `big`)
        unsigned int poison = 0;
#pragma unroll
        for (int j = 0; j < kTS; j++) dps[j] = 0u;

        for (int64_t l = l0; l < l1; l++) {
            mbar_wait(&full[stage], phase);
            unsigned char* st = ring + (size_t)stage * stage_bytes;
            // 4 calls = 12 int16 = 24 B (8-byte aligned; 24-byte lane stride is conflict-free for 64-bit accesses)
            uint2* gsrc = reinterpret_cast<uint2*>(st + (size_t)tid * (kTS * 6));
            const uint2 v0 = gsrc[0], v1 = gsrc[1], v2 = gsrc[2];
            uint32_t w[6] = {v0.x, v0.y, v1.x, v1.y, v2.x, v2.y};
            // no-call = some haplotype is -1.  Both haplotypes of a call as one packed int16 pair (even calls sit in one
            // word, odd calls straddle two), digits min(h + 2, 2) = {0 pad, 1 no-call, 2 allele} by one VIADDMNMX.U16x2
            const uint32_t pr[kTS] = {w[0], __byte_perm(w[1], w[2], 0x5432), w[3], __byte_perm(w[4], w[5], 0x5432)};
            uint32_t nocall = 0;
#pragma unroll
            for (int j = 0; j < kTS; j++)
                nocall |= (__vminu2(__vadd2(pr[j], 0x00020002u), 0x00020002u) & 0x00010001u) ? (1u << j) : 0u;
            uint32_t fired_any = 0;
            uint32_t fired[kTS];
            if (WANT_MASK) {
#pragma unroll
                for (int j = 0; j < kTS; j++) fired[j] = 0;
            }
            int32_t deni[kTS] = {0, 0, 0, 0};
            if (q.dp_slot >= 0) {
                const int4 a = *reinterpret_cast<const int4*>(st + kTmaGtBytes + (size_t)q.dp_slot * kTmaFieldBytes + (size_t)tid * 16);
                deni[0] = a.x; deni[1] = a.y; deni[2] = a.z; deni[3] = a.w;
            }
            for (int f = 0; f < p.n_specs; f++) {
                const int variant = p.specs[f].variant;
                const int4 a = *reinterpret_cast<const int4*>(st + kTmaGtBytes + (size_t)q.slot_of_spec[f] * kTmaFieldBytes + (size_t)tid * 16);
                const int32_t raw[kTS] = {a.x, a.y, a.z, a.w};
                uint32_t hit = 0;
                switch (variant) {      // uniform across the grid
                    case CFV_MIN_I32: {
                        const int t = p.specs[f].thr_i32;
#pragma unroll
                        for (int j = 0; j < kTS; j++) hit |= (raw[j] < t) ? (1u << j) : 0u;
                    } break;
                    case CFV_MAX_I32: {
                        const int t = p.specs[f].thr_i32;
#pragma unroll
                        for (int j = 0; j < kTS; j++) hit |= (raw[j] > t) ? (1u << j) : 0u;
                    } break;
                    case CFV_MIN_F32: {
                        const float t = p.specs[f].thr_f32;
#pragma unroll
                        for (int j = 0; j < kTS; j++) hit |= (__int_as_float(raw[j]) < t) ? (1u << j) : 0u;
                    } break;
                    case CFV_MAX_F32: {
                        const float t = p.specs[f].thr_f32;
#pragma unroll
                        for (int j = 0; j < kTS; j++) hit |= (__int_as_float(raw[j]) > t) ? (1u << j) : 0u;
                    } break;
                    case CFV_RATIO_TABLE: {
                        // numpy: int32 / int32 -> float64, filtered when the quotient > thr.  For depths 0 .. 255 the host has
                        // tabulated, with that very float64 division, the largest numerator that is NOT filtered, so the
                        // test is one shared-memory lookup and an integer compare.  A call whose two depth fields are both
                        // '.' (INT32_MIN / INT32_MIN = 1.0) is resolved on the host as well; anything else (negative or
                        // huge depths) takes the exact division, one rare branch per thread.
                        const int* cut = cut_tab + p.specs[f].ratio_slot * kRatioTable;
                        const uint32_t miss_hit = p.specs[f].hit_missing ? kAll : 0u;
                        uint32_t slow = 0, missing = 0;
#pragma unroll
                        for (int j = 0; j < kTS; j++) {
                            const bool fastp = (unsigned)deni[j] < (unsigned)kRatioTable;
                            const int c = cut[fastp ? deni[j] : 0];
                            hit |= (fastp & (raw[j] > c)) ? (1u << j) : 0u;
                            const bool ms = (deni[j] == INT_MIN) & (raw[j] == INT_MIN);
                            missing |= ms ? (1u << j) : 0u;
                            slow |= (fastp | ms) ? 0u : (1u << j);
                        }
                        hit |= missing & miss_hit;
                        if (slow) {
                            const double t = p.specs[f].thr;
#pragma unroll
                            for (int j = 0; j < kTS; j++) {
                                if (!((slow >> j) & 1u)) continue;
                                if (((double)raw[j] / (double)deni[j]) > t) hit |= 1u << j;
                            }
                        }
                    } break;
                    case CFV_RATIO_EXACT: {
                        const double t = p.specs[f].thr;
#pragma unroll
                        for (int j = 0; j < kTS; j++) hit |= (((double)raw[j] / (double)deni[j]) > t) ? (1u << j) : 0u;
                    } break;
                    case CFV_HOST: {
#pragma unroll
                        for (int j = 0; j < kTS; j++) hit |= !isnan(__int_as_float(raw[j])) ? (1u << j) : 0u;
                    } break;
                    case CFV_ALWAYS: hit = kAll; break;
                    default: break;
                }
                hit &= valid_mask;
                fired_any |= hit;
                fcnt[f * kTmaCons + tid] += ((hit & ~nocall) * 0x00204081u) & 0x01010101u;
                if (WANT_MASK) {
#pragma unroll
                    for (int j = 0; j < kTS; j++) fired[j] |= ((hit >> j) & 1u) << f;
                }
            }
            const uint32_t pass = ~fired_any & ~nocall & valid_mask;
            ncalls4 += (pass * 0x00204081u) & 0x01010101u;
            if (q.acc_slot >= 0) {
                const int4 a = *reinterpret_cast<const int4*>(st + kTmaGtBytes + (size_t)q.acc_slot * kTmaFieldBytes + (size_t)tid * 16);
                const int32_t dpv[kTS] = {a.x, a.y, a.z, a.w};
                int32_t neg = 0;
                uint32_t big = 0;
                uint32_t val[kTS];
#pragma unroll
                for (int j = 0; j < kTS; j++) {
                    const int32_t m = (int32_t)(pass << (31 - j)) >> 31;      // all ones iff call j passed
                    neg |= dpv[j] & m;                                        // sign bit: a passing call with a negative depth
                    val[j] = (uint32_t)(max(dpv[j], 0) & m);
                    big |= val[j];
                }
                if ((big >> 23) == 0u) {
                    // 255 loci x 2^23 < 2^31: the item's 32-bit sums cannot wrap
#pragma unroll
                    for (int j = 0; j < kTS; j++) dps[j] += val[j];
                } else {
#pragma unroll
                    for (int j = 0; j < kTS; j++)
                        if (val[j]) atomicAdd((unsigned long long*)&p.dpsum[sb + j], (unsigned long long)val[j]);
                }
                if (neg < 0) {                                // rare: missing depth poisons, negative depth is an error
#pragma unroll
                    for (int j = 0; j < kTS; j++) {
                        if (!((pass >> j) & 1u) || dpv[j] >= 0) continue;
                        if (dpv[j] == INT_MIN) poison |= 1u << j;
                        else atomicMin(p.neg_dp_locus, (int)l);
                    }
                }
            }
            if (WANT_MASK && sb < p.S) {
                uint32_t* cm = p.call_mask + (size_t)l * p.S + sb;
                uint32_t m[kTS];
#pragma unroll
                for (int j = 0; j < kTS; j++) m[j] = fired[j] | (((nocall >> j) & 1u) ? 0x80000000u : 0u);
                if (sb + kTS <= p.S) {
                    reinterpret_cast<uint4*>(cm)[0] = make_uint4(m[0], m[1], m[2], m[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < kTS; j++)
                        if (sb + j < p.S) cm[j] = m[j];
                }
            }
            const uint32_t filt = fired_any & ~nocall;       // filtered calls: every haplotype -> '.', unphased
            if (filt) {
#pragma unroll
                for (int j = 0; j < kTS; j++) {
                    if (!((filt >> j) & 1u)) continue;
#pragma unroll
                    for (int k = 3 * j; k < 3 * j + 3; k++) {
                        const uint32_t val = (k == 3 * j + 2) ? 0u : 0xffffu;
                        w[k >> 1] = (k & 1) ? ((w[k >> 1] & 0x0000ffffu) | (val << 16)) : ((w[k >> 1] & 0xffff0000u) | val);
                    }
                }
                gsrc[0] = make_uint2(w[0], w[1]);
                gsrc[1] = make_uint2(w[2], w[3]);
                gsrc[2] = make_uint2(w[4], w[5]);
            }
            fence_proxy_async_smem();                 // generic-proxy writes -> visible to the store warp's bulk store
            __syncwarp();
            if (lane == 0) mbar_arrive(&done[stage]);
            if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
        // ---- flush the per-sample accumulators of this item -----------------------------------------
#pragma unroll
        for (int j = 0; j < kTS; j++) {
            if (sb + j >= p.S) continue;
            const unsigned nc = (ncalls4 >> (8 * j)) & 0xffu;
            if (nc) atomicAdd((unsigned long long*)&p.numcalls[sb + j], (unsigned long long)nc);
            if (dps[j]) atomicAdd((unsigned long long*)&p.dpsum[sb + j], (unsigned long long)dps[j]);
            if ((poison >> j) & 1u) atomicOr(&p.dp_poison[sb + j], 1u);
            for (int f = 0; f < p.n_specs; f++) {
                const unsigned int c = (fcnt[f * kTmaCons + tid] >> (8 * j)) & 0xffu;
                if (c) atomicAdd((unsigned long long*)&p.filter_counts[(size_t)f * p.S + sb + j], (unsigned long long)c);
            }
        }
    }
}

// generic ploidy variant (P != 2): one thread per call, plain atomics; correctness path only
__global__ void call_filter_generic_kernel(CfParams p, int P) {
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < p.L * p.S;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t l = idx / p.S, s = idx % p.S;
        const int16_t* g = (const int16_t*)((const char*)p.gt + (size_t)l * p.pitch) + s * (P + 1);
        int16_t* go = (int16_t*)((char*)p.gt_out + (size_t)l * p.pitch) + s * (P + 1);
        bool nocall = false;
        for (int h = 0; h < P; h++) nocall |= (g[h] == -1);
        uint32_t fired = 0;
        for (int f = 0; f < p.n_specs; f++) {
            const CfSpec sp = p.specs[f];
            bool hit;
            double val;
            if (sp.kind == TRT_CF_MIN || sp.kind == TRT_CF_MAX) {
                if (sp.is_float) {
                    const float x = ((const float*)p.fmt[sp.field])[l * p.S + s];
                    hit = sp.kind == TRT_CF_MIN ? x < sp.thr_f32 : x > sp.thr_f32;
                    val = x;
                } else {
                    const double x = ((const int32_t*)p.fmt[sp.field])[l * p.S + s];
                    hit = sp.kind == TRT_CF_MIN ? x < sp.thr : x > sp.thr;
                    val = x;
                }
            } else if (sp.kind == TRT_CF_RATIO_GT) {
                const double r = (double)((const int32_t*)p.fmt[sp.field])[l * p.S + s] /
                                 (double)((const int32_t*)p.fmt[TRT_FMT_DP])[l * p.S + s];
                hit = r > sp.thr;
                val = r;
            } else if (sp.kind == TRT_CF_HOST_VALUE) {
                const float x = ((const float*)p.fmt[sp.field])[l * p.S + s];
                hit = !isnan(x);
                val = x;
            } else {
                const float* q = (const float*)p.fmt[TRT_FMT_QEXP] + (l * p.S + s) * 3;
                const float x = sp.kind == TRT_CF_QEXP_HET ? q[1] : (sp.kind == TRT_CF_QEXP_HOM ? q[2] : __fadd_rn(q[1], q[2]));
                hit = !nocall && x < sp.thr_f32;
                val = x;
            }
            if (hit) {
                fired |= 1u << f;
                if (!nocall) atomicAdd((unsigned long long*)&p.filter_counts[(size_t)f * p.S + s], 1ull);
            }
            if (p.trig) p.trig[((size_t)f * p.L + l) * p.S + s] = hit ? val : nan("");
        }
        const bool pass = fired == 0 && !nocall;
        if (pass) {
            atomicAdd((unsigned long long*)&p.numcalls[s], 1ull);
            if (p.dp_field >= 0 && p.dp_is_float) {
                // float32 depth (dumpSTR.py:688-713 on a float array: NaN never compares, no INT_MIN sentinel)
                const float d = ((const float*)p.fmt[p.dp_field])[l * p.S + s];
                if (d < 0.f) atomicMin(p.neg_dp_locus, (int)l);
                else if (d > 0.f) atomicAdd((double*)&p.dpsum[s], (double)d);
            } else if (p.dp_field >= 0) {
                const int d = ((const int32_t*)p.fmt[p.dp_field])[l * p.S + s];
                if (d == INT_MIN) atomicOr(&p.dp_poison[s], 1u);
                else if (d < 0) atomicMin(p.neg_dp_locus, (int)l);
                else if (d > 0) atomicAdd((unsigned long long*)&p.dpsum[s], (unsigned long long)d);
            }
        }
        const bool filtered = fired != 0 && !nocall;
        for (int h = 0; h < P; h++) go[h] = filtered ? (int16_t)-1 : g[h];
        go[P] = filtered ? (int16_t)0 : g[P];
        if (p.call_mask) p.call_mask[idx] = fired | (nocall ? 0x80000000u : 0u);
    }
}

struct LfSpec {
    int kind;
    double thr;
};

__global__ void locus_flags_kernel(int64_t L, int64_t S, int n_specs, const LfSpec* __restrict__ specs,
                                   const long long* __restrict__ lc, const double* __restrict__ het,
                                   const double* __restrict__ hwep, const int32_t* __restrict__ hrun,
                                   const int32_t* __restrict__ period, int has_period_info, uint32_t* __restrict__ flags,
                                   double* __restrict__ het_out, double* __restrict__ hwep_out,
                                   long long* __restrict__ n_called_out) {
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    const long long n_called = lc[l * TRT_LC_N + TRT_LC_NFULL];
    uint32_t fl = 0;
    for (int i = 0; i < n_specs; i++) {
        const LfSpec sp = specs[i];
        bool hit = false;
        switch (sp.kind) {
            case TRT_LF_CALLRATE: hit = ((double)n_called / (double)S) < sp.thr; break;   // filters.py:57-59
            case TRT_LF_HWE: hit = hwep[l] < sp.thr; break;                               // NaN never filters
            case TRT_LF_HETLOW: hit = het[l] < sp.thr; break;
            case TRT_LF_HETHIGH: hit = het[l] > sp.thr; break;
            case TRT_LF_HRUN:                                                              // filters.py:205-214
                hit = has_period_info && (period[l] == 5 || period[l] == 6) && hrun[l] >= period[l];
                break;
        }
        if (hit) fl |= 1u << i;
    }
    if (n_called == 0) fl |= 0x80000000u;   // NO_CALLS_REMAINING dumpSTR.py:957-965
    flags[l] = fl;
    het_out[l] = n_called > 0 ? het[l] : -1.0;      // INFO HET/HWEP dumpSTR.py:1314-1336
    hwep_out[l] = n_called > 0 ? hwep[l] : -1.0;
    n_called_out[l] = n_called;
}

}  // namespace

extern "C" {

int trt_call_filters(trt_ctx* ctx, const trt_call_filter_spec* specs, int n_specs, int dp_field_id,
                     trt_call_filter_out* out) {
    if (!ctx || !ctx->block_open || !ctx->have_gt)
        return trt_set_error(ctx, TRT_ESTATE, "trt_call_filters: needs a block with GT");
    if (!out || n_specs < 0 || n_specs > kMaxSpecs || (n_specs > 0 && !specs))
        return trt_set_error(ctx, TRT_EINVAL, "trt_call_filters: bad arguments (at most %d filters)", kMaxSpecs);
    TRT_CUDA(cudaSetDevice(ctx->device));
    const int64_t L = ctx->L, S = ctx->S;
    CfParams p;
    memset(&p, 0, sizeof(p));
    p.gt = ctx->d_gt;
    p.pitch = ctx->gt_pitch;
    p.L = L;
    p.S = S;
    for (int i = 0; i < TRT_FMT_NFIELDS; i++) {
        p.fmt[i] = ctx->d_fmt[i];
        p.fmt_ncol[i] = ctx->fmt_ncol[i];
    }
    p.n_specs = n_specs;
    for (int f = 0; f < n_specs; f++) {
        const trt_call_filter_spec& s = specs[f];
        if (s.kind < TRT_CF_MIN || s.kind > TRT_CF_HOST_VALUE)
            return trt_set_error(ctx, TRT_EINVAL, "trt_call_filters: unknown filter kind %d", s.kind);
        int field = s.field_id;
        if (s.kind >= TRT_CF_QEXP_HET && s.kind <= TRT_CF_QEXP_TOT) field = TRT_FMT_QEXP;
        if (field < 0 || field >= TRT_FMT_NFIELDS || !ctx->d_fmt[field])
            return trt_set_error(ctx, TRT_ESTATE, "trt_call_filters: FORMAT field %d needed by filter %d is not in the block", field, f);
        if (s.kind == TRT_CF_RATIO_GT && !ctx->d_fmt[TRT_FMT_DP])
            return trt_set_error(ctx, TRT_ESTATE, "trt_call_filters: ratio filter %d needs FORMAT DP", f);
        if (s.kind >= TRT_CF_QEXP_HET && s.kind <= TRT_CF_QEXP_TOT && ctx->fmt_ncol[TRT_FMT_QEXP] != 3)
            return trt_set_error(ctx, TRT_EINVAL, "trt_call_filters: QEXP must have 3 columns");
        p.specs[f].kind = s.kind;
        p.specs[f].field = field;
        p.specs[f].is_float = ctx->fmt_is_float[field];
        if (s.kind == TRT_CF_HOST_VALUE && !ctx->fmt_is_float[field])
            return trt_set_error(ctx, TRT_EINVAL, "trt_call_filters: HOST_VALUE filter %d needs a float32 field", f);
        if (s.kind == TRT_CF_RATIO_GT && (ctx->fmt_is_float[field] || ctx->fmt_is_float[TRT_FMT_DP]))
            return trt_set_error(ctx, TRT_EINVAL, "trt_call_filters: ratio filter %d needs int32 fields", f);
        p.specs[f].thr = s.threshold;
        p.specs[f].thr_f32 = (float)s.threshold;
        {   // comparison variant of the TMA kernel
            CfSpec& c = p.specs[f];
            const double t = s.threshold;
            c.thr_i64 = 0;
            c.thr_i32 = 0;
            c.ratio_slot = -1;
            c.hit_missing = 0;
            c.variant = CFV_NEVER;
            if (s.kind == TRT_CF_HOST_VALUE) c.variant = CFV_HOST;
            else if (s.kind == TRT_CF_RATIO_GT) c.variant = (t >= 0.0 && std::isfinite(t)) ? CFV_RATIO_TABLE : CFV_RATIO_EXACT;
            else if (s.kind == TRT_CF_MIN || s.kind == TRT_CF_MAX) {
                if (c.is_float) c.variant = (s.kind == TRT_CF_MIN) ? CFV_MIN_F32 : CFV_MAX_F32;   // NaN threshold: never true
                else if (!std::isnan(t)) {
                    // int32 value v:  v < t <=> v < ceil(t);  v > t <=> v > floor(t)  (exact in integers)
                    const double lim = 4.0e18;
                    const double e = (s.kind == TRT_CF_MIN) ? ceil(t) : floor(t);
                    c.thr_i64 = (long long)std::max(-lim, std::min(lim, e));
                    if (s.kind == TRT_CF_MIN) {
                        if (c.thr_i64 > (long long)INT_MAX) c.variant = CFV_ALWAYS;
                        else if (c.thr_i64 <= (long long)INT_MIN) c.variant = CFV_NEVER;
                        else { c.variant = CFV_MIN_I32; c.thr_i32 = (int)c.thr_i64; }
                    } else {
                        if (c.thr_i64 >= (long long)INT_MAX) c.variant = CFV_NEVER;
                        else if (c.thr_i64 < (long long)INT_MIN) c.variant = CFV_ALWAYS;
                        else { c.variant = CFV_MAX_I32; c.thr_i32 = (int)c.thr_i64; }
                    }
                }
            }
        }
    }
    p.dp_field = (dp_field_id >= 0 && dp_field_id < TRT_FMT_NFIELDS && ctx->d_fmt[dp_field_id]) ? dp_field_id : -1;
    p.dp_is_float = (p.dp_field >= 0) ? ctx->fmt_is_float[p.dp_field] : 0;
    // outputs / scratch
    TRT_TRY(trt_ensure(ctx, ctx->gt_masked_buf, ctx->gt_pitch * (size_t)L + 16));
    p.gt_out = (int16_t*)ctx->gt_masked_buf.p;
    const size_t n_ctr = (size_t)(n_specs + 2) * S;
    TRT_TRY(trt_ensure(ctx, ctx->samp_counts, n_ctr * 8 + 16));
    TRT_TRY(trt_ensure(ctx, ctx->samp_dp, (size_t)S * 4 + 32));
    TRT_CUDA(cudaMemsetAsync(ctx->samp_counts.p, 0, n_ctr * 8 + 16, ctx->stream));
    TRT_CUDA(cudaMemsetAsync(ctx->samp_dp.p, 0, (size_t)S * 4 + 16, ctx->stream));
    p.filter_counts = (long long*)ctx->samp_counts.p;
    p.numcalls = p.filter_counts + (size_t)n_specs * S;
    p.dpsum = p.numcalls + S;
    p.dp_poison = (unsigned int*)ctx->samp_dp.p;
    p.neg_dp_locus = (int*)((char*)ctx->samp_dp.p + (((size_t)S * 4 + 15) & ~size_t(15)));
    const int big = 0x7fffffff;
    TRT_CUDA(cudaMemcpyAsync(p.neg_dp_locus, &big, 4, cudaMemcpyHostToDevice, ctx->stream));
    if (out->call_mask) {
        TRT_TRY(trt_ensure(ctx, ctx->call_mask, (size_t)L * S * 4 + 16));
        p.call_mask = (uint32_t*)ctx->call_mask.p;
    }
    if (out->trigger_values && n_specs) {
        TRT_TRY(trt_ensure(ctx, ctx->trig, (size_t)n_specs * L * S * 8 + 16));
        p.trig = (double*)ctx->trig.p;
    }
    trt_timer_begin(ctx);
    TRT_CUDA(cudaEventRecord(ctx->ev_s0, ctx->stream));
    if (L > 0 && S > 0) {
        // TMA-pipelined path: diploid, plain [L][S] 4-byte fields, vector-aligned rows, no per-call float64 output
        bool tma_ok = ctx->P == 2 && !p.dp_is_float && !p.trig && (S % 4) == 0 && S >= kSlabSamples && (ctx->gt_pitch % 16) == 0 &&
                      ((uintptr_t)ctx->d_gt % 16) == 0 && !getenv("TRT_CF_LEGACY");
        CfTmaParams q;
        memset(&q, 0, sizeof(q));
        q.dp_slot = q.acc_slot = -1;
        if (tma_ok) {
            auto slot_of = [&](int field) -> int {
                for (int k = 0; k < q.n_fields; k++)
                    if (q.field_of_slot[k] == field) return k;
                if (q.n_fields == kTmaMaxFields || ctx->fmt_ncol[field] != 1 || ((uintptr_t)ctx->d_fmt[field] % 16) != 0) return -1;
                q.field_of_slot[q.n_fields] = field;
                return q.n_fields++;
            };
            for (int f = 0; f < n_specs && tma_ok; f++) {
                if (p.specs[f].kind >= TRT_CF_QEXP_HET && p.specs[f].kind <= TRT_CF_QEXP_TOT) { tma_ok = false; break; }
                q.slot_of_spec[f] = slot_of(p.specs[f].field);
                if (q.slot_of_spec[f] < 0) tma_ok = false;
                if (p.specs[f].kind == TRT_CF_RATIO_GT) {
                    q.dp_slot = slot_of(TRT_FMT_DP);
                    if (q.dp_slot < 0) tma_ok = false;
                }
            }
            if (tma_ok && p.dp_field >= 0) {
                q.acc_slot = slot_of(p.dp_field);
                if (q.acc_slot < 0) tma_ok = false;
            }
        }
        if (tma_ok) {
            q.base = p;
            const size_t stage_bytes = (size_t)kTmaGtBytes + (size_t)q.n_fields * kTmaFieldBytes;
            // cut tables of the ratio filters (see CFV_RATIO_TABLE): largest numerator NOT filtered, per depth 0 .. 255,
            // decided with the float64 division numpy performs
            std::vector<int32_t> cuts;
            for (int f = 0; f < n_specs; f++) {
                CfSpec& c = p.specs[f];
                if (c.variant != CFV_RATIO_TABLE) continue;
                c.ratio_slot = q.n_ratio++;
                c.hit_missing = (1.0 > c.thr) ? 1 : 0;
                for (int den = 0; den < kRatioTable; den++) {
                    long long r = 0;
                    if (den > 0) {
                        const double guess = floor(c.thr * (double)den);
                        r = (long long)std::max(0.0, std::min(guess, 2147483646.0));
                        while (r < 2147483647LL && !((double)(r + 1) / (double)den > c.thr)) r++;
                        while (r > 0 && ((double)r / (double)den > c.thr)) r--;
                    }
                    cuts.push_back((int32_t)r);
                }
            }
            if (q.n_ratio) {
                TRT_TRY(trt_ensure(ctx, ctx->cf_specs, cuts.size() * 4 + 16));
                TRT_CUDA(cudaMemcpyAsync(ctx->cf_specs.p, cuts.data(), cuts.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
                TRT_CUDA(cudaStreamSynchronize(ctx->stream));      // `cuts` is a local
                q.ratio_cut = (const int32_t*)ctx->cf_specs.p;
            }
            q.base = p;
            const size_t fixed = (size_t)std::max(n_specs, 1) * kTmaCons * 4 + (size_t)q.n_ratio * kRatioTable * 4 + 8 +
                                 3 * kTmaMaxStages * 8 + 128;
            int stages = (int)(((size_t)ctx->max_smem_optin - fixed) / stage_bytes);
            stages = std::min(stages, kTmaMaxStages);
            if (stages < 3) tma_ok = false;
            else {
                q.stages = stages;
                q.n_slabs = (S + kSlabSamples - 1) / kSlabSamples;
                // items small enough to balance 148 persistent CTAs, large enough to amortise the counter flush
                int64_t per = std::max<int64_t>(32, std::min<int64_t>(255, (L * q.n_slabs) / ((int64_t)ctx->sm_count * 24) + 1));   // <= 255: 8-bit counters
                q.loci_per_item = (int)per;
                q.n_items = q.n_slabs * ((L + per - 1) / per);
                const size_t smem = (size_t)stages * stage_bytes + fixed;
                const int grid = (int)std::min<int64_t>(q.n_items, ctx->sm_count);
                if (p.call_mask) {
                    TRT_CUDA(cudaFuncSetAttribute(call_filter_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    call_filter_tma_kernel<true><<<grid, kTmaCons + 64, smem, ctx->stream>>>(q);
                } else {
                    TRT_CUDA(cudaFuncSetAttribute(call_filter_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    call_filter_tma_kernel<false><<<grid, kTmaCons + 64, smem, ctx->stream>>>(q);
                }
            }
        }
        if (tma_ok) {
        } else if (ctx->P == 2 && !p.dp_is_float) {
            const int64_t slabs = (S + kSlabSamples - 1) / kSlabSamples;
            // enough CTAs to fill the machine several times over, chunks of at least 64 loci
            int64_t chunks = std::max<int64_t>(1, ((int64_t)ctx->sm_count * 16 + slabs - 1) / slabs);
            int64_t per = std::max<int64_t>(64, (L + chunks - 1) / chunks);
            chunks = (L + per - 1) / per;
            p.loci_per_chunk = (int)per;
            dim3 grid((unsigned)slabs, (unsigned)chunks);
            call_filter_kernel<<<grid, kCfThreads, (size_t)std::max(n_specs, 1) * kSlabSamples * 4, ctx->stream>>>(p);
        } else {
            const int blocks = (int)std::min<int64_t>((L * S + 255) / 256, (int64_t)ctx->sm_count * 16);
            call_filter_generic_kernel<<<blocks, 256, 0, ctx->stream>>>(p, ctx->P);
        }
        TRT_KERNEL_CHECK();
    }
    TRT_CUDA(cudaEventRecord(ctx->ev_s1, ctx->stream));
    trt_timer_end(ctx);
    {
        float ms = 0.f;
        TRT_CUDA(cudaEventElapsedTime(&ms, ctx->ev_s0, ctx->ev_s1));
        ctx->last_scan_ms = ms;
    }
    // ---- results ------------------------------------------------------------------------------------
    std::vector<long long> ctr(n_ctr);
    std::vector<unsigned int> poison((size_t)S);
    int neg = big;
    if (n_ctr) TRT_CUDA(cudaMemcpyAsync(ctr.data(), ctx->samp_counts.p, n_ctr * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (S) TRT_CUDA(cudaMemcpyAsync(poison.data(), ctx->samp_dp.p, (size_t)S * 4, cudaMemcpyDeviceToHost, ctx->stream));
    TRT_CUDA(cudaMemcpyAsync(&neg, p.neg_dp_locus, 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (out->call_mask && L * S)
        TRT_CUDA(cudaMemcpyAsync(out->call_mask, ctx->call_mask.p, (size_t)L * S * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (out->trigger_values && n_specs && L * S)
        TRT_CUDA(cudaMemcpyAsync(out->trigger_values, ctx->trig.p, (size_t)n_specs * L * S * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (out->gt_masked && L * S) {
        const size_t row = (size_t)S * (ctx->P + 1) * 2;
        TRT_CUDA(cudaMemcpy2DAsync(out->gt_masked, row, ctx->gt_masked_buf.p, ctx->gt_pitch, row, (size_t)L,
                                   cudaMemcpyDeviceToHost, ctx->stream));
    }
    TRT_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int f = 0; f < n_specs; f++)
        if (out->filter_counts)
            for (int64_t s = 0; s < S; s++) out->filter_counts[(size_t)f * S + s] += ctr[(size_t)f * S + s];
    for (int64_t s = 0; s < S; s++) {
        if (out->numcalls) out->numcalls[s] += ctr[(size_t)n_specs * S + s];
        if (out->totaldp) {
            if (p.dp_field < 0) out->totaldp[s] = nan("");                       // dumpSTR.py:714-715
            else if (poison[s]) out->totaldp[s] = nan("");                       // dumpSTR.py:710-713
            else if (p.dp_is_float) {
                double d;
                memcpy(&d, &ctr[(size_t)(n_specs + 1) * S + s], 8);
                out->totaldp[s] += d;
            } else out->totaldp[s] += (double)ctr[(size_t)(n_specs + 1) * S + s];
        }
    }
    if (out->negative_dp_locus) *out->negative_dp_locus = (neg == big) ? -1 : neg;
    // later statistics read the masked genotypes (the rebuilt TRRecord of dumpSTR.py:748-774)
    ctx->d_gt_active = (const int16_t*)ctx->gt_masked_buf.p;
    ctx->gt_active_pitch = ctx->gt_pitch;
    ctx->have_packed = false;
    return TRT_OK;
}

int trt_locus_filters(trt_ctx* ctx, const trt_locus_filter_spec* specs, int n_specs, int use_length,
                      trt_locus_filter_out* out) {
    if (!ctx || !ctx->block_open || !ctx->have_gt || !ctx->harmonized)
        return trt_set_error(ctx, TRT_ESTATE, "trt_locus_filters: needs a block with GT and trt_harmonize");
    if (!out || n_specs < 0 || n_specs > 31 || (n_specs > 0 && !specs))
        return trt_set_error(ctx, TRT_EINVAL, "trt_locus_filters: bad arguments");
    TRT_CUDA(cudaSetDevice(ctx->device));
    const int64_t L = ctx->L, S = ctx->S, nA = ctx->nA;
    TRT_TRY(trt_ensure(ctx, ctx->ac, (size_t)nA * 4 + 16));
    TRT_TRY(trt_ensure(ctx, ctx->lc, (size_t)L * TRT_LC_N * 8 + 16));
    TRT_TRY(trt_ensure(ctx, ctx->stat_f64, (size_t)L * 8 * 12 + 16));
    // scratch for the flag kernel: [flags u32 L][het f64 L][hwep f64 L][n_called i64 L][specs]
    const size_t off_het = (((size_t)L * 4 + 15) & ~size_t(15));
    const size_t off_hwep = off_het + (size_t)L * 8, off_nc = off_hwep + (size_t)L * 8, off_specs = off_nc + (size_t)L * 8;
    TRT_TRY(trt_ensure(ctx, ctx->misc, off_specs + 32 * sizeof(LfSpec) + 16));
    std::vector<LfSpec> hs((size_t)std::max(n_specs, 1));
    for (int i = 0; i < n_specs; i++) {
        if (specs[i].kind < TRT_LF_CALLRATE || specs[i].kind > TRT_LF_HRUN)
            return trt_set_error(ctx, TRT_EINVAL, "trt_locus_filters: unknown filter kind %d", specs[i].kind);
        hs[i].kind = specs[i].kind;
        hs[i].thr = specs[i].threshold;
    }
    char* base = (char*)ctx->misc.p;
    if (n_specs)
        TRT_CUDA(cudaMemcpyAsync(base + off_specs, hs.data(), n_specs * sizeof(LfSpec), cudaMemcpyHostToDevice, ctx->stream));
    trt_timer_begin(ctx);
    TRT_TRY(trt_prepare_ranks(ctx));
    TRT_CUDA(cudaMemsetAsync(ctx->ac.p, 0, (size_t)nA * 4 + 16, ctx->stream));
    TRT_CUDA(cudaMemsetAsync(ctx->lc.p, 0, (size_t)L * TRT_LC_N * 8 + 16, ctx->stream));
    if (L > 0) {
        TRT_CUDA(cudaEventRecord(ctx->ev_s0, ctx->stream));
        TRT_TRY(trt_run_scan(ctx, nullptr, 1));
        TRT_CUDA(cudaEventRecord(ctx->ev_s1, ctx->stream));
        TRT_TRY(trt_run_epilogue(ctx, use_length, 0.01, 1));
        const double* f = (const double*)ctx->stat_f64.p;
        const int has_period = (ctx->vcftype == TRT_VCF_HIPSTR || ctx->vcftype == TRT_VCF_LONGTR) ? 1 : 0;
        locus_flags_kernel<<<(unsigned)((L + 127) / 128), 128, 0, ctx->stream>>>(
            L, S, n_specs, (const LfSpec*)(base + off_specs), (const long long*)ctx->lc.p, f + L /*het*/, f + 6 * L /*hwep*/,
            (const int32_t*)ctx->hrun.p, (const int32_t*)ctx->period.p, has_period, (uint32_t*)base, (double*)(base + off_het),
            (double*)(base + off_hwep), (long long*)(base + off_nc));
        TRT_KERNEL_CHECK();
    }
    trt_timer_end(ctx);
    if (L > 0) {
        float ms = 0.f;
        TRT_CUDA(cudaEventElapsedTime(&ms, ctx->ev_s0, ctx->ev_s1));
        ctx->last_scan_ms = ms;
    }
#define D2H(dst, src, bytes) \
    if ((dst) && (bytes)) TRT_CUDA(cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyDeviceToHost, ctx->stream))
    D2H(out->flags, base, (size_t)L * 4);
    D2H(out->het, base + off_het, (size_t)L * 8);
    D2H(out->hwep, base + off_hwep, (size_t)L * 8);
    D2H(out->n_called, base + off_nc, (size_t)L * 8);
    D2H(out->ac, ctx->ac.p, (size_t)nA * 4);
    D2H(out->hrun, ctx->hrun.p, (size_t)L * 4);
#undef D2H
    TRT_CUDA(cudaStreamSynchronize(ctx->stream));
    return TRT_OK;
}

}  // extern "C"
