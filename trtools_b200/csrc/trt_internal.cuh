// Internal definitions shared by the translation units of libtrtools_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/trtools_b200.h"

struct DevBuf {
    void*  p = nullptr;
    size_t cap = 0;
};

// a gather whose rows are staged but whose exchange has not been issued yet (trt_dist.cu)
struct trt_pending_gather {
    int slot = 0, buf = 0;
    int64_t nbytes = 0;
    std::vector<int64_t> per_rank;
    int dst = 0;
    void* host_out = nullptr;
};

struct trt_ctx {
    int device = -1;
    int sm_count = 0;
    int max_smem_optin = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_s0 = nullptr, ev_s1 = nullptr, ev_u0 = nullptr, ev_u1 = nullptr;
    double last_scan_ms = 0.0;
    std::string err;
    int64_t launches = 0;
    double last_ms = 0.0;
    bool   timer_pending = false;   // trt_timer_end_async: ev1 recorded, elapsed time not read yet

    // ---- block state ----------------------------------------------------------------
    bool    block_open = false;
    int64_t L = 0, S = 0;
    int     P = 0;           // ploidy (columns of GT minus the phase column)
    int     vcftype = -1;
    size_t  gt_pitch = 0;    // bytes per locus row of d_gt (multiple of 16)
    const int16_t* d_gt = nullptr;        // native GT rows (owned: gt_buf, or external)
    const int16_t* d_gt_active = nullptr; // what stats read: d_gt, or the masked copy after call filters
    size_t  gt_active_pitch = 0;
    DevBuf  gt_buf, gt_masked_buf, gt_packed_buf;
    bool    have_gt = false;
    const void* d_fmt[TRT_FMT_NFIELDS] = {nullptr};
    int     fmt_ncol[TRT_FMT_NFIELDS] = {0};
    int     fmt_is_float[TRT_FMT_NFIELDS] = {0};
    DevBuf  fmt_buf[TRT_FMT_NFIELDS];

    // allele table
    bool    have_alleles = false;
    int64_t nA = 0;          // total alleles in the block
    int64_t seq_bytes = 0;
    int     maxA = 0;        // max alleles of a locus in the block
    DevBuf  seqs, allele_off, locus_off, pos, start, end, period, given_len, motif_in;
    bool    have_motif_in = false;
    std::vector<int32_t> h_locus_off;
    std::vector<int32_t> h_period;

    // harmonize outputs
    bool    harmonized = false;
    DevBuf  allele_len, trim_off, trim_len, len_class, seq_class, len_order, seq_order, hrun, hflags, motif, motif_off;
    int64_t motif_bytes = 0;

    // packed tensor
    DevBuf  packed;
    bool    have_packed = false;

    // Beagle allele probabilities (FORMAT AP1 / AP2) and the dosage tensor (trt_dosage.cu)
    DevBuf  ap1, ap2, has_ap, dosage, dosage_err, dos_meta, dos_out;
    DevBuf  reduce_buf;              // scratch of the qcSTR / compareSTR reductions (trt_reduce.cu)
    bool    have_ap = false;

    // stats scratch (device)
    DevBuf  ac, ac_part, lc, group_masks, stat_f64, stat_i32, work_counter;
    DevBuf  scan_lists;              // per-tier locus lists of the current block (trt_scan.cu)
    DevBuf  scan_gbits;              // one byte per sample: membership bits of the sample groups of one scan pass
    bool    scan_lists_valid = false;
    int     scan_tier_off[8] = {0};   // list offsets per tier (+ end)
    int     scan_n_tier[8] = {0}, scan_max_in_tier[8] = {0}, scan_rows_in_tier[8] = {0}, scan_lists_fast = -1;
    bool    want_ac_part = false;
    // dumpSTR scratch
    DevBuf  cf_specs, call_mask, trig, samp_counts, samp_dp, misc;
    // associaTR
    DevBuf  covars, outcome, sample_index, design_row_of_sample, assoc_acc, assoc_out, assoc_tot;
    DevBuf  assoc_zt, assoc_fast_tiles, assoc_tile_fast, assoc_masks, assoc_mom_part;   // fast path (trt_assoc_tile.cu)
    DevBuf  assoc_flags, assoc_mma_tab, assoc_xd, assoc_mma_part, assoc_mma_masks, assoc_colscale;   // tensor path (trt_assoc_mma.cu)
    int64_t n_design = 0;
    int     K = 0;
    bool    have_design = false;
    int64_t design_checked_S = -1;   // sample count the design's sample indices were last validated against

    // NCCL (opaque here)
    void*   nccl_comm = nullptr;
    int     rank = 0, world = 1;
    DevBuf  dist_send, dist_recv;
    DevBuf  dist_recv_r[5];                 // receive buffer per result region (+ one for host-payload gathers)
    DevBuf  dist_stage_r[5][2];             // the rank's own rows of a gather, staged for the side stream (two per slot, alternating)
    int     dist_seq[5] = {0, 0, 0, 0, 0};  // gathers made per slot (parity picks the staging buffer)
    cudaStream_t copy_stream = nullptr;     // device->host copies of gathered tables (overlap the next step's kernels)
    cudaEvent_t  ev_gathered = nullptr, ev_copied[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t  ev_staged[5][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
    cudaEvent_t  ev_sent[5][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
    cudaEvent_t  ev_after_scan = nullptr;
    std::vector<trt_pending_gather> dist_pending;
};

int  trt_set_error(trt_ctx* ctx, int code, const char* fmt, ...);
int  trt_ensure(trt_ctx* ctx, DevBuf& b, size_t bytes);
void trt_free_buf(DevBuf& b);
void trt_timer_begin(trt_ctx* ctx);
void trt_timer_end(trt_ctx* ctx);
void trt_timer_end_async(trt_ctx* ctx);
int  trt_dist_flush_after_scan(trt_ctx* ctx);     // trt_dist.cu: deferred gathers start behind the scan just queued

#define TRT_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            return trt_set_error(ctx, TRT_ECUDA, "%s failed: %s (%s:%d)", #call,                \
                                 cudaGetErrorString(e__), __FILE__, __LINE__);                  \
    } while (0)

#define TRT_TRY(call)                   \
    do {                                \
        int rc__ = (call);              \
        if (rc__ != TRT_OK) return rc__;\
    } while (0)

#define TRT_KERNEL_CHECK()                                                                      \
    do {                                                                                        \
        ctx->launches++;                                                                        \
        cudaError_t e__ = cudaGetLastError();                                                   \
        if (e__ != cudaSuccess)                                                                 \
            return trt_set_error(ctx, TRT_ECUDA, "kernel launch failed: %s (%s:%d)",            \
                                 cudaGetErrorString(e__), __FILE__, __LINE__);                  \
    } while (0)

// per-locus counters produced by the scan kernels (int64 [G][L][TRT_LC_N])
enum {
    TRT_LC_NFULL = 0,      // samples with no -1 haplotype (strictly called)
    TRT_LC_NNONSTRICT = 1, // samples with at least one called haplotype
    TRT_LC_NPAD = 2,       // fully-called samples containing a -2 pad
    TRT_LC_HOM_IDX = 3,    // first two sorted haplotypes equal as allele indices
    TRT_LC_HOM_LEN = 4,    // ... equal as length classes
    TRT_LC_HOM_SEQ = 5,    // ... equal as sequence classes
    TRT_LC_N = 8
};

// ---- device helpers ---------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 1-D bulk async copy global -> shared (TMA engine, no tensor map): size and both addresses
// must be multiples of 16 bytes; completion is signalled on the mbarrier as transaction bytes.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif
