// Context, error handling, device buffers and the blocked ingest stage of libtrtools_b200.so.
#include <stdarg.h>

#include <algorithm>

#include "trt_internal.cuh"

static thread_local std::string g_init_error;

int trt_set_error(trt_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx)
        ctx->err = buf;
    else
        g_init_error = buf;
    return code;
}

int trt_ensure(trt_ctx* ctx, DevBuf& b, size_t bytes) {
    if (bytes == 0) bytes = 16;
    if (b.cap >= bytes) return TRT_OK;
    if (b.p) {
        cudaFree(b.p);
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = (bytes + 255) & ~size_t(255);
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        b.p = nullptr;
        return trt_set_error(ctx, TRT_ENOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    b.cap = want;
    return TRT_OK;
}

void trt_free_buf(DevBuf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

void trt_timer_begin(trt_ctx* ctx) { cudaEventRecord(ctx->ev0, ctx->stream); }
void trt_timer_end(trt_ctx* ctx) {
    cudaEventRecord(ctx->ev1, ctx->stream);
    cudaEventSynchronize(ctx->ev1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_ms = ms;
    ctx->timer_pending = false;
}
// the same without blocking the host: the elapsed time is read when somebody asks for it (trt_last_kernel_ms)
void trt_timer_end_async(trt_ctx* ctx) {
    cudaEventRecord(ctx->ev1, ctx->stream);
    ctx->timer_pending = true;
}
static void timer_resolve(trt_ctx* ctx) {
    if (!ctx->timer_pending) return;
    cudaEventSynchronize(ctx->ev1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_ms = ms;
    ctx->timer_pending = false;
}

// ---- packed GT transfer form <-> native cyvcf2 rows ---------------------------------------------------------------
// packing side: 8 calls per thread, 48 bytes of int16 triples -> 16 packed bytes (+ one phase byte)
__global__ void __launch_bounds__(256) gt_pack_kernel(const int16_t* __restrict__ gt, size_t pitch, int64_t l0, int64_t n, int64_t S,
                                                      uint8_t* __restrict__ packed, uint8_t* __restrict__ phase, int* __restrict__ bad) {
    const int64_t groups = (S + 7) / 8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * groups; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t l = i / groups, gidx = i % groups;
        const int64_t s0 = gidx * 8;
        const int16_t* src = (const int16_t*)((const char*)gt + (size_t)(l0 + l) * pitch) + s0 * 3;
        uint8_t* dst = packed + ((size_t)l * S + s0) * 2;
        unsigned ph = 0;
        for (int j = 0; j < 8 && s0 + j < S; j++) {
            const int a = src[3 * j], b = src[3 * j + 1];
            if (a > 252 || b > 252 || a < -2 || b < -2) *bad = 1;
            dst[2 * j] = (uint8_t)(a >= 0 ? a : 256 + a);
            dst[2 * j + 1] = (uint8_t)(b >= 0 ? b : 256 + b);
            ph |= (src[3 * j + 2] ? 1u : 0u) << j;
        }
        if (phase) phase[(size_t)l * groups + gidx] = (uint8_t)ph;
    }
}

// Both transfer forms -> native rows, one 2048-call tile of one locus per block iteration: a thread decodes 8 calls into
// 48 bytes of shared memory (48-byte stride: conflict-free 16-byte stores), then the block writes the 12 KB tile with
// fully coalesced 16-byte stores (a thread-per-8-calls kernel storing straight to global memory strides its 16-byte
// stores by 48 bytes and reached 2.5-3.4 TB/s; the rows are written once, this is the kernel's whole cost).
template <bool NIBBLE>
__global__ void __launch_bounds__(256) gt_expand_tile_kernel(const uint8_t* __restrict__ packed, const uint8_t* __restrict__ phase,
                                                             int64_t L, int64_t S, int16_t* __restrict__ gt, size_t pitch) {
    __shared__ uint4 tile[768];
    const int tid = threadIdx.x;
    const int64_t tiles_per_row = (S + 2047) / 2048;
    const int64_t groups = (S + 7) / 8;
    for (int64_t t = blockIdx.x; t < L * tiles_per_row; t += gridDim.x) {
        const int64_t l = t / tiles_per_row;
        const int64_t c0 = (t % tiles_per_row) * 2048;
        const int64_t s0 = c0 + (int64_t)tid * 8;
        const int nvalid = (int)max((int64_t)0, min((int64_t)8, S - s0));
        uint32_t o[12];
#pragma unroll
        for (int k = 0; k < 12; k++) o[k] = 0xFEFEFEFEu;                // beyond the row: the pad pattern
        if (nvalid > 0) {
            const unsigned ph = phase ? phase[(size_t)l * groups + (s0 >> 3)] : 0u;
            int16_t h[24];
            if (NIBBLE) {
                const uint8_t* src = packed + (size_t)l * S + s0;
                unsigned long long v = 0;
                if (nvalid == 8 && (((size_t)l * S + s0) & 7) == 0) v = *reinterpret_cast<const unsigned long long*>(src);
                else for (int j = 0; j < nvalid; j++) v |= (unsigned long long)src[j] << (8 * j);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const unsigned b = (unsigned)(v >> (8 * j)) & 0xffu, a0 = b & 15u, a1 = b >> 4;
                    h[3 * j] = (int16_t)(a0 >= 14u ? (int)a0 - 16 : (int)a0);
                    h[3 * j + 1] = (int16_t)(a1 >= 14u ? (int)a1 - 16 : (int)a1);
                    h[3 * j + 2] = (int16_t)((ph >> j) & 1u);
                }
            } else {
                const uint8_t* src = packed + ((size_t)l * S + s0) * 2;
                uint32_t w[4] = {0u, 0u, 0u, 0u};
                if (nvalid == 8 && ((((size_t)l * S + s0) * 2) & 15) == 0) {
                    const uint4 v = *reinterpret_cast<const uint4*>(src);
                    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
                } else {
                    for (int j = 0; j < 2 * nvalid; j++) w[j >> 2] |= (uint32_t)src[j] << (8 * (j & 3));
                }
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const unsigned a = (w[j >> 1] >> (16 * (j & 1))) & 0xffu, b = (w[j >> 1] >> (16 * (j & 1) + 8)) & 0xffu;
                    h[3 * j] = (int16_t)(a >= 254u ? (int)a - 256 : (int)a);
                    h[3 * j + 1] = (int16_t)(b >= 254u ? (int)b - 256 : (int)b);
                    h[3 * j + 2] = (int16_t)((ph >> j) & 1u);
                }
            }
#pragma unroll
            for (int k = 0; k < 12; k++)
                if (2 * k < 3 * nvalid) {
                    const uint32_t lo = (uint16_t)h[2 * k];
                    const uint32_t hi = (2 * k + 1 < 3 * nvalid) ? (uint32_t)(uint16_t)h[2 * k + 1] : 0xFEFEu;
                    o[k] = lo | (hi << 16);
                }
        }
        tile[3 * tid] = make_uint4(o[0], o[1], o[2], o[3]);
        tile[3 * tid + 1] = make_uint4(o[4], o[5], o[6], o[7]);
        tile[3 * tid + 2] = make_uint4(o[8], o[9], o[10], o[11]);
        __syncthreads();
        const int64_t row_vecs = (int64_t)(pitch / 16);                  // the row incl. its pad, in 16-byte units
        const int64_t v0 = c0 * 6 / 16;                                  // c0 is a multiple of 2048: 768 vectors per tile
        uint4* dst = reinterpret_cast<uint4*>((char*)gt + (size_t)l * pitch) + v0;
        const int n_vec = (int)min((int64_t)768, row_vecs - v0);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int i = k * 256 + tid;
            if (i < n_vec) dst[i] = tile[i];
        }
        __syncthreads();
    }
}

// nibble transfer form (alleles <= 13): one byte per call, first haplotype in the low nibble; 14 = ploidy pad, 15 = no call
__global__ void __launch_bounds__(256) gt_pack4_kernel(const int16_t* __restrict__ gt, size_t pitch, int64_t l0, int64_t n, int64_t S,
                                                       uint8_t* __restrict__ packed, uint8_t* __restrict__ phase, int* __restrict__ bad) {
    const int64_t groups = (S + 7) / 8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * groups; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t l = i / groups, gidx = i % groups;
        const int64_t s0 = gidx * 8;
        const int16_t* src = (const int16_t*)((const char*)gt + (size_t)(l0 + l) * pitch) + s0 * 3;
        uint8_t* dst = packed + (size_t)l * S + s0;
        unsigned ph = 0;
        for (int j = 0; j < 8 && s0 + j < S; j++) {
            const int a = src[3 * j], b = src[3 * j + 1];
            if (a > 13 || b > 13 || a < -2 || b < -2) *bad = 1;
            dst[j] = (uint8_t)((a & 15) | ((b & 15) << 4));          // -1 -> 15, -2 -> 14
            ph |= (src[3 * j + 2] ? 1u : 0u) << j;
        }
        if (phase) phase[(size_t)l * groups + gidx] = (uint8_t)ph;
    }
}

template <typename T>
static int upload(trt_ctx* ctx, DevBuf& b, const T* host, size_t n) {
    TRT_TRY(trt_ensure(ctx, b, n * sizeof(T) + 16));
    if (n) TRT_CUDA(cudaMemcpyAsync(b.p, host, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return TRT_OK;
}

extern "C" {

int trt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

const char* trt_last_error(const trt_ctx* ctx) { return ctx ? ctx->err.c_str() : g_init_error.c_str(); }

int trt_init(int device_ordinal, trt_ctx** out) {
    if (!out) return trt_set_error(nullptr, TRT_EINVAL, "trt_init: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return trt_set_error(nullptr, TRT_ENODEV,
                             "no CUDA device available (%s); trtools_b200 has no CPU fallback",
                             e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device_ordinal < 0 || device_ordinal >= n)
        return trt_set_error(nullptr, TRT_EINVAL, "device ordinal %d out of range [0,%d)", device_ordinal, n);
    trt_ctx* ctx = new trt_ctx();
    ctx->device = device_ordinal;
    if ((e = cudaSetDevice(device_ordinal)) != cudaSuccess) {
        int rc = trt_set_error(nullptr, TRT_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
        delete ctx;
        return rc;
    }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device_ordinal);
    ctx->sm_count = prop.multiProcessorCount;
    ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->ev0)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev1)) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->ev_s0)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev_s1)) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->ev_u0)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev_u1)) != cudaSuccess) {
        int rc = trt_set_error(nullptr, TRT_ECUDA, "stream/event creation: %s", cudaGetErrorString(e));
        delete ctx;
        return rc;
    }
    *out = ctx;
    return TRT_OK;
}

void trt_destroy(trt_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    trt_dist_finalize(ctx);
    if (ctx->copy_stream) {
        cudaStreamDestroy(ctx->copy_stream);
        cudaEventDestroy(ctx->ev_gathered);
        for (int i = 0; i < 5; i++) {
            cudaEventDestroy(ctx->ev_copied[i]);
            for (int b = 0; b < 2; b++) {
                cudaEventDestroy(ctx->ev_staged[i][b]);
                cudaEventDestroy(ctx->ev_sent[i][b]);
            }
        }
        cudaEventDestroy(ctx->ev_after_scan);
    }
    for (int i = 0; i < 5; i++) {
        trt_free_buf(ctx->dist_recv_r[i]);
        trt_free_buf(ctx->dist_stage_r[i][0]);
        trt_free_buf(ctx->dist_stage_r[i][1]);
    }
    DevBuf* bufs[] = {&ctx->gt_buf, &ctx->gt_masked_buf, &ctx->gt_packed_buf, &ctx->seqs, &ctx->allele_off, &ctx->locus_off, &ctx->pos,
                      &ctx->start, &ctx->end, &ctx->period, &ctx->given_len, &ctx->motif_in, &ctx->allele_len, &ctx->trim_off,
                      &ctx->trim_len, &ctx->len_class, &ctx->seq_class, &ctx->len_order, &ctx->seq_order, &ctx->hrun,
                      &ctx->hflags, &ctx->motif, &ctx->motif_off, &ctx->packed, &ctx->ac, &ctx->ac_part, &ctx->lc, &ctx->group_masks,
                      &ctx->stat_f64, &ctx->stat_i32, &ctx->work_counter, &ctx->scan_lists, &ctx->scan_gbits, &ctx->ap1, &ctx->ap2, &ctx->has_ap, &ctx->dosage,
                      &ctx->dosage_err, &ctx->dos_meta, &ctx->dos_out, &ctx->reduce_buf, &ctx->cf_specs, &ctx->call_mask, &ctx->trig,
                      &ctx->samp_counts, &ctx->samp_dp, &ctx->misc, &ctx->covars, &ctx->outcome, &ctx->sample_index,
                      &ctx->design_row_of_sample, &ctx->assoc_acc, &ctx->assoc_out, &ctx->assoc_tot, &ctx->assoc_zt, &ctx->assoc_fast_tiles,
                      &ctx->assoc_tile_fast, &ctx->assoc_masks, &ctx->assoc_mom_part, &ctx->assoc_flags, &ctx->assoc_mma_tab, &ctx->assoc_xd,
                      &ctx->assoc_mma_part, &ctx->assoc_mma_masks, &ctx->assoc_colscale, &ctx->dist_send,
                      &ctx->dist_recv};
    for (DevBuf* b : bufs) trt_free_buf(*b);
    for (int i = 0; i < TRT_FMT_NFIELDS; i++) trt_free_buf(ctx->fmt_buf[i]);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int trt_device_info(trt_ctx* ctx, trt_devinfo* out) {
    if (!ctx || !out) return TRT_EINVAL;
    cudaDeviceProp prop;
    TRT_CUDA(cudaGetDeviceProperties(&prop, ctx->device));
    memset(out, 0, sizeof(*out));
    strncpy(out->name, prop.name, sizeof(out->name) - 1);
    out->cc_major = prop.major;
    out->cc_minor = prop.minor;
    out->sm_count = prop.multiProcessorCount;
    out->total_mem_bytes = (int64_t)prop.totalGlobalMem;
    size_t fr = 0, tot = 0;
    cudaMemGetInfo(&fr, &tot);
    out->free_mem_bytes = (int64_t)fr;
    out->l2_bytes = prop.l2CacheSize;
    out->abi_version = TRT_ABI_VERSION;
    return TRT_OK;
}

int trt_synchronize(trt_ctx* ctx) {
    if (!ctx) return TRT_EINVAL;
    TRT_CUDA(cudaStreamSynchronize(ctx->stream));
    return TRT_OK;
}

void* trt_host_alloc(trt_ctx* ctx, size_t bytes) {
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        trt_set_error(ctx, TRT_ENOMEM, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
        return nullptr;
    }
    return p;
}

int trt_host_free(trt_ctx* ctx, void* p) {
    if (p) TRT_CUDA(cudaFreeHost(p));
    return TRT_OK;
}

int64_t trt_launch_count(const trt_ctx* ctx) { return ctx ? ctx->launches : 0; }
double trt_last_kernel_ms(const trt_ctx* ctx) {
    if (!ctx) return 0.0;
    timer_resolve(const_cast<trt_ctx*>(ctx));
    return ctx->last_ms;
}
double trt_last_scan_ms(const trt_ctx* ctx) { return ctx ? ctx->last_scan_ms : 0.0; }
int trt_stopwatch_start(trt_ctx* ctx) {
    if (!ctx) return TRT_EINVAL;
    TRT_CUDA(cudaStreamSynchronize(ctx->stream));
    TRT_CUDA(cudaEventRecord(ctx->ev_u0, ctx->stream));
    return TRT_OK;
}
int trt_stopwatch_stop(trt_ctx* ctx, double* ms_out) {
    if (!ctx || !ms_out) return TRT_EINVAL;
    TRT_CUDA(cudaEventRecord(ctx->ev_u1, ctx->stream));
    TRT_CUDA(cudaEventSynchronize(ctx->ev_u1));
    float ms = 0.f;
    TRT_CUDA(cudaEventElapsedTime(&ms, ctx->ev_u0, ctx->ev_u1));
    *ms_out = ms;
    return TRT_OK;
}

// ---- block ingest ----------------------------------------------------------------------------
int trt_block_begin(trt_ctx* ctx, int64_t n_loci, int64_t n_samples, int ploidy, int vcftype) {
    if (!ctx) return TRT_EINVAL;
    if (n_loci < 0 || n_samples < 0 || ploidy < 1 || ploidy > 8)
        return trt_set_error(ctx, TRT_EINVAL, "trt_block_begin: bad shape L=%lld S=%lld P=%d", (long long)n_loci,
                             (long long)n_samples, ploidy);
    if (vcftype < TRT_VCF_GANGSTR || vcftype > TRT_VCF_LONGTR)
        return trt_set_error(ctx, TRT_EINVAL, "trt_block_begin: unknown vcftype %d", vcftype);
    TRT_CUDA(cudaSetDevice(ctx->device));
    ctx->block_open = true;
    ctx->L = n_loci;
    ctx->S = n_samples;
    ctx->P = ploidy;
    ctx->vcftype = vcftype;
    ctx->have_gt = false;
    ctx->have_alleles = false;
    ctx->scan_lists_valid = false;
    ctx->harmonized = false;
    ctx->have_packed = false;
    ctx->have_ap = false;
    ctx->d_gt = nullptr;
    ctx->d_gt_active = nullptr;
    for (int i = 0; i < TRT_FMT_NFIELDS; i++) {
        ctx->d_fmt[i] = nullptr;
        ctx->fmt_ncol[i] = 0;
    }
    return TRT_OK;
}

static size_t gt_row_bytes(const trt_ctx* ctx) { return (size_t)ctx->S * (ctx->P + 1) * sizeof(int16_t); }

int trt_block_set_gt(trt_ctx* ctx, const int16_t* gt_host) {
    if (!ctx || !ctx->block_open) return trt_set_error(ctx, TRT_ESTATE, "trt_block_set_gt: no open block");
    if (!gt_host && ctx->L * ctx->S > 0) return trt_set_error(ctx, TRT_EINVAL, "trt_block_set_gt: NULL array");
    size_t row = gt_row_bytes(ctx);
    size_t pitch = (row + 15) & ~size_t(15);
    if (pitch == 0) pitch = 16;
    TRT_TRY(trt_ensure(ctx, ctx->gt_buf, pitch * (size_t)ctx->L + 16));
    if (ctx->L > 0 && row > 0) {
        if (pitch != row)  // keep the pad bytes defined
            TRT_CUDA(cudaMemsetAsync(ctx->gt_buf.p, 0xFE, pitch * (size_t)ctx->L, ctx->stream));
        TRT_CUDA(cudaMemcpy2DAsync(ctx->gt_buf.p, pitch, gt_host, row, row, (size_t)ctx->L, cudaMemcpyHostToDevice,
                                   ctx->stream));
    }
    ctx->d_gt = (const int16_t*)ctx->gt_buf.p;
    ctx->gt_pitch = pitch;
    ctx->d_gt_active = ctx->d_gt;
    ctx->gt_active_pitch = pitch;
    ctx->have_gt = true;
    ctx->have_packed = false;
    return TRT_OK;
}

int trt_block_set_gt_packed(trt_ctx* ctx, const uint8_t* gt2_host, const uint8_t* phase_bits_host) {
    if (!ctx || !ctx->block_open) return trt_set_error(ctx, TRT_ESTATE, "trt_block_set_gt_packed: no open block");
    if (ctx->P != 2) return trt_set_error(ctx, TRT_EINVAL, "trt_block_set_gt_packed: the packed form is diploid (block ploidy %d)", ctx->P);
    if (!gt2_host && ctx->L * ctx->S > 0) return trt_set_error(ctx, TRT_EINVAL, "trt_block_set_gt_packed: NULL array");
    const int64_t L = ctx->L, S = ctx->S;
    const size_t row = gt_row_bytes(ctx);
    size_t pitch = (row + 15) & ~size_t(15);
    if (pitch == 0) pitch = 16;
    const size_t pbytes = (size_t)(S + 7) / 8;
    TRT_TRY(trt_ensure(ctx, ctx->gt_buf, pitch * (size_t)L + 16));
    TRT_TRY(trt_ensure(ctx, ctx->gt_packed_buf, (size_t)L * S * 2 + (size_t)L * pbytes + 64));
    uint8_t* d_packed = (uint8_t*)ctx->gt_packed_buf.p;
    uint8_t* d_phase = d_packed + (((size_t)L * S * 2 + 15) & ~size_t(15));
    if (L > 0 && S > 0) {
        if (pitch != row) TRT_CUDA(cudaMemsetAsync(ctx->gt_buf.p, 0xFE, pitch * (size_t)L, ctx->stream));   // keep the pad bytes defined
        TRT_CUDA(cudaMemcpyAsync(d_packed, gt2_host, (size_t)L * S * 2, cudaMemcpyHostToDevice, ctx->stream));
        if (phase_bits_host)
            TRT_CUDA(cudaMemcpyAsync(d_phase, phase_bits_host, (size_t)L * pbytes, cudaMemcpyHostToDevice, ctx->stream));
        const int64_t work = L * ((S + 2047) / 2048);
        const unsigned blocks = (unsigned)std::min<int64_t>(work, (int64_t)ctx->sm_count * 16);
        gt_expand_tile_kernel<false><<<blocks, 256, 0, ctx->stream>>>(d_packed, phase_bits_host ? d_phase : nullptr, L, S,
                                                                      (int16_t*)ctx->gt_buf.p, pitch);
        TRT_KERNEL_CHECK();
    }
    ctx->d_gt = (const int16_t*)ctx->gt_buf.p;
    ctx->gt_pitch = pitch;
    ctx->d_gt_active = ctx->d_gt;
    ctx->gt_active_pitch = pitch;
    ctx->have_gt = true;
    ctx->have_packed = false;
    return TRT_OK;
}

int trt_block_get_gt_packed(trt_ctx* ctx, int64_t locus0, int64_t n, uint8_t* gt2_out_host, uint8_t* phase_bits_out_host) {
    if (!ctx || !ctx->have_gt) return trt_set_error(ctx, TRT_ESTATE, "trt_block_get_gt_packed: no GT in the block");
    if (ctx->P != 2) return trt_set_error(ctx, TRT_EINVAL, "trt_block_get_gt_packed: the packed form is diploid");
    if (locus0 < 0 || n < 0 || locus0 + n > ctx->L || (n && !gt2_out_host)) return trt_set_error(ctx, TRT_EINVAL, "trt_block_get_gt_packed: range");
    const int64_t S = ctx->S;
    const size_t pbytes = (size_t)(S + 7) / 8;
    TRT_TRY(trt_ensure(ctx, ctx->gt_packed_buf, (size_t)n * S * 2 + (size_t)n * pbytes + 64 + 16));
    uint8_t* d_packed = (uint8_t*)ctx->gt_packed_buf.p;
    uint8_t* d_phase = d_packed + (((size_t)n * S * 2 + 15) & ~size_t(15));
    int* d_bad = (int*)(d_phase + (((size_t)n * pbytes + 15) & ~size_t(15)));
    int bad = 0;
    TRT_CUDA(cudaMemsetAsync(d_bad, 0, 4, ctx->stream));
    if (n > 0 && S > 0) {
        const int64_t work = n * ((S + 7) / 8);
        const unsigned blocks = (unsigned)std::min<int64_t>((work + 255) / 256, (int64_t)ctx->sm_count * 32);
        gt_pack_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->d_gt_active, ctx->gt_active_pitch, locus0, n, S, d_packed,
                                                         phase_bits_out_host ? d_phase : nullptr, d_bad);
        TRT_KERNEL_CHECK();
        TRT_CUDA(cudaMemcpyAsync(gt2_out_host, d_packed, (size_t)n * S * 2, cudaMemcpyDeviceToHost, ctx->stream));
        if (phase_bits_out_host)
            TRT_CUDA(cudaMemcpyAsync(phase_bits_out_host, d_phase, (size_t)n * pbytes, cudaMemcpyDeviceToHost, ctx->stream));
    }
    TRT_CUDA(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
    TRT_CUDA(cudaStreamSynchronize(ctx->stream));
    if (bad) return trt_set_error(ctx, TRT_EINVAL, "trt_block_get_gt_packed: an allele index above 252 does not fit the packed form");
    return TRT_OK;
}

int trt_block_set_gt_nibble(trt_ctx* ctx, const uint8_t* g4_host, const uint8_t* phase_bits_host) {
    if (!ctx || !ctx->block_open) return trt_set_error(ctx, TRT_ESTATE, "trt_block_set_gt_nibble: no open block");
    if (ctx->P != 2) return trt_set_error(ctx, TRT_EINVAL, "trt_block_set_gt_nibble: the nibble form is diploid (block ploidy %d)", ctx->P);
    if (!g4_host && ctx->L * ctx->S > 0) return trt_set_error(ctx, TRT_EINVAL, "trt_block_set_gt_nibble: NULL array");
    const int64_t L = ctx->L, S = ctx->S;
    const size_t row = gt_row_bytes(ctx);
    size_t pitch = (row + 15) & ~size_t(15);
    if (pitch == 0) pitch = 16;
    const size_t pbytes = (size_t)(S + 7) / 8;
    TRT_TRY(trt_ensure(ctx, ctx->gt_buf, pitch * (size_t)L + 16));
    TRT_TRY(trt_ensure(ctx, ctx->gt_packed_buf, (size_t)L * S + (size_t)L * pbytes + 64));
    uint8_t* d_packed = (uint8_t*)ctx->gt_packed_buf.p;
    uint8_t* d_phase = d_packed + (((size_t)L * S + 15) & ~size_t(15));
    if (L > 0 && S > 0) {
        if (pitch != row) TRT_CUDA(cudaMemsetAsync(ctx->gt_buf.p, 0xFE, pitch * (size_t)L, ctx->stream));   // keep the pad bytes defined
        TRT_CUDA(cudaMemcpyAsync(d_packed, g4_host, (size_t)L * S, cudaMemcpyHostToDevice, ctx->stream));
        if (phase_bits_host)
            TRT_CUDA(cudaMemcpyAsync(d_phase, phase_bits_host, (size_t)L * pbytes, cudaMemcpyHostToDevice, ctx->stream));
        const int64_t work = L * ((S + 2047) / 2048);
        const unsigned blocks = (unsigned)std::min<int64_t>(work, (int64_t)ctx->sm_count * 16);
        gt_expand_tile_kernel<true><<<blocks, 256, 0, ctx->stream>>>(d_packed, phase_bits_host ? d_phase : nullptr, L, S,
                                                                     (int16_t*)ctx->gt_buf.p, pitch);
        TRT_KERNEL_CHECK();
    }
    ctx->d_gt = (const int16_t*)ctx->gt_buf.p;
    ctx->gt_pitch = pitch;
    ctx->d_gt_active = ctx->d_gt;
    ctx->gt_active_pitch = pitch;
    ctx->have_gt = true;
    ctx->have_packed = false;
    return TRT_OK;
}

int trt_block_get_gt_nibble(trt_ctx* ctx, int64_t locus0, int64_t n, uint8_t* g4_out_host, uint8_t* phase_bits_out_host) {
    if (!ctx || !ctx->have_gt) return trt_set_error(ctx, TRT_ESTATE, "trt_block_get_gt_nibble: no GT in the block");
    if (ctx->P != 2) return trt_set_error(ctx, TRT_EINVAL, "trt_block_get_gt_nibble: the nibble form is diploid");
    if (locus0 < 0 || n < 0 || locus0 + n > ctx->L || (n && !g4_out_host)) return trt_set_error(ctx, TRT_EINVAL, "trt_block_get_gt_nibble: range");
    const int64_t S = ctx->S;
    const size_t pbytes = (size_t)(S + 7) / 8;
    TRT_TRY(trt_ensure(ctx, ctx->gt_packed_buf, (size_t)n * S + (size_t)n * pbytes + 64 + 16));
    uint8_t* d_packed = (uint8_t*)ctx->gt_packed_buf.p;
    uint8_t* d_phase = d_packed + (((size_t)n * S + 15) & ~size_t(15));
    int* d_bad = (int*)(d_phase + (((size_t)n * pbytes + 15) & ~size_t(15)));
    int bad = 0;
    TRT_CUDA(cudaMemsetAsync(d_bad, 0, 4, ctx->stream));
    if (n > 0 && S > 0) {
        const int64_t work = n * ((S + 7) / 8);
        const unsigned blocks = (unsigned)std::min<int64_t>((work + 255) / 256, (int64_t)ctx->sm_count * 32);
        gt_pack4_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->d_gt_active, ctx->gt_active_pitch, locus0, n, S, d_packed,
                                                          phase_bits_out_host ? d_phase : nullptr, d_bad);
        TRT_KERNEL_CHECK();
        TRT_CUDA(cudaMemcpyAsync(g4_out_host, d_packed, (size_t)n * S, cudaMemcpyDeviceToHost, ctx->stream));
        if (phase_bits_out_host)
            TRT_CUDA(cudaMemcpyAsync(phase_bits_out_host, d_phase, (size_t)n * pbytes, cudaMemcpyDeviceToHost, ctx->stream));
    }
    TRT_CUDA(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
    TRT_CUDA(cudaStreamSynchronize(ctx->stream));
    if (bad) return trt_set_error(ctx, TRT_EINVAL, "trt_block_get_gt_nibble: an allele index above 13 does not fit the nibble form");
    return TRT_OK;
}

int trt_block_set_gt_device(trt_ctx* ctx, const int16_t* gt_dev, size_t row_pitch_bytes) {
    if (!ctx || !ctx->block_open) return trt_set_error(ctx, TRT_ESTATE, "trt_block_set_gt_device: no open block");
    if (row_pitch_bytes % 16 != 0 || row_pitch_bytes < gt_row_bytes(ctx) || ((uintptr_t)gt_dev & 15))
        return trt_set_error(ctx, TRT_EINVAL,
                             "trt_block_set_gt_device: pointer and row pitch must be 16-byte aligned and pitch >= S*(P+1)*2");
    ctx->d_gt = gt_dev;
    ctx->gt_pitch = row_pitch_bytes;
    ctx->d_gt_active = gt_dev;
    ctx->gt_active_pitch = row_pitch_bytes;
    ctx->have_gt = true;
    ctx->have_packed = false;
    return TRT_OK;
}

static int set_format_host(trt_ctx* ctx, int field_id, const void* v_host, int ncol, size_t elem, int is_float) {
    if (!ctx || !ctx->block_open) return trt_set_error(ctx, TRT_ESTATE, "trt_block_set_format: no open block");
    if (field_id < 0 || field_id >= TRT_FMT_NFIELDS || ncol < 1)
        return trt_set_error(ctx, TRT_EINVAL, "trt_block_set_format: bad field id %d / ncol %d", field_id, ncol);
    size_t bytes = (size_t)ctx->L * ctx->S * ncol * elem;
    TRT_TRY(trt_ensure(ctx, ctx->fmt_buf[field_id], bytes + 16));
    if (bytes) TRT_CUDA(cudaMemcpyAsync(ctx->fmt_buf[field_id].p, v_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ctx->d_fmt[field_id] = ctx->fmt_buf[field_id].p;
    ctx->fmt_ncol[field_id] = ncol;
    ctx->fmt_is_float[field_id] = is_float;
    return TRT_OK;
}

int trt_block_set_format_i32(trt_ctx* ctx, int field_id, const int32_t* v_host) {
    return set_format_host(ctx, field_id, v_host, 1, sizeof(int32_t), 0);
}
int trt_block_set_format_f32(trt_ctx* ctx, int field_id, const float* v_host, int ncol) {
    return set_format_host(ctx, field_id, v_host, ncol, sizeof(float), 1);
}
int trt_block_set_format_device(trt_ctx* ctx, int field_id, const void* v_dev, int ncol, int is_float) {
    if (!ctx || !ctx->block_open) return trt_set_error(ctx, TRT_ESTATE, "trt_block_set_format_device: no open block");
    if (field_id < 0 || field_id >= TRT_FMT_NFIELDS || ncol < 1 || ((uintptr_t)v_dev & 15))
        return trt_set_error(ctx, TRT_EINVAL, "trt_block_set_format_device: bad field id / ncol / alignment");
    ctx->d_fmt[field_id] = v_dev;
    ctx->fmt_ncol[field_id] = ncol;
    ctx->fmt_is_float[field_id] = is_float;
    return TRT_OK;
}

int trt_block_set_alleles(trt_ctx* ctx, const char* seqs, const int64_t* allele_off, const int32_t* locus_off,
                          const int32_t* pos, const int32_t* start, const int32_t* end, const int32_t* period,
                          const double* given_len, const char* motifs) {
    if (!ctx || !ctx->block_open) return trt_set_error(ctx, TRT_ESTATE, "trt_block_set_alleles: no open block");
    if (!allele_off || !locus_off || !pos || !start || !end || !period)
        return trt_set_error(ctx, TRT_EINVAL, "trt_block_set_alleles: NULL table");
    int64_t L = ctx->L;
    if (locus_off[0] != 0) return trt_set_error(ctx, TRT_EINVAL, "locus_off[0] must be 0");
    int maxA = 0;
    for (int64_t l = 0; l < L; l++) {
        int a = locus_off[l + 1] - locus_off[l];
        if (a < 1) return trt_set_error(ctx, TRT_EINVAL, "locus %lld has no REF allele", (long long)l);
        if (a > 32767) return trt_set_error(ctx, TRT_EINVAL, "locus %lld has %d alleles (int16 GT holds < 32768)", (long long)l, a);
        if (a > maxA) maxA = a;
    }
    int64_t nA = locus_off[L];
    if (allele_off[0] != 0) return trt_set_error(ctx, TRT_EINVAL, "allele_off[0] must be 0");
    for (int64_t a = 0; a < nA; a++)
        if (allele_off[a + 1] < allele_off[a]) return trt_set_error(ctx, TRT_EINVAL, "allele_off not monotone at %lld", (long long)a);
    int64_t nbytes = allele_off[nA];
    if (nbytes > 0 && !seqs) return trt_set_error(ctx, TRT_EINVAL, "trt_block_set_alleles: seqs is NULL");
    ctx->nA = nA;
    ctx->seq_bytes = nbytes;
    ctx->maxA = maxA;
    ctx->h_locus_off.assign(locus_off, locus_off + L + 1);
    ctx->h_period.assign(period, period + L);
    TRT_TRY(upload(ctx, ctx->seqs, seqs, (size_t)nbytes));
    TRT_TRY(upload(ctx, ctx->allele_off, allele_off, (size_t)nA + 1));
    TRT_TRY(upload(ctx, ctx->locus_off, locus_off, (size_t)L + 1));
    TRT_TRY(upload(ctx, ctx->pos, pos, (size_t)L));
    TRT_TRY(upload(ctx, ctx->start, start, (size_t)L));
    TRT_TRY(upload(ctx, ctx->end, end, (size_t)L));
    TRT_TRY(upload(ctx, ctx->period, period, (size_t)L));
    if (given_len) {
        TRT_TRY(upload(ctx, ctx->given_len, given_len, (size_t)nA));
    } else {
        // all-NaN table = "derive every length from the sequence"
        std::vector<double> nanv((size_t)nA, __builtin_nan(""));
        TRT_TRY(upload(ctx, ctx->given_len, nanv.data(), (size_t)nA));
        TRT_CUDA(cudaStreamSynchronize(ctx->stream));  // nanv goes out of scope
    }
    {   // motif offsets = exclusive prefix sum of max(period, 0) (the harmonize kernel's output layout)
        std::vector<int64_t> moff((size_t)L + 1, 0);
        for (int64_t l = 0; l < L; l++) moff[l + 1] = moff[l] + (period[l] > 0 ? period[l] : 0);
        ctx->motif_bytes = moff[L];
        TRT_TRY(upload(ctx, ctx->motif_off, moff.data(), (size_t)L + 1));
        TRT_CUDA(cudaStreamSynchronize(ctx->stream));  // moff goes out of scope
    }
    ctx->have_motif_in = false;
    if (motifs) {
        int64_t mbytes = 0;
        for (int64_t l = 0; l < L; l++) mbytes += period[l] > 0 ? period[l] : 0;
        TRT_TRY(upload(ctx, ctx->motif_in, motifs, (size_t)mbytes));
        ctx->have_motif_in = true;
    }
    // the tables above were copied from caller memory that may be pageable: the async copies
    // have been staged by the runtime, but synchronise so the caller may reuse its buffers.
    TRT_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->have_alleles = true;
    ctx->scan_lists_valid = false;
    ctx->harmonized = false;
    ctx->have_packed = false;
    return TRT_OK;
}

int trt_block_get_gt(trt_ctx* ctx, int64_t locus0, int64_t n, int16_t* out_host) {
    if (!ctx || !ctx->have_gt) return trt_set_error(ctx, TRT_ESTATE, "trt_block_get_gt: no GT in the block");
    if (locus0 < 0 || n < 0 || locus0 + n > ctx->L) return trt_set_error(ctx, TRT_EINVAL, "trt_block_get_gt: range");
    size_t row = gt_row_bytes(ctx);
    if (n && row)
        TRT_CUDA(cudaMemcpy2DAsync(out_host, row, (const char*)ctx->d_gt_active + (size_t)locus0 * ctx->gt_active_pitch,
                                   ctx->gt_active_pitch, row, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    TRT_CUDA(cudaStreamSynchronize(ctx->stream));
    return TRT_OK;
}

int trt_block_get_format(trt_ctx* ctx, int field_id, int64_t locus0, int64_t n, void* out_host) {
    if (!ctx || field_id < 0 || field_id >= TRT_FMT_NFIELDS || !ctx->d_fmt[field_id])
        return trt_set_error(ctx, TRT_ESTATE, "trt_block_get_format: field %d not set", field_id);
    if (locus0 < 0 || n < 0 || locus0 + n > ctx->L) return trt_set_error(ctx, TRT_EINVAL, "trt_block_get_format: range");
    size_t row = (size_t)ctx->S * ctx->fmt_ncol[field_id] * 4;
    if (n && row)
        TRT_CUDA(cudaMemcpyAsync(out_host, (const char*)ctx->d_fmt[field_id] + (size_t)locus0 * row, row * (size_t)n,
                                 cudaMemcpyDeviceToHost, ctx->stream));
    TRT_CUDA(cudaStreamSynchronize(ctx->stream));
    return TRT_OK;
}

}  // extern "C"
