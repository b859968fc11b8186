// Context, error handling, device buffers and the blocked ingest stage of libtrtools_b200.so.
#include <stdarg.h>

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
        for (int i = 0; i < 5; i++) cudaEventDestroy(ctx->ev_copied[i]);
    }
    for (int i = 0; i < 5; i++) trt_free_buf(ctx->dist_recv_r[i]);
    DevBuf* bufs[] = {&ctx->gt_buf, &ctx->gt_masked_buf, &ctx->seqs, &ctx->allele_off, &ctx->locus_off, &ctx->pos,
                      &ctx->start, &ctx->end, &ctx->period, &ctx->given_len, &ctx->motif_in, &ctx->allele_len, &ctx->trim_off,
                      &ctx->trim_len, &ctx->len_class, &ctx->seq_class, &ctx->len_order, &ctx->seq_order, &ctx->hrun,
                      &ctx->hflags, &ctx->motif, &ctx->motif_off, &ctx->packed, &ctx->ac, &ctx->ac_part, &ctx->lc, &ctx->group_masks,
                      &ctx->stat_f64, &ctx->stat_i32, &ctx->work_counter, &ctx->scan_lists, &ctx->ap1, &ctx->ap2, &ctx->has_ap, &ctx->dosage,
                      &ctx->dosage_err, &ctx->dos_meta, &ctx->dos_out, &ctx->cf_specs, &ctx->call_mask, &ctx->trig,
                      &ctx->samp_counts, &ctx->samp_dp, &ctx->misc, &ctx->covars, &ctx->outcome, &ctx->sample_index,
                      &ctx->design_row_of_sample, &ctx->assoc_acc, &ctx->assoc_out, &ctx->assoc_tot, &ctx->assoc_zt, &ctx->assoc_fast_tiles,
                      &ctx->assoc_tile_fast, &ctx->assoc_masks, &ctx->assoc_mom_part, &ctx->dist_send,
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
double trt_last_kernel_ms(const trt_ctx* ctx) { return ctx ? ctx->last_ms : 0.0; }
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
