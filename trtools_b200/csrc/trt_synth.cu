// Device twin of trtools_b200/synth.py::fill_calls — writes a synthetic HipSTR block
// (GT int16 [L][S][3], DP/DSTUTTER/DFLANKINDEL int32 [L][S], Q float32 [L][S]) straight into HBM.
// Every value is a pure function of (seed, field, global locus index, sample) built from integer
// operations, so the numpy implementation and this kernel agree bit for bit and blocks far larger
// than host memory / PCIe allow can be benchmarked and spot-checked against the oracle.
#include "trt_internal.cuh"

namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

constexpr uint64_t C_L = 0xD1B54A32D192ED03ull;
constexpr uint64_t C_S = 0x8CB92BA72F3D8DD7ull;
constexpr uint64_t C_F = 0xDB4F0B9175AE2165ull;
enum { F_GT0 = 0, F_GT1, F_MISS, F_DP, F_DP2, F_FLANK, F_STUT, F_Q };

__device__ __forceinline__ uint64_t locus_key(uint64_t seed, int fld, uint64_t locus) {
    uint64_t k = mix64(seed ^ ((uint64_t)fld * C_F));
    return mix64(k + locus * C_L);
}
__device__ __forceinline__ uint64_t call_hash(uint64_t lkey, uint64_t sample) { return mix64(lkey + sample * C_S); }

__global__ void __launch_bounds__(256) synth_fill_kernel(uint64_t seed, int64_t locus_offset, int64_t L, int64_t S,
                                                         const uint32_t* __restrict__ cum_freq, uint32_t miss_thresh,
                                                         uint32_t half_thresh, int16_t* __restrict__ gt, size_t pitch,
                                                         int32_t* __restrict__ dp, int32_t* __restrict__ dst,
                                                         int32_t* __restrict__ dfl, float* __restrict__ q) {
    __shared__ uint32_t cum[16];
    __shared__ uint64_t keys[8];
    for (int64_t l = blockIdx.y; l < L; l += gridDim.y) {
        __syncthreads();
        if (threadIdx.x < 16) cum[threadIdx.x] = cum_freq[l * 16 + threadIdx.x];
        if (threadIdx.x < 8) keys[threadIdx.x] = locus_key(seed, threadIdx.x, (uint64_t)(l + locus_offset));
        __syncthreads();
        int16_t* row = (int16_t*)((char*)gt + (size_t)l * pitch);
        for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < S; s += (int64_t)gridDim.x * blockDim.x) {
            const uint32_t u0 = (uint32_t)(call_hash(keys[F_GT0], s) >> 32);
            const uint32_t u1 = (uint32_t)(call_hash(keys[F_GT1], s) >> 32);
            const uint32_t um = (uint32_t)(call_hash(keys[F_MISS], s) >> 32);
            int a0 = 0, a1 = 0;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                a0 += (u0 > cum[k]);
                a1 += (u1 > cum[k]);
            }
            int ph = 1;
            const bool missing = um < miss_thresh;
            const bool half = !missing && um < miss_thresh + half_thresh;
            if (missing) {
                a0 = -1;
                a1 = -2;
                ph = 0;
            } else if (half) {
                a1 = -1;
            }
            row[s * 3 + 0] = (int16_t)a0;
            row[s * 3 + 1] = (int16_t)a1;
            row[s * 3 + 2] = (int16_t)ph;
            if (dp || dst || dfl || q) {
                const uint64_t h_dp = call_hash(keys[F_DP], s), h_dp2 = call_hash(keys[F_DP2], s);
                long long d = (long long)__popcll(h_dp) + (long long)(h_dp2 % 33ull) - 16;
                if (d < 0) d = 0;
                uint64_t f = call_hash(keys[F_FLANK], s), accf = f;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    f = mix64(f);
                    accf &= f;
                }
                uint64_t g = call_hash(keys[F_STUT], s), accs = g;
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    g = mix64(g);
                    accs &= g;
                }
                int fl = min(__popcll(accf), (int)d), st = min(__popcll(accs), (int)d);
                const uint64_t hq = call_hash(keys[F_Q], s);
                const uint64_t qa = hq >> 40, qb = (hq >> 16) & 0xFFFFFFull;
                const uint64_t kq = 16777215ull - ((qa * qb) >> 28);
                float qv = (float)kq / 16777216.0f;
                int dv = (int)d;
                if (missing) {
                    dv = fl = st = INT_MIN;
                    qv = __int_as_float(0x7fc00000);
                }
                if (dp) dp[l * S + s] = dv;
                if (dfl) dfl[l * S + s] = fl;
                if (dst) dst[l * S + s] = st;
                if (q) q[l * S + s] = qv;
            }
        }
    }
}

}  // namespace

extern "C" int trt_synth_fill(trt_ctx* ctx, uint64_t seed, int64_t locus_offset, int64_t n_loci, int64_t n_samples,
                              const uint32_t* cum_freq_host, uint32_t miss_thresh, uint32_t half_thresh, int with_format) {
    if (!ctx || !ctx->block_open) return trt_set_error(ctx, TRT_ESTATE, "trt_synth_fill: call trt_block_begin first");
    if (n_loci != ctx->L || n_samples != ctx->S || ctx->P != 2)
        return trt_set_error(ctx, TRT_EINVAL, "trt_synth_fill: shape must match the open (diploid) block");
    TRT_CUDA(cudaSetDevice(ctx->device));
    const int64_t L = n_loci, S = n_samples;
    const size_t row = (size_t)S * 6, pitch = std::max<size_t>((row + 15) & ~size_t(15), 16);
    TRT_TRY(trt_ensure(ctx, ctx->gt_buf, pitch * (size_t)L + 16));
    TRT_TRY(trt_ensure(ctx, ctx->misc, (size_t)L * 16 * 4 + 16));
    if (L) TRT_CUDA(cudaMemcpyAsync(ctx->misc.p, cum_freq_host, (size_t)L * 16 * 4, cudaMemcpyHostToDevice, ctx->stream));
    int32_t *dp = nullptr, *dst = nullptr, *dfl = nullptr;
    float* q = nullptr;
    const int ids[4] = {TRT_FMT_DP, TRT_FMT_DSTUTTER, TRT_FMT_DFLANKINDEL, TRT_FMT_Q};
    if (with_format == 1) with_format = 15;   // "all four fields"
    {
        const size_t b = (size_t)L * S * 4 + 16;
        for (int i = 0; i < 4; i++)
            if (with_format & (1 << i)) TRT_TRY(trt_ensure(ctx, ctx->fmt_buf[ids[i]], b));
        if (with_format & 1) dp = (int32_t*)ctx->fmt_buf[TRT_FMT_DP].p;
        if (with_format & 2) dst = (int32_t*)ctx->fmt_buf[TRT_FMT_DSTUTTER].p;
        if (with_format & 4) dfl = (int32_t*)ctx->fmt_buf[TRT_FMT_DFLANKINDEL].p;
        if (with_format & 8) q = (float*)ctx->fmt_buf[TRT_FMT_Q].p;
    }
    if (L > 0 && S > 0) {
        if (pitch != row) TRT_CUDA(cudaMemsetAsync(ctx->gt_buf.p, 0xFE, pitch * (size_t)L, ctx->stream));
        trt_timer_begin(ctx);
        dim3 grid((unsigned)std::min<int64_t>((S + 255) / 256, 32), (unsigned)std::min<int64_t>(L, 16384));
        synth_fill_kernel<<<grid, 256, 0, ctx->stream>>>(seed, locus_offset, L, S, (const uint32_t*)ctx->misc.p, miss_thresh,
                                                         half_thresh, (int16_t*)ctx->gt_buf.p, pitch, dp, dst, dfl, q);
        TRT_KERNEL_CHECK();
        trt_timer_end(ctx);
    }
    TRT_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->d_gt = (const int16_t*)ctx->gt_buf.p;
    ctx->gt_pitch = pitch;
    ctx->d_gt_active = ctx->d_gt;
    ctx->gt_active_pitch = pitch;
    ctx->have_gt = true;
    ctx->have_packed = false;
    {
        for (int i = 0; i < 4; i++) {
            if (!(with_format & (1 << i))) continue;
            ctx->d_fmt[ids[i]] = ctx->fmt_buf[ids[i]].p;
            ctx->fmt_ncol[ids[i]] = 1;
            ctx->fmt_is_float[ids[i]] = (ids[i] == TRT_FMT_Q) ? 1 : 0;
        }
    }
    return TRT_OK;
}
