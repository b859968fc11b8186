// K5 — dosages: TRRecord.GetDosages for a whole block, one float32 per call.
//
// Reference semantics reproduced (trtools/utils/tr_harmonizer.py:1098-1208, numpy >= 2 promotion rules):
//   bestguess       sum of the haplotypes' repeat-unit lengths (float64), '.' and ploidy pads count 0, cast to float32
//   beagleap        clip(AP1 . alt_lengths, 0, max alt length) + clip(AP2 . alt_lengths, ..) in float64
//                   + clip(1 - sum(AP1), 0, 1) * ref_length + clip(1 - sum(AP2), 0, 1) * ref_length in float32, cast to float32;
//                   AP rows summing to more than 1.1 or holding a negative entry make the RECORD invalid
//   *_norm          (dosage - 2 min_length) / (max_length - min_length) in float32, all zero when the locus has one
//                   length, invalid record when a value is >= 2.1 or <= -0.1, clipped to [0, 2]; bestguess_norm turns
//                   '.' and pads into NaN
// The output is the float32 [L][S] tensor annotaTR batches into PGEN files (annotaTR.py:673-703); record-level
// validation failures come back as one code per locus and become the reference's ValueError / warning + NaN row at the
// Python API edge.  AP1 / AP2 arrive exactly as cyvcf2 returns them per record — float32 [S][A-1] — stacked over loci.
#include <float.h>
#include <limits.h>
#include <math.h>

#include "trt_internal.cuh"

namespace {

struct DosParams {
    int64_t L, S;
    int P;
    const int16_t* gt;
    size_t pitch;
    const int32_t* locus_off;
    const double* allele_len;
    const float* ap1;
    const float* ap2;
    const uint8_t* has_ap;     // [L]
    int type;
    float* out;                // [L][S]
    int32_t* err;              // [L], pre-set to INT_MAX; the smallest code wins (the reference's check order)
};

__device__ float np_sum_f32(const float* a, int n) {
    // np.sum over the contiguous last axis of a float32 array = numpy's pairwise_sum of the row
    // (numpy/_core/src/umath/loops_utils.h.src): plain left-to-right loop below 8 elements, eight partial sums
    // combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) up to 128, halves split at a multiple of 8 above
    if (n < 8) {
        float res = 0.0f;
        for (int i = 0; i < n; i++) res = __fadd_rn(res, a[i]);
        return res;
    }
    if (n <= 128) {
        float r[8];
        for (int j = 0; j < 8; j++) r[j] = a[j];
        int i = 8;
        for (; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] = __fadd_rn(r[j], a[i + j]);
        float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])), __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
        for (; i < n; i++) res = __fadd_rn(res, a[i]);
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __fadd_rn(np_sum_f32(a, n2), np_sum_f32(a + n2, n - n2));
}

__device__ __forceinline__ void flag(int32_t* err, int code) { atomicMin(err, code); }
// np.clip / np.maximum propagate NaN (a missing AP entry makes the sample's dosage NaN); fminf / fmaxf would drop it
__device__ __forceinline__ float clip_f32(float x, float lo, float hi) { return (x != x) ? x : fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ double clip_f64(double x, double lo, double hi) { return (x != x) ? x : fmin(fmax(x, lo), hi); }

__global__ void __launch_bounds__(256) dosage_kernel(DosParams p) {
    const float NaNf = __int_as_float(0x7fc00000);
    for (int64_t l = blockIdx.y; l < p.L; l += gridDim.y) {
        const int a0 = p.locus_off[l];
        const int A = p.locus_off[l + 1] - a0;
        const int nalt = A - 1;
        const double* len = p.allele_len + a0;
        double min_len = len[0], max_len = len[0], max_alt = -DBL_MAX;
        for (int a = 1; a < A; a++) {
            min_len = fmin(min_len, len[a]);
            max_len = fmax(max_len, len[a]);
            max_alt = fmax(max_alt, len[a]);
        }
        const bool beagle = (p.type == TRT_DOSAGE_BEAGLEAP || p.type == TRT_DOSAGE_BEAGLEAP_NORM);
        const bool norm = (p.type == TRT_DOSAGE_BESTGUESS_NORM || p.type == TRT_DOSAGE_BEAGLEAP_NORM);
        const bool no_ap = beagle && !p.has_ap[l];
        if (no_ap && blockIdx.x == 0 && threadIdx.x == 0) flag(&p.err[l], TRT_DE_NO_AP);
        const int16_t* row = (const int16_t*)((const char*)p.gt + (size_t)l * p.pitch);
        // locus l's AP rows start S * (alts before it) floats into the stacked arrays
        const size_t ap_base = (size_t)p.S * (size_t)(a0 - l);
        for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < p.S; s += (int64_t)gridDim.x * blockDim.x) {
            float unnorm;
            if (!beagle) {
                double acc = 0.0;
                bool any_nan = false;
                for (int h = 0; h < p.P; h++) {
                    const int a = row[s * (p.P + 1) + h];
                    if (a >= 0 && a < A) acc += len[a];
                    else if (norm) any_nan = true;            // '.' / pad -> NaN when normalising, else 0
                }
                unnorm = any_nan ? NaNf : (float)acc;
            } else if (no_ap) {
                unnorm = NaNf;
            } else {
                const float* r1 = p.ap1 + ap_base + (size_t)s * nalt;
                const float* r2 = p.ap2 + ap_base + (size_t)s * nalt;
                const float s1 = np_sum_f32(r1, nalt), s2 = np_sum_f32(r2, nalt);
                if (s1 > 1.1f || s2 > 1.1f) flag(&p.err[l], TRT_DE_AP_SUM);
                double d1 = 0.0, d2 = 0.0;
                bool neg = false;
                for (int i = 0; i < nalt; i++) {
                    neg |= (r1[i] < 0.0f) | (r2[i] < 0.0f);
                    d1 = fma((double)r1[i], len[i + 1], d1);
                    d2 = fma((double)r2[i], len[i + 1], d2);
                }
                if (neg) flag(&p.err[l], TRT_DE_AP_NEGATIVE);
                if (nalt > 0) {
                    d1 = clip_f64(d1, 0.0, max_alt);
                    d2 = clip_f64(d2, 0.0, max_alt);
                }
                const float ref1 = clip_f32(__fsub_rn(1.0f, s1), 0.0f, 1.0f), ref2 = clip_f32(__fsub_rn(1.0f, s2), 0.0f, 1.0f);
                const float reflen = (float)len[0];
                const float rd1 = __fmul_rn(ref1, reflen), rd2 = __fmul_rn(ref2, reflen);
                unnorm = (float)(((d1 + d2) + (double)rd1) + (double)rd2);
            }
            float v = unnorm;
            if (norm) {
                if (min_len == max_len) {
                    v = 0.0f;
                } else {
                    v = __fdiv_rn(__fsub_rn(unnorm, (float)(2.0 * min_len)), (float)(max_len - min_len));
                    if (v >= 2.1f || v <= -0.1f) flag(&p.err[l], TRT_DE_NORM_RANGE);
                    v = (v != v) ? v : fminf(fmaxf(v, 0.0f), 2.0f);
                }
            }
            p.out[l * p.S + s] = v;
        }
    }
}

__global__ void fill_i32_kernel(int32_t* a, int64_t n, int32_t v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}

}  // namespace

extern "C" {

int trt_block_set_ap(trt_ctx* ctx, const float* ap1_host, const float* ap2_host, const uint8_t* has_ap_host) {
    if (!ctx || !ctx->block_open || !ctx->have_alleles)
        return trt_set_error(ctx, TRT_ESTATE, "trt_block_set_ap: needs an open block with its allele table");
    TRT_CUDA(cudaSetDevice(ctx->device));
    const int64_t L = ctx->L, S = ctx->S;
    const size_t n = (size_t)S * (size_t)(ctx->nA - L);          // S x (alternate alleles of the block)
    if (n > 0 && (!ap1_host || !ap2_host)) return trt_set_error(ctx, TRT_EINVAL, "trt_block_set_ap: NULL array");
    TRT_TRY(trt_ensure(ctx, ctx->ap1, n * 4 + 16));
    TRT_TRY(trt_ensure(ctx, ctx->ap2, n * 4 + 16));
    TRT_TRY(trt_ensure(ctx, ctx->has_ap, (size_t)L + 16));
    if (n) {
        TRT_CUDA(cudaMemcpyAsync(ctx->ap1.p, ap1_host, n * 4, cudaMemcpyHostToDevice, ctx->stream));
        TRT_CUDA(cudaMemcpyAsync(ctx->ap2.p, ap2_host, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (has_ap_host) {
        if (L) TRT_CUDA(cudaMemcpyAsync(ctx->has_ap.p, has_ap_host, (size_t)L, cudaMemcpyHostToDevice, ctx->stream));
    } else {
        TRT_CUDA(cudaMemsetAsync(ctx->has_ap.p, 1, (size_t)L + 16, ctx->stream));
    }
    TRT_CUDA(cudaStreamSynchronize(ctx->stream));      // the caller may reuse its (possibly pageable) arrays
    ctx->have_ap = true;
    return TRT_OK;
}

int trt_dosages(trt_ctx* ctx, int dosage_type, float* dosage_out, int32_t* error_out) {
    if (!ctx || !ctx->block_open || !ctx->have_gt || !ctx->harmonized)
        return trt_set_error(ctx, TRT_ESTATE, "trt_dosages: needs a block with GT and trt_harmonize");
    if (dosage_type < TRT_DOSAGE_BESTGUESS || dosage_type > TRT_DOSAGE_BEAGLEAP_NORM)
        return trt_set_error(ctx, TRT_EINVAL, "trt_dosages: unknown dosage type %d", dosage_type);
    const bool beagle = (dosage_type == TRT_DOSAGE_BEAGLEAP || dosage_type == TRT_DOSAGE_BEAGLEAP_NORM);
    if (beagle && !ctx->have_ap) return trt_set_error(ctx, TRT_ESTATE, "trt_dosages: Beagle dosages need trt_block_set_ap");
    TRT_CUDA(cudaSetDevice(ctx->device));
    const int64_t L = ctx->L, S = ctx->S;
    TRT_TRY(trt_ensure(ctx, ctx->dosage, (size_t)L * S * 4 + 16));
    TRT_TRY(trt_ensure(ctx, ctx->dosage_err, (size_t)L * 4 + 16));
    trt_timer_begin(ctx);
    if (L > 0) {
        fill_i32_kernel<<<(unsigned)((L + 255) / 256), 256, 0, ctx->stream>>>((int32_t*)ctx->dosage_err.p, L, INT_MAX);
        TRT_KERNEL_CHECK();
    }
    if (L > 0 && S > 0) {
        DosParams p;
        p.L = L; p.S = S; p.P = ctx->P;
        p.gt = ctx->d_gt_active;
        p.pitch = ctx->gt_active_pitch;
        p.locus_off = (const int32_t*)ctx->locus_off.p;
        p.allele_len = (const double*)ctx->allele_len.p;
        p.ap1 = (const float*)ctx->ap1.p;
        p.ap2 = (const float*)ctx->ap2.p;
        p.has_ap = (const uint8_t*)ctx->has_ap.p;
        p.type = dosage_type;
        p.out = (float*)ctx->dosage.p;
        p.err = (int32_t*)ctx->dosage_err.p;
        dim3 grid((unsigned)std::min<int64_t>((S + 255) / 256, 64), (unsigned)std::min<int64_t>(L, 32768));
        dosage_kernel<<<grid, 256, 0, ctx->stream>>>(p);
        TRT_KERNEL_CHECK();
    }
    trt_timer_end(ctx);
    if (dosage_out && L * S) TRT_CUDA(cudaMemcpyAsync(dosage_out, ctx->dosage.p, (size_t)L * S * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (error_out && L) TRT_CUDA(cudaMemcpyAsync(error_out, ctx->dosage_err.p, (size_t)L * 4, cudaMemcpyDeviceToHost, ctx->stream));
    TRT_CUDA(cudaStreamSynchronize(ctx->stream));
    if (error_out)
        for (int64_t l = 0; l < L; l++)
            if (error_out[l] == INT_MAX) error_out[l] = TRT_DE_OK;
    return TRT_OK;
}

}  // extern "C"
