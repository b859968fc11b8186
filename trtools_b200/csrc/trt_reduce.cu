// K6 — reductions of the qcSTR and compareSTR consumers over the resident block.
//
// trt_qc_reduce   (trtools/qcSTR/qcSTR.py:523-570): per sample, the number of loci at which it has a call — a call is
//     anything but "every haplotype is '.'" (:532-533; a lone '.' of a diploid record, [-1, -2], therefore counts) —
//     per locus the number of such calls among the selected samples, and the quality-field sums the quality plots
//     start from (:536-556): with no-calls reading 0, or skipped (--quality-ignore-no-call).
// trt_compare     (trtools/compareSTR/compareSTR.py:508-643): for two call sets of the same loci and a list of shared
//     samples, per locus: samples called in both, sequence- and length-concordant calls (haplotypes compared as sorted
//     pairs when the calls are unphased), and the five sums of the summed length differences from the reference that
//     the overall R^2 needs; per shared sample: the same three counts accumulated over the loci.
// Both are one pass over GT rows already in HBM; integers are exact, the float sums accumulate in FP64.
#include <limits.h>
#include <math.h>

#include <algorithm>

#include "trt_internal.cuh"

namespace {

struct QcParams {
    const int16_t* gt;
    size_t pitch;
    int64_t L, S;
    int P;
    const uint8_t* mask;        // [S] or null
    const float* q;             // [L][S][ncol] or null
    int ncol;
    int ignore_no_call;
    const int32_t* rec_ploidy;       // [L] GT columns of the record itself (cyvcf2's array width - 1), or null: P
    unsigned long long* sample_calls;   // [S]
    double* sample_quality;             // [S]
    unsigned long long* locus_calls;    // [L]
    double* locus_qsum;                 // [L]
    unsigned long long* locus_qn;       // [L]
    int loci_per_block;
};

// grid (sample slabs of 256, locus chunks): a thread owns one sample and walks the chunk's loci, so the per-sample sums
// live in registers; per-locus sums are reduced per warp and added once per warp and locus
__global__ void __launch_bounds__(256) qc_reduce_kernel(QcParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t l0 = (int64_t)blockIdx.y * p.loci_per_block, l1 = min(p.L, l0 + p.loci_per_block);
    const bool in = s < p.S && (!p.mask || p.mask[s]);
    unsigned long long calls = 0;
    double qsum = 0.0;
    for (int64_t l = l0; l < l1; l++) {
        bool call = false;
        float qv = 0.f;
        bool qok = false;
        if (in) {
            const int16_t* g = (const int16_t*)((const char*)p.gt + (size_t)l * p.pitch) + s * (p.P + 1);
            bool all_missing = true;
            const int width = p.rec_ploidy ? p.rec_ploidy[l] : p.P;      // block columns past the record's own are pads
            for (int h = 0; h < width; h++) all_missing = all_missing && (g[h] == -1);
            call = !all_missing;
            if (p.q) {
                float x = p.q[((size_t)l * p.S + s) * p.ncol];
                if (!call) x = nanf("");                       // quality_scores[~calls] = nan
                if (!p.ignore_no_call) {                       // NaN (no call, or a missing value) reads 0
                    qv = (x != x) ? 0.f : x;
                    qok = true;
                } else {
                    qok = (x == x);
                    qv = qok ? x : 0.f;
                }
            }
        }
        calls += call ? 1ull : 0ull;
        if (qok) qsum += (double)qv;
        const unsigned cm = __ballot_sync(0xffffffffu, call);
        if (p.q) {
            const double wq = warp_sum_d(qok ? (double)qv : 0.0);
            const unsigned qm = __ballot_sync(0xffffffffu, qok);
            if (lane == 0 && qm) {
                atomicAdd(&p.locus_qsum[l], wq);
                atomicAdd(&p.locus_qn[l], (unsigned long long)__popc(qm));
            }
        }
        if (lane == 0 && cm) atomicAdd(&p.locus_calls[l], (unsigned long long)__popc(cm));
    }
    if (in) {
        if (calls) atomicAdd(&p.sample_calls[s], calls);
        if (p.q && qsum != 0.0) atomicAdd(&p.sample_quality[s], qsum);
    }
}

struct CmpParams {
    const int16_t* gt1;
    size_t pitch1;
    const int16_t* gt2;          // [L][S2][P+1] dense
    int64_t L, S1, S2, n;
    int P;
    const int32_t* idx1;
    const int32_t* idx2;
    const int32_t* locus_off1;
    const int32_t* locus_off2;
    const int32_t* seq_class1;   // harmonize: first allele of the locus with the same trimmed sequence
    const double* len1;
    const int32_t* seq_id2;      // set-2 allele -> set-1 sequence class, or a negative id of its own
    const double* len2;
    const double* reflen;        // [L] len(record1.ref_allele) / period  (compareSTR.py:548)
    int ignore_phasing;
    // pass 1
    unsigned long long* numcalls;    // [L]
    unsigned long long* n_unphased;  // [L] both-called samples whose two calls are both unphased
    int* status;                     // [L]
    // pass 2
    unsigned long long* conc_seq;
    unsigned long long* conc_len;
    double* len_sums;                // [L][5]
    unsigned long long* s_numcalls;  // [n]
    unsigned long long* s_conc_seq;
    unsigned long long* s_conc_len;
};

__device__ __forceinline__ bool cmp_called(const int16_t* g, int P) {
    bool ok = true;
    for (int h = 0; h < P; h++) ok = ok && (g[h] != -1);
    return ok;
}

// pass 1: one warp per locus — who is called in both sets, are the ploidies equal, how many of those calls are unphased
__global__ void __launch_bounds__(256) compare_count_kernel(CmpParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t l = warp; l < p.L; l += nwarps) {
        const int16_t* r1 = (const int16_t*)((const char*)p.gt1 + (size_t)l * p.pitch1);
        const int16_t* r2 = p.gt2 + (size_t)l * p.S2 * (p.P + 1);
        unsigned long long nc = 0, nu = 0;
        bool ploidy_differs = false;
        for (int64_t i = lane; i < p.n; i += 32) {
            const int16_t* g1 = r1 + (size_t)p.idx1[i] * (p.P + 1);
            const int16_t* g2 = r2 + (size_t)p.idx2[i] * (p.P + 1);
            if (!(cmp_called(g1, p.P) && cmp_called(g2, p.P))) continue;
            nc++;
            int p1 = 0, p2 = 0;
            for (int h = 0; h < p.P; h++) { p1 += (g1[h] != -2); p2 += (g2[h] != -2); }
            ploidy_differs |= (p1 != p2);
            nu += (g1[p.P] == 0 && g2[p.P] == 0) ? 1ull : 0ull;
        }
        nc = (unsigned long long)warp_sum_ll((long long)nc);
        nu = (unsigned long long)warp_sum_ll((long long)nu);
        const bool pd = __any_sync(0xffffffffu, ploidy_differs);
        if (lane == 0) {
            p.numcalls[l] = nc;
            p.n_unphased[l] = nu;
            int st = 0;
            if (nc > 0 && pd) st = 1;                                                       // compareSTR.py:573-575
            else if (nc > 0 && !p.ignore_phasing && nu != 0 && nu != nc) st = 2;            // :581-586
            p.status[l] = st;
        }
    }
}

// pass 2: concordance of every both-called sample under the locus' phasedness
__global__ void __launch_bounds__(256) compare_conc_kernel(CmpParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t l = warp; l < p.L; l += nwarps) {
        if (p.numcalls[l] == 0 || p.status[l] != 0) continue;
        const bool sorted = p.ignore_phasing || p.n_unphased[l] == p.numcalls[l];
        const int a1 = p.locus_off1[l], A1 = p.locus_off1[l + 1] - a1;
        const int a2 = p.locus_off2[l], A2 = p.locus_off2[l + 1] - a2;
        const double reflen = p.reflen[l];
        const int16_t* r1 = (const int16_t*)((const char*)p.gt1 + (size_t)l * p.pitch1);
        const int16_t* r2 = p.gt2 + (size_t)l * p.S2 * (p.P + 1);
        unsigned long long cs = 0, cl = 0;
        double s1 = 0, s2 = 0, s11 = 0, s12 = 0, s22 = 0;
        for (int64_t i = lane; i < p.n; i += 32) {
            const int16_t* g1 = r1 + (size_t)p.idx1[i] * (p.P + 1);
            const int16_t* g2 = r2 + (size_t)p.idx2[i] * (p.P + 1);
            if (!(cmp_called(g1, p.P) && cmp_called(g2, p.P))) continue;
            // sequence class ids (a ploidy pad is its own symbol, equal in both sets) and lengths (pad = -2) of P <= 2 haplotypes
            int q1[2] = {INT_MIN, INT_MIN}, q2[2] = {INT_MIN, INT_MIN};
            double v1[2] = {0, 0}, v2[2] = {0, 0};
            for (int h = 0; h < p.P; h++) {
                const int x = g1[h], y = g2[h];
                q1[h] = (x >= 0 && x < A1) ? p.seq_class1[a1 + x] : INT_MIN + 1;
                v1[h] = (x >= 0 && x < A1) ? p.len1[a1 + x] : -2.0;
                q2[h] = (y >= 0 && y < A2) ? p.seq_id2[a2 + y] : INT_MIN + 1;
                v2[h] = (y >= 0 && y < A2) ? p.len2[a2 + y] : -2.0;
            }
            bool same_seq, same_len;
            if (p.P == 1) {
                same_seq = q1[0] == q2[0];
                same_len = v1[0] == v2[0];
            } else if (sorted) {
                same_seq = (q1[0] == q2[0] && q1[1] == q2[1]) || (q1[0] == q2[1] && q1[1] == q2[0]);
                same_len = (v1[0] == v2[0] && v1[1] == v2[1]) || (v1[0] == v2[1] && v1[1] == v2[0]);
            } else {
                same_seq = q1[0] == q2[0] && q1[1] == q2[1];
                same_len = v1[0] == v2[0] && v1[1] == v2[1];
            }
            cs += same_seq;
            cl += same_len;
            double d1 = 0, d2 = 0;
            for (int h = 0; h < p.P; h++) { d1 += v1[h] - reflen; d2 += v2[h] - reflen; }
            s1 += d1; s2 += d2; s11 += d1 * d1; s12 += d1 * d2; s22 += d2 * d2;
            atomicAdd(&p.s_numcalls[i], 1ull);
            if (same_seq) atomicAdd(&p.s_conc_seq[i], 1ull);
            if (same_len) atomicAdd(&p.s_conc_len[i], 1ull);
        }
        cs = (unsigned long long)warp_sum_ll((long long)cs);
        cl = (unsigned long long)warp_sum_ll((long long)cl);
        s1 = warp_sum_d(s1); s2 = warp_sum_d(s2); s11 = warp_sum_d(s11); s12 = warp_sum_d(s12); s22 = warp_sum_d(s22);
        if (lane == 0) {
            p.conc_seq[l] = cs;
            p.conc_len[l] = cl;
            double* o = p.len_sums + l * 5;
            o[0] = s1; o[1] = s2; o[2] = s11; o[3] = s12; o[4] = s22;
        }
    }
}

}  // namespace

extern "C" {

int trt_qc_reduce(trt_ctx* ctx, const uint8_t* sample_mask_host, const int32_t* rec_ploidy_host, int quality_field,
                  int ignore_no_call, trt_qc_out* out) {
    if (!ctx || !ctx->block_open || !ctx->have_gt) return trt_set_error(ctx, TRT_ESTATE, "trt_qc_reduce: needs a block with GT");
    if (!out) return trt_set_error(ctx, TRT_EINVAL, "trt_qc_reduce: out is NULL");
    if (quality_field >= TRT_FMT_NFIELDS || (quality_field >= 0 && (!ctx->d_fmt[quality_field] || !ctx->fmt_is_float[quality_field])))
        return trt_set_error(ctx, TRT_ESTATE, "trt_qc_reduce: quality field %d is not a float32 field of the block", quality_field);
    TRT_CUDA(cudaSetDevice(ctx->device));
    const int64_t L = ctx->L, S = ctx->S;
    // scratch: [S] sample calls (u64) | [S] sample quality (f64) | [L] locus calls | [L] quality sums | [L] quality counts
    const size_t bytes = ((size_t)2 * S + (size_t)3 * L) * 8 + 64;
    const size_t mask_bytes = ((size_t)S + 15) & ~(size_t)15;
    TRT_TRY(trt_ensure(ctx, ctx->reduce_buf, bytes + mask_bytes + (size_t)L * 4 + 16));
    char* base = (char*)ctx->reduce_buf.p;
    TRT_CUDA(cudaMemsetAsync(base, 0, bytes, ctx->stream));
    QcParams p;
    p.gt = ctx->d_gt_active;
    p.pitch = ctx->gt_active_pitch;
    p.L = L; p.S = S; p.P = ctx->P;
    p.mask = nullptr;
    if (sample_mask_host) {
        uint8_t* dm = (uint8_t*)(base + bytes);
        if (S) TRT_CUDA(cudaMemcpyAsync(dm, sample_mask_host, (size_t)S, cudaMemcpyHostToDevice, ctx->stream));
        p.mask = dm;
    }
    p.rec_ploidy = nullptr;
    if (rec_ploidy_host && L) {
        for (int64_t l = 0; l < L; l++)
            if (rec_ploidy_host[l] < 1 || rec_ploidy_host[l] > ctx->P)
                return trt_set_error(ctx, TRT_EINVAL, "trt_qc_reduce: rec_ploidy[%lld] = %d outside 1..%d", (long long)l,
                                     rec_ploidy_host[l], ctx->P);
        int32_t* dp = (int32_t*)(base + bytes + mask_bytes);
        TRT_CUDA(cudaMemcpyAsync(dp, rec_ploidy_host, (size_t)L * 4, cudaMemcpyHostToDevice, ctx->stream));
        p.rec_ploidy = dp;
    }
    p.q = quality_field >= 0 ? (const float*)ctx->d_fmt[quality_field] : nullptr;
    p.ncol = quality_field >= 0 ? ctx->fmt_ncol[quality_field] : 1;
    p.ignore_no_call = ignore_no_call;
    p.sample_calls = (unsigned long long*)base;
    p.sample_quality = (double*)(base + (size_t)S * 8);
    p.locus_calls = (unsigned long long*)(base + (size_t)2 * S * 8);
    p.locus_qsum = (double*)(base + ((size_t)2 * S + L) * 8);
    p.locus_qn = (unsigned long long*)(base + ((size_t)2 * S + 2 * L) * 8);
    trt_timer_begin(ctx);
    if (L > 0 && S > 0) {
        const int64_t slabs = (S + 255) / 256;
        int64_t chunks = std::max<int64_t>(1, std::min<int64_t>(L, ((int64_t)ctx->sm_count * 8 + slabs - 1) / slabs));
        p.loci_per_block = (int)((L + chunks - 1) / chunks);
        chunks = (L + p.loci_per_block - 1) / p.loci_per_block;
        qc_reduce_kernel<<<dim3((unsigned)slabs, (unsigned)chunks), 256, 0, ctx->stream>>>(p);
        TRT_KERNEL_CHECK();
    }
    trt_timer_end(ctx);
    std::vector<unsigned long long> sc((size_t)S), lc((size_t)L), qn((size_t)L);
    std::vector<double> sq((size_t)S), lq((size_t)L);
    if (S) {
        TRT_CUDA(cudaMemcpyAsync(sc.data(), p.sample_calls, (size_t)S * 8, cudaMemcpyDeviceToHost, ctx->stream));
        TRT_CUDA(cudaMemcpyAsync(sq.data(), p.sample_quality, (size_t)S * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (L) {
        TRT_CUDA(cudaMemcpyAsync(lc.data(), p.locus_calls, (size_t)L * 8, cudaMemcpyDeviceToHost, ctx->stream));
        TRT_CUDA(cudaMemcpyAsync(lq.data(), p.locus_qsum, (size_t)L * 8, cudaMemcpyDeviceToHost, ctx->stream));
        TRT_CUDA(cudaMemcpyAsync(qn.data(), p.locus_qn, (size_t)L * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    TRT_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int64_t s = 0; s < S; s++) {
        if (out->sample_calls) out->sample_calls[s] += (int64_t)sc[s];
        if (out->sample_quality) out->sample_quality[s] += sq[s];
    }
    for (int64_t l = 0; l < L; l++) {
        if (out->locus_calls) out->locus_calls[l] = (int64_t)lc[l];
        if (out->locus_quality) out->locus_quality[l] = qn[l] ? lq[l] / (double)qn[l] : nan("");   // np.mean of nothing is NaN
    }
    return TRT_OK;
}

int trt_compare(trt_ctx* ctx, const trt_compare_in* in, trt_compare_out* out) {
    if (!ctx || !ctx->block_open || !ctx->have_gt || !ctx->harmonized)
        return trt_set_error(ctx, TRT_ESTATE, "trt_compare: needs a block with GT and trt_harmonize");
    if (!in || !out || !in->gt2 || !in->locus_off2 || !in->seq_id2 || !in->len2 || !in->reflen || (in->n_shared > 0 && (!in->idx1 || !in->idx2)))
        return trt_set_error(ctx, TRT_EINVAL, "trt_compare: NULL argument");
    if (ctx->P > 2) return trt_set_error(ctx, TRT_EINVAL, "trt_compare: at most two haplotypes per call (block ploidy %d)", ctx->P);
    const int64_t L = ctx->L, S1 = ctx->S, S2 = in->S2, n = in->n_shared;
    const int P = ctx->P;
    for (int64_t i = 0; i < n; i++)
        if (in->idx1[i] < 0 || in->idx1[i] >= S1 || in->idx2[i] < 0 || in->idx2[i] >= S2)
            return trt_set_error(ctx, TRT_EINVAL, "trt_compare: shared sample %lld is out of range", (long long)i);
    if (in->locus_off2[0] != 0) return trt_set_error(ctx, TRT_EINVAL, "trt_compare: locus_off2[0] must be 0");
    const int64_t nA2 = in->locus_off2[L];
    TRT_CUDA(cudaSetDevice(ctx->device));
    const size_t gt2_bytes = (size_t)L * S2 * (P + 1) * 2;
    // layout: gt2 | idx1 | idx2 | locus_off2 | seq_id2 | len2 | per-locus u64 x4 | status | len_sums | per-sample u64 x3
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~size_t(15); return o; };
    const size_t o_gt2 = take(gt2_bytes), o_i1 = take((size_t)n * 4), o_i2 = take((size_t)n * 4), o_lo2 = take((size_t)(L + 1) * 4),
                 o_sid = take((size_t)nA2 * 4), o_len2 = take((size_t)nA2 * 8), o_rl = take((size_t)L * 8), o_zero = off;
    const size_t o_nc = take((size_t)L * 8), o_nu = take((size_t)L * 8), o_cs = take((size_t)L * 8), o_cl = take((size_t)L * 8),
                 o_st = take((size_t)L * 4), o_ls = take((size_t)L * 5 * 8), o_sn = take((size_t)n * 8), o_ss = take((size_t)n * 8),
                 o_sl = take((size_t)n * 8);
    TRT_TRY(trt_ensure(ctx, ctx->reduce_buf, off + 64));
    char* b = (char*)ctx->reduce_buf.p;
    if (gt2_bytes) TRT_CUDA(cudaMemcpyAsync(b + o_gt2, in->gt2, gt2_bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (n) {
        TRT_CUDA(cudaMemcpyAsync(b + o_i1, in->idx1, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
        TRT_CUDA(cudaMemcpyAsync(b + o_i2, in->idx2, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    TRT_CUDA(cudaMemcpyAsync(b + o_lo2, in->locus_off2, (size_t)(L + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (nA2) {
        TRT_CUDA(cudaMemcpyAsync(b + o_sid, in->seq_id2, (size_t)nA2 * 4, cudaMemcpyHostToDevice, ctx->stream));
        TRT_CUDA(cudaMemcpyAsync(b + o_len2, in->len2, (size_t)nA2 * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (L) TRT_CUDA(cudaMemcpyAsync(b + o_rl, in->reflen, (size_t)L * 8, cudaMemcpyHostToDevice, ctx->stream));
    TRT_CUDA(cudaMemsetAsync(b + o_zero, 0, off - o_zero, ctx->stream));
    CmpParams p;
    p.gt1 = ctx->d_gt_active; p.pitch1 = ctx->gt_active_pitch;
    p.gt2 = (const int16_t*)(b + o_gt2);
    p.L = L; p.S1 = S1; p.S2 = S2; p.n = n; p.P = P;
    p.idx1 = (const int32_t*)(b + o_i1); p.idx2 = (const int32_t*)(b + o_i2);
    p.locus_off1 = (const int32_t*)ctx->locus_off.p; p.locus_off2 = (const int32_t*)(b + o_lo2);
    p.seq_class1 = (const int32_t*)ctx->seq_class.p; p.len1 = (const double*)ctx->allele_len.p;
    p.seq_id2 = (const int32_t*)(b + o_sid); p.len2 = (const double*)(b + o_len2);
    p.reflen = (const double*)(b + o_rl);
    p.ignore_phasing = in->ignore_phasing;
    p.numcalls = (unsigned long long*)(b + o_nc); p.n_unphased = (unsigned long long*)(b + o_nu);
    p.conc_seq = (unsigned long long*)(b + o_cs); p.conc_len = (unsigned long long*)(b + o_cl);
    p.status = (int*)(b + o_st); p.len_sums = (double*)(b + o_ls);
    p.s_numcalls = (unsigned long long*)(b + o_sn); p.s_conc_seq = (unsigned long long*)(b + o_ss);
    p.s_conc_len = (unsigned long long*)(b + o_sl);
    trt_timer_begin(ctx);
    if (L > 0) {
        const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>((L + 7) / 8, (int64_t)ctx->sm_count * 8));
        compare_count_kernel<<<blocks, 256, 0, ctx->stream>>>(p);
        TRT_KERNEL_CHECK();
        compare_conc_kernel<<<blocks, 256, 0, ctx->stream>>>(p);
        TRT_KERNEL_CHECK();
    }
    trt_timer_end(ctx);
#define D2H(dst, src, bytes) \
    if ((dst) && (bytes)) TRT_CUDA(cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyDeviceToHost, ctx->stream))
    D2H(out->numcalls, b + o_nc, (size_t)L * 8);
    D2H(out->conc_seq, b + o_cs, (size_t)L * 8);
    D2H(out->conc_len, b + o_cl, (size_t)L * 8);
    D2H(out->len_sums, b + o_ls, (size_t)L * 5 * 8);
    D2H(out->status, b + o_st, (size_t)L * 4);
    std::vector<long long> sn((size_t)n), ss((size_t)n), sl((size_t)n);
    if (n) {
        TRT_CUDA(cudaMemcpyAsync(sn.data(), b + o_sn, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
        TRT_CUDA(cudaMemcpyAsync(ss.data(), b + o_ss, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
        TRT_CUDA(cudaMemcpyAsync(sl.data(), b + o_sl, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
#undef D2H
    TRT_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int64_t i = 0; i < n; i++) {
        if (out->sample_numcalls) out->sample_numcalls[i] += sn[i];
        if (out->sample_conc_seq) out->sample_conc_seq[i] += ss[i];
        if (out->sample_conc_len) out->sample_conc_len[i] += sl[i];
    }
    return TRT_OK;
}

}  // extern "C"
