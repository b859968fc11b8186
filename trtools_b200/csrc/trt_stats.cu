// K2b — per-locus statistics from the O(alleles) count tables the GT scan (trt_scan.cu) produces.
//
//   locus_epilogue_kernel   : thread per (group, locus): allele-frequency statistics and the
//       exact two-sided binomial HWE test in FP64.
//
// Reference semantics reproduced (file:line in the reference tree):
//   TRRecord.GetAlleleCounts    trtools/utils/tr_harmonizer.py:1420-1499  (-1/-2 dropped; partial calls count)
//   TRRecord.GetGenotypeCounts  :1326-1418 (sorted haplotypes; genotypes with a no-call dropped, pads kept)
//   TRRecord.GetCalledSamples   :864-897, GetMaxAllele :1542-1575
//   utils.GetHeterozygosity/GetEntropy/GetMean/GetMode/GetVariance  trtools/utils/utils.py:142-296
//   utils.GetHardyWeinbergBinomialTest :298-338 -> scipy.stats.binomtest (two-sided, exact)
#include <float.h>
#include <stdlib.h>
#include <math.h>

#include <algorithm>

#include "trt_internal.cuh"
#include "trt_scan.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// epilogue: exact binomial machinery (scipy.stats.binomtest two-sided)
// ------------------------------------------------------------------------------------------------
__device__ double stirlerr(double n) {
    // log(n!) - log(sqrt(2 pi n) (n/e)^n) for integer n (Loader's saddle-point algorithm)
    const double sfe[16] = {0.0, 0.0810614667953272582196702, 0.0413406959554092940938221,
                            0.02767792568499833914878929, 0.02079067210376509311152277,
                            0.01664469118982119216319487, 0.01387612882307074799874573,
                            0.01189670994589177009505572, 0.010411265261972096497478567,
                            0.009255462182712732917728637, 0.008330563433362871256469318,
                            0.007573675487951840794972024, 0.006942840107209529865664152,
                            0.006408994188004207068439631, 0.005951370112758847735624416,
                            0.005554733551962801371038690};
    if (n < 16.0) return sfe[(int)n];
    const double S0 = 1.0 / 12.0, S1 = 1.0 / 360.0, S2 = 1.0 / 1260.0, S3 = 1.0 / 1680.0, S4 = 1.0 / 1188.0;
    const double nn = n * n;
    if (n > 500.0) return (S0 - S1 / nn) / n;
    if (n > 80.0) return (S0 - (S1 - S2 / nn) / nn) / n;
    if (n > 35.0) return (S0 - (S1 - (S2 - S3 / nn) / nn) / nn) / n;
    return (S0 - (S1 - (S2 - (S3 - S4 / nn) / nn) / nn) / nn) / n;
}

__device__ double bd0(double x, double np) {
    // deviance part x log(x/np) + np - x, evaluated stably
    if (fabs(x - np) < 0.1 * (x + np)) {
        double v = (x - np) / (x + np);
        double s = (x - np) * v;
        if (fabs(s) < DBL_MIN) return s;
        double ej = 2.0 * x * v;
        v = v * v;
        for (int j = 1; j < 1000; j++) {
            ej *= v;
            const double s1 = s + ej / (double)((j << 1) + 1);
            if (s1 == s) return s1;
            s = s1;
        }
    }
    return x * log(x / np) + np - x;
}

__device__ double binom_pmf(double x, double n, double pr) {
    const double q = 1.0 - pr;
    if (x < 0.0 || x > n) return 0.0;
    if (pr == 0.0) return x == 0.0 ? 1.0 : 0.0;
    if (q == 0.0) return x == n ? 1.0 : 0.0;
    if (x == 0.0) {
        if (n == 0.0) return 1.0;
        const double lc = (pr < 0.1) ? -bd0(n, n * q) - n * pr : n * log(q);
        return exp(lc);
    }
    if (x == n) {
        const double lc = (q < 0.1) ? -bd0(n, n * pr) - n * q : n * log(pr);
        return exp(lc);
    }
    const double lc = stirlerr(n) - stirlerr(x) - stirlerr(n - x) - bd0(x, n * pr) - bd0(n - x, n * q);
    const double lf = 1.837877066409345483560659472811 + log(x) + log1p(-x / n);
    return exp(lc - 0.5 * lf);
}

// Tail sums.  Consecutive terms differ by a rational factor, t' = t * a / b; dividing per term makes every locus a chain
// of ~2 sqrt(n) dependent FP64 divisions (the epilogue was bound by exactly that latency).  Instead kTailBlock terms are
// accumulated over their common denominator — acc / Pd = sum over the block of prod(a) / prod(b) — with ONE reciprocal
// per block.  a, b <= ~1e10 so eight factors stay far inside the FP64 range.
constexpr int kTailBlock = 8;

// P(X <= k): terms summed outward from k (decreasing away from the mode)
__device__ double binom_cdf(double k, double n, double pr) {
    if (k < 0.0) return 0.0;
    if (k >= n) return 1.0;
    const double q = 1.0 - pr;
    if (pr == 0.0) return 1.0;
    if (q == 0.0) return 0.0;
    double t = binom_pmf(k, n, pr), sum = t;
    const double r = q / pr;
    double i = k;
    while (i >= 1.0) {
        double Pn = 1.0, Pd = 1.0, acc = 0.0;
        bool falling = false;
        for (int c = 0; c < kTailBlock && i >= 1.0; c++, i -= 1.0) {
            const double a = i * r, b = n - i + 1.0;      // t_{i-1} = t_i * a / b
            Pn *= a;
            Pd *= b;
            acc = fma(acc, b, Pn);
            falling = a < b;
        }
        const double inv = 1.0 / Pd;
        sum += t * (acc * inv);
        t *= Pn * inv;
        if (falling && t < sum * 1e-18) break;
    }
    return sum;
}

// P(X >= k)
__device__ double binom_upper(double k, double n, double pr) {
    if (k <= 0.0) return 1.0;
    if (k > n) return 0.0;
    const double q = 1.0 - pr;
    if (pr == 0.0) return 0.0;
    if (q == 0.0) return 1.0;
    double t = binom_pmf(k, n, pr), sum = t;
    const double r = pr / q;
    double i = k;
    while (i < n) {
        double Pn = 1.0, Pd = 1.0, acc = 0.0;
        bool falling = false;
        for (int c = 0; c < kTailBlock && i < n; c++, i += 1.0) {
            const double a = (n - i) * r, b = i + 1.0;     // t_{i+1} = t_i * a / b
            Pn *= a;
            Pd *= b;
            acc = fma(acc, b, Pn);
            falling = a < b;
        }
        const double inv = 1.0 / Pd;
        sum += t * (acc * inv);
        t *= Pn * inv;
        if (falling && t < sum * 1e-18) break;
    }
    return sum;
}

__device__ double binomtest_two_sided(double k, double n, double pr) {
    if (!(n >= 1.0) || !(pr >= 0.0 && pr <= 1.0) || k < 0.0 || k > n) return nan("");
    const double d = binom_pmf(k, n, pr);
    const double dr = d * (1.0 + 1e-7);
    double pval;
    if (k < pr * n) {
        double lo = ceil(pr * n), hi = n;
        while (lo < hi) {
            const double mid = lo + floor((hi - lo) / 2.0);
            const double v = -binom_pmf(mid, n, pr);
            if (v < -dr) lo = mid + 1.0;
            else if (v > -dr) hi = mid - 1.0;
            else { lo = mid; hi = mid; }
        }
        const double plo = binom_pmf(lo, n, pr);
        const double ix = ((lo <= n) && (-plo <= -dr)) ? lo : lo - 1.0;
        const double y = n - ix + ((dr == binom_pmf(ix, n, pr)) ? 1.0 : 0.0);
        // cdf(k) + sf(n - y),  sf(x) = P(X > x) = P(X >= x + 1)
        pval = binom_cdf(k, n, pr) + binom_upper(n - y + 1.0, n, pr);
    } else {
        double lo = 0.0, hi = floor(pr * n);
        while (lo < hi) {
            const double mid = lo + floor((hi - lo) / 2.0);
            const double v = binom_pmf(mid, n, pr);
            if (v < dr) lo = mid + 1.0;
            else if (v > dr) hi = mid - 1.0;
            else { lo = mid; hi = mid; }
        }
        const double plo = binom_pmf(lo, n, pr);
        const double ix = ((lo >= 0.0) && (plo <= dr)) ? lo : lo - 1.0;
        // cdf(y - 1) + sf(k - 1) with y = ix + 1
        pval = binom_cdf(ix, n, pr) + binom_upper(k, n, pr);
    }
    return fmin(1.0, pval);
}

struct EpiParams {
    int64_t L;
    int G;
    int64_t nA;
    const int32_t* locus_off;
    const double* allele_len;
    const int32_t* len_class;
    const int32_t* seq_class;
    const int32_t* len_order;
    const int32_t* seq_order;
    const int32_t* ac;       // [G][nA]
    const long long* lc;     // [G][L][8]
    int use_length;
    double nalleles_thresh;
    int P;
    // outputs [G][L]
    double *thresh, *het, *entropy, *mean, *mode, *var, *hwep;
    int32_t* nalleles;
    long long* n_hom;
    long long *n_called, *n_nonstrict, *n_padded;
    unsigned long long* n_bad;   // device-wide count of genotype entries outside [-2, A)
};

__global__ void __launch_bounds__(64, 11) locus_epilogue_kernel(EpiParams p) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.L * p.G) return;
    const int64_t g = idx / p.L, l = idx % p.L;
    const int a0 = p.locus_off[l];
    const int A = p.locus_off[l + 1] - a0;
    const int32_t* ac = p.ac + g * p.nA + a0;
    const long long* lc = p.lc + (g * p.L + l) * TRT_LC_N;
    long long total = 0;
    for (int a = 0; a < A; a++) total += ac[a];
    const double tot = (double)total;
    const double NaN = nan("");

    // ---- length-keyed statistics: thresh, mean, mode, var (always by length) -------------------
    double mean = NaN, mode = NaN, var = NaN, thresh = NaN;
    // statistics under the selected relation
    double het = NaN, ent = NaN, sumsq = NaN;
    int nall = 0;
    if (total > 0) {
        const int32_t* order = p.len_order + a0;
        const int32_t* cls = p.len_class + a0;
        // pass 1: mean, mode, thresh, (het/entropy when use_length)
        double m = 0.0, best_f = -1.0, best_len = NaN, fsum = 0.0, ssq = 0.0;
        int i = 0;
        while (i < A) {
            const int c = cls[order[i]];
            long long cnt = 0;
            int j = i;
            while (j < A && cls[order[j]] == c) { cnt += ac[order[j]]; j++; }
            if (cnt > 0) {
                const double f = (double)cnt / tot;
                const double len = p.allele_len[a0 + c];
                m += len * f;
                if (f > best_f) { best_f = f; best_len = len; }   // ascending length: ties keep the smallest
                thresh = len;                                       // last present class = max allele
                fsum += f;
                ssq += f * f;
                if (p.use_length && f >= p.nalleles_thresh) nall++;
            }
            i = j;
        }
        mean = m;
        mode = best_len;
        // pass 2: variance
        double v = 0.0;
        i = 0;
        while (i < A) {
            const int c = cls[order[i]];
            long long cnt = 0;
            int j = i;
            while (j < A && cls[order[j]] == c) { cnt += ac[order[j]]; j++; }
            if (cnt > 0) {
                const double f = (double)cnt / tot;
                const double dl = p.allele_len[a0 + c] - m;
                v += f * dl * dl;
            }
            i = j;
        }
        var = v;
        const int32_t* rorder = p.use_length ? order : p.seq_order + a0;
        const int32_t* rcls = p.use_length ? cls : p.seq_class + a0;
        if (!p.use_length) {
            fsum = 0.0;
            ssq = 0.0;
            i = 0;
            while (i < A) {
                const int c = rcls[rorder[i]];
                long long cnt = 0;
                int j = i;
                while (j < A && rcls[rorder[j]] == c) { cnt += ac[rorder[j]]; j++; }
                if (cnt > 0) {
                    const double f = (double)cnt / tot;
                    fsum += f;
                    ssq += f * f;
                    if (f >= p.nalleles_thresh) nall++;
                }
                i = j;
            }
        }
        sumsq = ssq;
        het = 1.0 - ssq;
        // entropy: scipy.stats.entropy normalises pk by its sum, sums -p ln p, divides by ln 2
        double e = 0.0;
        i = 0;
        while (i < A) {
            const int c = rcls[rorder[i]];
            long long cnt = 0;
            int j = i;
            while (j < A && rcls[rorder[j]] == c) { cnt += ac[rorder[j]]; j++; }
            if (cnt > 0) {
                const double pk = ((double)cnt / tot) / fsum;
                e += -pk * log(pk);
            }
            i = j;
        }
        ent = e / 0.693147180559945309417232121458;
        if (fabs(1.0 - fsum) > 0.001) {   // ValidateAlleleFreqs utils.py:140 (cannot trigger for count-derived freqs)
            het = ent = mean = mode = var = NaN;
        }
    }
    // ---- HWE ---------------------------------------------------------------------------------
    const long long n_full = lc[TRT_LC_NFULL];
    const long long n_hom = p.use_length ? lc[TRT_LC_HOM_LEN] : lc[TRT_LC_HOM_SEQ];
    double hwep = NaN;
    if (total > 0 && lc[TRT_LC_NPAD] == 0 && p.P >= 2 && !isnan(het))
        hwep = binomtest_two_sided((double)n_hom, (double)n_full, sumsq);
    const int64_t o = g * p.L + l;
    if (p.thresh) p.thresh[o] = thresh;
    if (p.het) p.het[o] = het;
    if (p.entropy) p.entropy[o] = ent;
    if (p.mean) p.mean[o] = mean;
    if (p.mode) p.mode[o] = mode;
    if (p.var) p.var[o] = var;
    if (p.hwep) p.hwep[o] = hwep;
    if (p.nalleles) p.nalleles[o] = nall;
    if (p.n_hom) p.n_hom[o] = n_hom;
    p.n_called[o] = n_full;
    p.n_nonstrict[o] = lc[TRT_LC_NNONSTRICT];
    p.n_padded[o] = lc[TRT_LC_NPAD];
    if (lc[6]) atomicAdd(p.n_bad, (unsigned long long)lc[6]);
}

// dense class ranks (needed for the sorted-genotype semantics when ploidy > 2)
__global__ void class_rank_kernel(const int32_t* __restrict__ locus_off, const int32_t* __restrict__ order,
                                  const int32_t* __restrict__ cls, int64_t L, int32_t* __restrict__ rank_out) {
    int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    const int a0 = locus_off[l], A = locus_off[l + 1] - a0;
    int rank = -1, prev = -1;
    for (int i = 0; i < A; i++) {
        const int a = order[a0 + i];
        const int c = cls[a0 + a];
        if (c != prev) { rank++; prev = c; }
        rank_out[a0 + a] = rank;
    }
}

// dense table of sorted-index genotypes of ONE locus: key digit = allele + 2 (-2 -> 0, -1 -> 1)
__global__ void genotype_table_kernel(const int16_t* __restrict__ row, int64_t S, int P, int A,
                                      const uint8_t* __restrict__ mask, unsigned long long* __restrict__ table,
                                      unsigned long long* __restrict__ n_bad) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < S; s += (int64_t)gridDim.x * blockDim.x) {
        if (mask && !mask[s]) continue;
        int k[8];
        bool bad = false;
        for (int h = 0; h < P; h++) {
            int a = row[s * (P + 1) + h];
            if (a < -2 || a >= A) bad = true;
            k[h] = a + 2;
        }
        if (bad) {
            atomicAdd(n_bad, 1ull);
            continue;
        }
        for (int i = 1; i < P; i++) {   // insertion sort, P <= 8
            int v = k[i], j = i - 1;
            while (j >= 0 && k[j] > v) { k[j + 1] = k[j]; j--; }
            k[j + 1] = v;
        }
        long long idx = 0;
        for (int h = 0; h < P; h++) idx = idx * (A + 2) + k[h];
        atomicAdd(&table[idx], 1ull);
    }
}

}  // namespace

// (re)compute dense class ranks into ctx->stat_i32 = [len_rank[nA] | seq_rank[nA]]
int trt_prepare_ranks(trt_ctx* ctx) {
    const int64_t L = ctx->L, nA = ctx->nA;
    TRT_TRY(trt_ensure(ctx, ctx->stat_i32, (size_t)nA * 8 + 16));
    if (L > 0) {
        class_rank_kernel<<<(unsigned)((L + 127) / 128), 128, 0, ctx->stream>>>(
            (const int32_t*)ctx->locus_off.p, (const int32_t*)ctx->len_order.p, (const int32_t*)ctx->len_class.p, L,
            (int32_t*)ctx->stat_i32.p);
        TRT_KERNEL_CHECK();
        class_rank_kernel<<<(unsigned)((L + 127) / 128), 128, 0, ctx->stream>>>(
            (const int32_t*)ctx->locus_off.p, (const int32_t*)ctx->seq_order.p, (const int32_t*)ctx->seq_class.p, L,
            (int32_t*)ctx->stat_i32.p + nA);
        TRT_KERNEL_CHECK();
    }
    return TRT_OK;
}

// epilogue over ctx->ac / ctx->lc -> ctx->stat_f64 = [thresh|het|entropy|mean|mode|var|hwep|nalleles(i32)|n_hom(i64)|
// n_called|n_called_nonstrict|n_padded (i64)] x G*L, then one u64 bad-entry counter
int trt_run_epilogue(trt_ctx* ctx, int use_length, double nalleles_thresh, int G) {
    const int64_t L = ctx->L, nA = ctx->nA;
    const size_t n_out = (size_t)G * L;
    if (n_out == 0) return TRT_OK;
    double* f = (double*)ctx->stat_f64.p;
    EpiParams ep;
    ep.L = L; ep.G = G; ep.nA = nA;
    ep.locus_off = (const int32_t*)ctx->locus_off.p;
    ep.allele_len = (const double*)ctx->allele_len.p;
    ep.len_class = (const int32_t*)ctx->len_class.p;
    ep.seq_class = (const int32_t*)ctx->seq_class.p;
    ep.len_order = (const int32_t*)ctx->len_order.p;
    ep.seq_order = (const int32_t*)ctx->seq_order.p;
    ep.ac = (const int32_t*)ctx->ac.p;
    ep.lc = (const long long*)ctx->lc.p;
    ep.use_length = use_length;
    ep.nalleles_thresh = nalleles_thresh;
    ep.P = ctx->P;
    ep.thresh = f; ep.het = f + n_out; ep.entropy = f + 2 * n_out; ep.mean = f + 3 * n_out;
    ep.mode = f + 4 * n_out; ep.var = f + 5 * n_out; ep.hwep = f + 6 * n_out;
    ep.nalleles = (int32_t*)(f + 7 * n_out);
    ep.n_hom = (long long*)(f + 8 * n_out);
    ep.n_called = (long long*)(f + 9 * n_out);
    ep.n_nonstrict = (long long*)(f + 10 * n_out);
    ep.n_padded = (long long*)(f + 11 * n_out);
    ep.n_bad = (unsigned long long*)(f + 12 * n_out);
    TRT_CUDA(cudaMemsetAsync(ep.n_bad, 0, 8, ctx->stream));
    locus_epilogue_kernel<<<(unsigned)((n_out + 63) / 64), 64, 0, ctx->stream>>>(ep);
    TRT_KERNEL_CHECK();
    return TRT_OK;
}

extern "C" int trt_locus_stats(trt_ctx* ctx, int use_length, const uint8_t* group_masks, int n_groups,
                               double nalleles_thresh, trt_locus_stats_out* out) {
    if (!ctx || !ctx->block_open || !ctx->have_gt || !ctx->harmonized)
        return trt_set_error(ctx, TRT_ESTATE, "trt_locus_stats: needs a block with GT and trt_harmonize");
    if (!out || n_groups < 1 || (n_groups > 1 && !group_masks))
        return trt_set_error(ctx, TRT_EINVAL, "trt_locus_stats: bad arguments");
    TRT_CUDA(cudaSetDevice(ctx->device));
    const int64_t L = ctx->L, S = ctx->S, nA = ctx->nA;
    const int G = n_groups;
    TRT_TRY(trt_ensure(ctx, ctx->ac, (size_t)G * nA * 4 + 16));
    TRT_TRY(trt_ensure(ctx, ctx->lc, (size_t)G * L * TRT_LC_N * 8 + 16));
    const size_t n_out = (size_t)G * L;
    TRT_TRY(trt_ensure(ctx, ctx->stat_f64, n_out * 8 * 12 + 16));
    if (group_masks) {
        TRT_TRY(trt_ensure(ctx, ctx->group_masks, (size_t)G * S + 16));
        TRT_CUDA(cudaMemcpyAsync(ctx->group_masks.p, group_masks, (size_t)G * S, cudaMemcpyHostToDevice, ctx->stream));
    }
    trt_timer_begin(ctx);
    TRT_TRY(trt_prepare_ranks(ctx));
    TRT_CUDA(cudaMemsetAsync(ctx->ac.p, 0, (size_t)G * nA * 4 + 16, ctx->stream));
    TRT_CUDA(cudaMemsetAsync(ctx->lc.p, 0, (size_t)G * L * TRT_LC_N * 8 + 16, ctx->stream));
    if (L > 0) {
        TRT_CUDA(cudaEventRecord(ctx->ev_s0, ctx->stream));
        // every sample group in as few passes over the GT rows as shared memory allows (up to three groups per pass)
        TRT_TRY(trt_run_scan(ctx, group_masks ? (const uint8_t*)ctx->group_masks.p : nullptr, G));
        TRT_CUDA(cudaEventRecord(ctx->ev_s1, ctx->stream));
        TRT_TRY(trt_dist_flush_after_scan(ctx));     // multi-GPU: the previous step's rows travel from here on
        TRT_TRY(trt_run_epilogue(ctx, use_length, nalleles_thresh, G));
    }
    trt_timer_end_async(ctx);       // the copies below queue up behind the epilogue: one host synchronisation per call
    // ---- results to the caller's host buffers --------------------------------------------------
    const double* f = (const double*)ctx->stat_f64.p;
#define D2H(dst, src, bytes) \
    if ((dst) && (bytes)) TRT_CUDA(cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyDeviceToHost, ctx->stream))
    D2H(out->ac, ctx->ac.p, (size_t)G * nA * 4);
    D2H(out->thresh, f, n_out * 8);
    D2H(out->het, f + n_out, n_out * 8);
    D2H(out->entropy, f + 2 * n_out, n_out * 8);
    D2H(out->mean, f + 3 * n_out, n_out * 8);
    D2H(out->mode, f + 4 * n_out, n_out * 8);
    D2H(out->var, f + 5 * n_out, n_out * 8);
    D2H(out->hwep, f + 6 * n_out, n_out * 8);
    D2H(out->nalleles, f + 7 * n_out, n_out * 4);
    D2H(out->n_hom, f + 8 * n_out, n_out * 8);
    D2H(out->n_called, f + 9 * n_out, n_out * 8);
    D2H(out->n_called_nonstrict, f + 10 * n_out, n_out * 8);
    D2H(out->n_padded, f + 11 * n_out, n_out * 8);
    unsigned long long bad = 0;
    if (n_out) TRT_CUDA(cudaMemcpyAsync(&bad, f + 12 * n_out, 8, cudaMemcpyDeviceToHost, ctx->stream));
#undef D2H
    TRT_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->last_scan_ms = 0.0;
    if (L > 0) {
        float ms = 0.f;
        TRT_CUDA(cudaEventElapsedTime(&ms, ctx->ev_s0, ctx->ev_s1));
        ctx->last_scan_ms = ms;
    }
    if (bad && getenv("TRT_DEBUG_BAD")) {
        // diagnostics: which loci report out-of-range genotype entries, and what the scan counted for them
        std::vector<long long> lcv(n_out * TRT_LC_N);
        cudaMemcpy(lcv.data(), ctx->lc.p, n_out * TRT_LC_N * 8, cudaMemcpyDeviceToHost);
        int shown = 0;
        for (size_t i = 0; i < n_out && shown < 12; i++) {
            if (!lcv[i * TRT_LC_N + 6]) continue;
            const int64_t l = (int64_t)(i % L);
            const int A = ctx->h_locus_off[l + 1] - ctx->h_locus_off[l];
            std::vector<int16_t> rowv((size_t)S * (ctx->P + 1));
            cudaMemcpy(rowv.data(), (const char*)ctx->d_gt_active + (size_t)l * ctx->gt_active_pitch, rowv.size() * 2, cudaMemcpyDeviceToHost);
            long long truly = 0, first = -1;
            for (int64_t s2 = 0; s2 < S; s2++)
                for (int h = 0; h < ctx->P; h++) {
                    const int a = rowv[(size_t)s2 * (ctx->P + 1) + h];
                    if (a < -2 || a >= A) { truly++; if (first < 0) first = s2; }
                }
            fprintf(stderr, "[bad] locus %lld A=%d tier=%d n_full=%lld n_non=%lld n_pad=%lld n_bad=%lld | host recount of the row: %lld out-of-range entries (first at sample %lld)\n",
                    (long long)l, A, scan_tier(A), lcv[i * TRT_LC_N + 0], lcv[i * TRT_LC_N + 1], lcv[i * TRT_LC_N + 2], lcv[i * TRT_LC_N + 6], truly, first);
            shown++;
        }
    }
    if (bad)
        return trt_set_error(ctx, TRT_ERECORD,
                             "%lld genotype entries index an allele the record does not have (or are < -2)", (long long)bad);
    return TRT_OK;
}

// TRRecord.GetGenotypeCounts tr_harmonizer.py:1326-1418 for one locus (API edge): counts of the
// index genotypes with haplotypes sorted, as a dense base-(A+2) table (digit = allele + 2).
extern "C" int trt_genotype_counts(trt_ctx* ctx, int64_t locus, const uint8_t* mask_host, int64_t* table_host,
                                   int64_t table_len) {
    if (!ctx || !ctx->block_open || !ctx->have_gt || !ctx->have_alleles)
        return trt_set_error(ctx, TRT_ESTATE, "trt_genotype_counts: needs a block with GT and alleles");
    if (locus < 0 || locus >= ctx->L) return trt_set_error(ctx, TRT_EINVAL, "trt_genotype_counts: locus out of range");
    const int A = ctx->h_locus_off[locus + 1] - ctx->h_locus_off[locus];
    double need = 1.0;
    for (int h = 0; h < ctx->P; h++) need *= (double)(A + 2);
    if (need > (double)(1 << 26) || (int64_t)need != table_len)
        return trt_set_error(ctx, TRT_EINVAL, "trt_genotype_counts: table_len must be (A+2)^P = %.0f (max 2^26)", need);
    TRT_CUDA(cudaSetDevice(ctx->device));
    TRT_TRY(trt_ensure(ctx, ctx->misc, (size_t)(table_len + 1) * 8 + 16));
    TRT_CUDA(cudaMemsetAsync(ctx->misc.p, 0, (size_t)(table_len + 1) * 8, ctx->stream));
    const uint8_t* d_mask = nullptr;
    if (mask_host) {
        TRT_TRY(trt_ensure(ctx, ctx->group_masks, (size_t)ctx->S + 16));
        TRT_CUDA(cudaMemcpyAsync(ctx->group_masks.p, mask_host, (size_t)ctx->S, cudaMemcpyHostToDevice, ctx->stream));
        d_mask = (const uint8_t*)ctx->group_masks.p;
    }
    if (ctx->S > 0) {
        const int16_t* row = (const int16_t*)((const char*)ctx->d_gt_active + (size_t)locus * ctx->gt_active_pitch);
        const int blocks = (int)std::min<int64_t>((ctx->S + 255) / 256, (int64_t)ctx->sm_count * 4);
        genotype_table_kernel<<<blocks, 256, 0, ctx->stream>>>(row, ctx->S, ctx->P, A, d_mask,
                                                               (unsigned long long*)ctx->misc.p,
                                                               (unsigned long long*)ctx->misc.p + table_len);
        TRT_KERNEL_CHECK();
    }
    std::vector<int64_t> tmp((size_t)table_len + 1);
    TRT_CUDA(cudaMemcpyAsync(tmp.data(), ctx->misc.p, (size_t)(table_len + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    TRT_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(table_host, tmp.data(), (size_t)table_len * 8);
    if (tmp[table_len])
        return trt_set_error(ctx, TRT_ERECORD, "%lld genotypes index an allele the record does not have", (long long)tmp[table_len]);
    return TRT_OK;
}
