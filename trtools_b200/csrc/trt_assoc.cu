// K4 — associaTR: per-locus OLS of the outcome on (summed length genotype, intercept, covariates) over the
// called samples of each locus, all accumulation in FP64.
//
// Reference semantics reproduced (file:line in the reference tree):
//   load_and_filter_genotypes.load_trs  trtools/associaTR/load_and_filter_genotypes.py:157-259 (non-dosage)
//       called = GetCalledSamples() & sample filter; allele frequencies over those samples by length rounded to
//       2 dp; filters 'No called samples' / 'Only one called allele' / 'non-major allele count<cutoff'
//   associaTR.perform_gwas_helper       trtools/associaTR/associaTR.py:246-291
//       'n covars >= n samples'; g = sum of the haplotype lengths (a -2 ploidy pad counts as -2, like the
//       reference's float array); g standardised; statsmodels OLS(outcome[called], covars[called]) ->
//       pvalues[0], params[0]/std, bse[0]/std, rsquared
//
// Because the intercept is in the design, standardising g only rescales its coefficient, so the kernel works
// with the raw g (shifted by a per-locus constant for conditioning) and uses the Frisch-Waugh form:
//   M = C'C, v = C'y, u = C'g over the called rows;  a = M^-1 u, b = M^-1 v;
//   gg~ = g'g - u'a, gy~ = g'y - u'b, yy~ = y'y - v'b;  beta = gy~/gg~;  SSR = yy~ - beta*gy~;
//   se = sqrt(SSR/(n-K)/gg~);  p = 2 T_{n-K}.sf(|beta/se|);  R^2 = 1 - SSR/sum((y-ybar)^2).
// M, v, y'y over the called rows = totals over all design rows (once per design) minus the outer products of
// the (few) uncalled rows of that locus (exact per-locus missingness, SURVEY.md §7 "hard parts").
//
// Kernels
//   design_totals_kernel  : C'C, C'y, y'y over all design rows (thread per matrix entry; once per design)
//   assoc_moments_kernel  : tiles of 16 loci x 256-sample chunks; the covariate chunk is staged once in shared
//                           memory and reused by every locus of the tile (each thread: 1 sample x 2 loci in
//                           registers), so L2->SM covariate traffic is ~1x the GT traffic instead of 16x
//   assoc_downdate_kernel : warp per locus; lanes own entries of the outer product of the uncalled rows
//   assoc_solve_kernel    : thread per locus; Cholesky of M, t-test p-value (regularised incomplete beta), filters
#include <math.h>

#include <algorithm>

#include "trt_internal.cuh"
#include "trt_scan.cuh"
#include "trt_assoc_tile.cuh"

namespace {

constexpr int kMaxK = 32;                 // design columns incl. genotype and intercept
constexpr int kTileLoci = 16;
constexpr int kMomThreads = 256;
constexpr int kMomWarps = kMomThreads / 32;
constexpr int kLociPerWarp = kTileLoci / kMomWarps;   // 2
static_assert(kLociPerWarp == 2, "the moments kernel keeps two loci per thread in registers");

// z-vector of a design row: z[0..K-2] = covars columns 1..K-1 (intercept first), z[K-1] = outcome
__host__ __device__ inline int tri_entries(int K) { return K * (K + 1) / 2; }

// one CTA per triangle entry; strided partial sums + a fixed-order tree (bit-reproducible)
__global__ void __launch_bounds__(256) design_totals_kernel(const double* __restrict__ covars, const double* __restrict__ outcome,
                                                            int64_t n, int K, double* __restrict__ tot) {
    __shared__ double part[256];
    const int e = blockIdx.x;
    // entry e -> (a, b), a <= b, row-major upper triangle
    int a = 0, rem = e;
    while (rem >= K - a) { rem -= K - a; a++; }
    const int b = a + rem;
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 256) {
        const double za = (a == K - 1) ? outcome[i] : covars[i * K + a + 1];
        const double zb = (b == K - 1) ? outcome[i] : covars[i * K + b + 1];
        s += za * zb;
    }
    part[threadIdx.x] = s;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) part[threadIdx.x] += part[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) tot[e] = part[0];
}

struct AssocParams {
    const int16_t* gt;
    size_t pitch;
    int64_t L, S;
    int P;
    const int32_t* locus_off;
    const double* allele_len;
    const int32_t* row_of_sample;   // [S] design row or -1
    const double* covars;           // [n][K]
    const double* outcome;          // [n]
    int K;
    double* mom;                    // [L][K+3]: n, sum g', sum g'^2, g'.y, g'.c_1..c_{K-1}
    double* dd;                     // [L][K(K+1)/2] outer products of the uncalled design rows
    const int32_t* list;            // loci handled by the generic kernels (null: all L loci)
    int64_t n_list;
    // a short list is spread over the SMs by cutting the sample axis into segments (grid.y): partial sums go to
    // part [nseg][n_list][K+3 | K(K+1)/2] and are added up in segment order by assoc_generic_reduce_kernel
    int nseg;
    int64_t seg_len;                // multiple of 256
    double* mom_part;
    double* dd_part;
};

// summed length genotype of one call; returns false if the sample is not (strictly) called
__device__ __forceinline__ bool call_value(const int16_t* g, int P, int A, const double* len, double& val) {
    double v = 0.0;
    bool ok = true;
    for (int h = 0; h < P; h++) {
        const int a = g[h];
        if (a >= 0 && a < A) v += len[a];
        else if (a == -2) v += -2.0;       // GetLengthGenotypes keeps the pad sentinel as a number (:1239-1242)
        else ok = false;
    }
    val = v;
    return ok;
}

template <int KP>
__global__ void __launch_bounds__(kMomThreads) assoc_moments_kernel(AssocParams p) {
    extern __shared__ double zs[];       // [K][256] column-major chunk of z-vectors
    __shared__ int rows[256];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = p.K, nacc = K + 3;
    const int64_t n_loci = p.list ? p.n_list : p.L;
    const int64_t ntiles = (n_loci + kTileLoci - 1) / kTileLoci;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // q = locus ids of this warp's two slots (p.L = none)
        int64_t q[2] = {tile * kTileLoci + warp * 2, tile * kTileLoci + warp * 2 + 1};
        for (int j = 0; j < 2; j++) q[j] = (q[j] < n_loci) ? (p.list ? (int64_t)p.list[q[j]] : q[j]) : p.L;
        const int16_t* grow[2];
        const double* len[2];
        int A[2];
        double off[2];
        for (int j = 0; j < 2; j++) {
            const int64_t l = min(q[j], p.L - 1);
            const int a0 = p.locus_off[l];
            A[j] = p.locus_off[l + 1] - a0;
            len[j] = p.allele_len + a0;
            off[j] = (double)p.P * len[j][0];                 // conditioning shift: P * ref length
            grow[j] = (const int16_t*)((const char*)p.gt + (size_t)l * p.pitch);
        }
        double acc0[KP + 3], acc1[KP + 3];
#pragma unroll
        for (int k = 0; k < KP + 3; k++) { acc0[k] = 0.0; acc1[k] = 0.0; }
        const int64_t s_begin = (int64_t)blockIdx.y * p.seg_len, s_end = min(p.S, s_begin + p.seg_len);
        for (int64_t base = s_begin; base < s_end; base += 256) {
            __syncthreads();
            {   // stage the chunk's z-vectors (zeros for samples outside the design)
                const int64_t s = base + tid;
                const int r = (s < p.S) ? p.row_of_sample[s] : -1;
                rows[tid] = r;
                for (int k = 0; k < K - 1; k++) zs[k * 256 + tid] = (r >= 0) ? p.covars[(int64_t)r * K + k + 1] : 0.0;
                zs[(K - 1) * 256 + tid] = (r >= 0) ? p.outcome[r] : 0.0;
            }
            __syncthreads();
            for (int sub = 0; sub < 8; sub++) {
                const int i = sub * 32 + lane;
                const int64_t s = base + i;
                if (s >= p.S || rows[i] < 0) continue;
                double g0, g1;
                const bool c0 = call_value(grow[0] + s * (p.P + 1), p.P, A[0], len[0], g0) && q[0] < p.L;
                const bool c1 = call_value(grow[1] + s * (p.P + 1), p.P, A[1], len[1], g1) && q[1] < p.L;
                if (!(c0 | c1)) continue;
                g0 = c0 ? g0 - off[0] : 0.0;
                g1 = c1 ? g1 - off[1] : 0.0;
                const double y = zs[(K - 1) * 256 + i];
                acc0[0] += c0 ? 1.0 : 0.0; acc1[0] += c1 ? 1.0 : 0.0;
                acc0[1] += g0;             acc1[1] += g1;
                acc0[2] += g0 * g0;        acc1[2] += g1 * g1;
                acc0[3] += g0 * y;         acc1[3] += g1 * y;
#pragma unroll
                for (int k = 0; k < KP - 1; k++) {
                    if (k < K - 1) {
                        const double c = zs[k * 256 + i];
                        acc0[4 + k] += g0 * c;
                        acc1[4 + k] += g1 * c;
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < KP + 3; k++) {
            if (k < nacc) {
                const double a = warp_sum_d(acc0[k]), b = warp_sum_d(acc1[k]);
                if (lane == 0) {
                    if (p.nseg > 1) {
                        const int64_t i0 = tile * kTileLoci + warp * 2;
                        double* o = p.mom_part + ((size_t)blockIdx.y * n_loci + i0) * nacc + k;
                        if (q[0] < p.L) o[0] = a;
                        if (q[1] < p.L) o[nacc] = b;
                    } else {
                        if (q[0] < p.L) p.mom[q[0] * nacc + k] = a;
                        if (q[1] < p.L) p.mom[q[1] * nacc + k] = b;
                    }
                }
            }
        }
    }
}

template <int KP>
__global__ void __launch_bounds__(256) assoc_downdate_kernel(AssocParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int K = p.K, ne = tri_entries(K);
    constexpr int kR = (KP * (KP + 1) / 2 + 31) / 32;   // entries per lane (17 for K = 32, 5 for K <= 16)
    // lane's entries e = lane + 32 r  ->  (a, b)
    int ea[kR], eb[kR];
#pragma unroll
    for (int r = 0; r < kR; r++) {
        const int e = lane + 32 * r;
        int a = 0, rem = e;
        if (e < ne) {
            while (rem >= K - a) { rem -= K - a; a++; }
        } else {
            rem = 0;
        }
        ea[r] = a;
        eb[r] = a + rem;
    }
    const int64_t n_loci = p.list ? p.n_list : p.L;
    for (int64_t wi = warp; wi < n_loci * p.nseg; wi += nwarps) {
        const int64_t li = wi % n_loci;
        const int seg = (int)(wi / n_loci);
        const int64_t l = p.list ? (int64_t)p.list[li] : li;
        const int a0 = p.locus_off[l];
        const int A = p.locus_off[l + 1] - a0;
        const int16_t* row = (const int16_t*)((const char*)p.gt + (size_t)l * p.pitch);
        double acc[kR];
#pragma unroll
        for (int r = 0; r < kR; r++) acc[r] = 0.0;
        const int64_t s_begin = (int64_t)seg * p.seg_len, s_end = min(p.S, s_begin + p.seg_len);
        for (int64_t sb = s_begin; sb < s_end; sb += 32) {
            const int64_t s = sb + lane;
            int r_row = -1;
            bool uncalled = false;
            if (s < s_end) {
                r_row = p.row_of_sample[s];
                if (r_row >= 0) {
                    const int16_t* g = row + s * (p.P + 1);
                    bool ok = true;
                    for (int h = 0; h < p.P; h++) {
                        const int a = g[h];
                        ok = ok && ((a >= 0 && a < A) || a == -2);
                    }
                    uncalled = !ok;
                }
            }
            unsigned m = __ballot_sync(0xffffffffu, uncalled);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const int rr = __shfl_sync(0xffffffffu, r_row, src);
                // lane k holds z_k of that design row
                double z = 0.0;
                if (lane < K - 1) z = p.covars[(int64_t)rr * K + lane + 1];
                else if (lane == K - 1) z = p.outcome[rr];
#pragma unroll
                for (int r = 0; r < kR; r++) {
                    if (32 * r < ne) {
                        const double za = __shfl_sync(0xffffffffu, z, ea[r]);
                        const double zb = __shfl_sync(0xffffffffu, z, eb[r]);
                        acc[r] += za * zb;
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < kR; r++) {
            const int e = lane + 32 * r;
            if (e < ne) {
                if (p.nseg > 1) p.dd_part[((size_t)seg * n_loci + li) * ne + e] = acc[r];
                else p.dd[l * ne + e] = acc[r];
            }
        }
    }
}

__global__ void assoc_generic_reduce_kernel(AssocParams p, int nacc, int ne) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int w = nacc + ne;
    if (i >= p.n_list * w) return;
    const int64_t li = i / w;
    const int k = (int)(i % w);
    const int64_t l = p.list[li];
    double v = 0.0;
    if (k < nacc) {
        for (int sg = 0; sg < p.nseg; sg++) v += p.mom_part[((size_t)sg * p.n_list + li) * nacc + k];
        p.mom[l * nacc + k] = v;
    } else {
        for (int sg = 0; sg < p.nseg; sg++) v += p.dd_part[((size_t)sg * p.n_list + li) * ne + (k - nacc)];
        p.dd[l * ne + (k - nacc)] = v;
    }
}

// ---- Student t two-sided p-value: I_x(df/2, 1/2), x = df/(df+t^2) ----------------------------------------------
__device__ double betacf(double a, double b, double x) {
    // modified Lentz continued fraction for the incomplete beta function
    const double tiny = 1e-300, eps = 1e-16;
    const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (fabs(d) < tiny) d = tiny;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m <= 100000; m++) {
        const double m2 = 2.0 * m;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c;
        if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c;
        if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < eps) break;
    }
    return h;
}

__device__ double t_two_sided_p(double t, double df) {
    if (isnan(t) || !(df > 0.0)) return nan("");
    if (isinf(t)) return 0.0;
    const double t2 = t * t;
    const double a = 0.5 * df, b = 0.5;
    // ln x = -log1p(t^2/df), ln(1-x) = ln(t^2/(df+t^2))
    const double lnx = -log1p(t2 / df);
    const double x = exp(lnx);
    const double omx = t2 / (df + t2);
    const double lbeta = lgamma(a) + lgamma(b) - lgamma(a + b);
    if (x < (a + 1.0) / (a + b + 2.0)) {
        const double lnpref = a * lnx + b * log(omx) - lbeta;
        return exp(lnpref) * betacf(a, b, x) / a;
    }
    // near t = 0: complement with the roles swapped (converges quickly, result close to 1)
    const double lnpref = b * log(omx) + a * lnx - lbeta;
    const double comp = (omx > 0.0) ? exp(lnpref) * betacf(b, a, omx) / b : 0.0;
    return 1.0 - comp;
}

struct SolveParams {
    int64_t L;
    int K, P;
    const int32_t* locus_off;
    const double* allele_len;
    const int32_t* len_class;
    const int32_t* len_order;
    const int32_t* ac;        // [nA] allele counts over design samples incl. partial calls
    const int32_t* ac_part;   // [nA] the partial-call share
    const double* tot;        // [K(K+1)/2]
    const double* dd;         // [L][K(K+1)/2]
    const double* mom;        // [L][K+3]
    double cutoff;
    int dosage_mode;          // the allele-frequency filters are the caller's (dosage frequencies): only 'n covars >= n samples'
    int32_t* filter_code;
    long long* n_tested;
    double *pval, *coef, *se, *r2, *std_g;
    int32_t* ac_len;          // [nA] counts among tested samples
};

__device__ __forceinline__ int tri_index(int a, int b, int K) {   // a <= b, row-major upper triangle
    return a * K - a * (a - 1) / 2 + (b - a);
}

template <int KP>
__global__ void __launch_bounds__(64) assoc_solve_kernel(SolveParams p) {
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= p.L) return;
    const int K = p.K, Kc = K - 1, ne = tri_entries(K), nacc = K + 3;
    const int a0 = p.locus_off[l];
    const int A = p.locus_off[l + 1] - a0;
    const double* mom = p.mom + l * nacc;
    const double* dd = p.dd + l * ne;
    const double NaN = nan("");
    // ---- allele frequencies over the tested samples, by length rounded to 2 decimals (lafg.py:37-45,174) ----
    long long total = 0;
    if (!p.dosage_mode)
        for (int a = 0; a < A; a++) {
            const int c = p.ac[a0 + a] - p.ac_part[a0 + a];
            if (p.ac_len) p.ac_len[a0 + a] = c;
            total += c;
        }
    const long long n = (long long)llrint(mom[0]);
    int code = TRT_AF_OK;
    if (!p.dosage_mode) {
        // classes in ascending length; keys equal after rounding to 2 dp are merged (consecutive in this order).
        // pass 0: number of classes and the position of the FIRST maximum frequency (np.argmax);
        // pass 1: sum of the other frequencies in dictionary order (af.pop(argmax); np.sum(af)).
        const int32_t* order = p.len_order + a0;
        int nclass = 0, argmax = -1;
        double fmax = -1.0, rest = 0.0;
        for (int pass = 0; pass < 2; pass++) {
            int cls_idx = 0;
            bool have = false;
            double cur_key = 0.0;
            long long cur_cnt = 0;
            for (int i = 0; i <= A; i++) {
                double key = 0.0;
                long long cnt = 0;
                bool flush = (i == A);
                if (i < A) {
                    const int a = order[i];
                    cnt = p.ac[a0 + a] - p.ac_part[a0 + a];
                    key = rint(p.allele_len[a0 + a] * 100.0) / 100.0;
                    if (cnt == 0) continue;
                    if (have && key != cur_key) flush = true;
                }
                if (flush && have) {
                    const double f = (double)cur_cnt / (double)total;
                    if (pass == 0) {
                        if (f > fmax) { fmax = f; argmax = cls_idx; }
                    } else if (cls_idx != argmax) {
                        rest += f;
                    }
                    cls_idx++;
                    have = false;
                }
                if (i < A) {
                    if (!have) { have = true; cur_key = key; cur_cnt = 0; }
                    cur_cnt += cnt;
                }
            }
            nclass = cls_idx;
        }
        if (nclass == 0) code = TRT_AF_NO_CALLED;
        else if (nclass == 1) code = TRT_AF_ONE_ALLELE;
        else if (rest * (double)n * 2.0 < p.cutoff) code = TRT_AF_NON_MAJOR;
    }
    if (code == TRT_AF_OK && (long long)K >= n) code = TRT_AF_NCOVARS;
    p.filter_code[l] = code;
    p.n_tested[l] = n;
    double pval = NaN, coef = NaN, se = NaN, r2 = NaN, sdg = NaN;
    if (code == TRT_AF_OK) {
        const double dn = (double)n;
        const double sg = mom[1], sgg = mom[2], gy_raw = mom[3];
        const double mean_g = sg / dn;
        const double var_g = sgg / dn - mean_g * mean_g;
        sdg = sqrt(var_g > 0.0 ? var_g : 0.0);
        // M = C'C (Kc x Kc), v = C'y, yy over the called rows
        double M[(KP - 1) * (KP - 1)];
        double v[KP - 1], u[KP - 1], ya[KP - 1], yb[KP - 1];
        for (int a = 0; a < Kc; a++) {
            for (int b = a; b < Kc; b++) {
                const int e = tri_index(a, b, K);
                const double x = p.tot[e] - dd[e];
                M[a * Kc + b] = x;
                M[b * Kc + a] = x;
            }
            const int ey = tri_index(a, K - 1, K);
            v[a] = p.tot[ey] - dd[ey];
            u[a] = mom[4 + a];
        }
        const int eyy = tri_index(K - 1, K - 1, K);
        const double yy = p.tot[eyy] - dd[eyy];
        // Cholesky M = L L' (in place, lower)
        bool ok = true;
        for (int j = 0; j < Kc && ok; j++) {
            double s = M[j * Kc + j];
            for (int k = 0; k < j; k++) s -= M[j * Kc + k] * M[j * Kc + k];
            if (!(s > 0.0)) { ok = false; break; }
            const double d = sqrt(s);
            M[j * Kc + j] = d;
            for (int i = j + 1; i < Kc; i++) {
                double t = M[i * Kc + j];
                for (int k = 0; k < j; k++) t -= M[i * Kc + k] * M[j * Kc + k];
                M[i * Kc + j] = t / d;
            }
        }
        if (ok && sdg > 0.0) {
            // solve M a = u, M b = v
            for (int pass = 0; pass < 2; pass++) {
                const double* rhs = pass ? v : u;
                double* x = pass ? yb : ya;
                for (int i = 0; i < Kc; i++) {
                    double t = rhs[i];
                    for (int k = 0; k < i; k++) t -= M[i * Kc + k] * x[k];
                    x[i] = t / M[i * Kc + i];
                }
                for (int i = Kc - 1; i >= 0; i--) {
                    double t = x[i];
                    for (int k = i + 1; k < Kc; k++) t -= M[k * Kc + i] * x[k];
                    x[i] = t / M[i * Kc + i];
                }
            }
            double ua = 0.0, ub = 0.0, vb = 0.0;
            for (int i = 0; i < Kc; i++) { ua += u[i] * ya[i]; ub += u[i] * yb[i]; vb += v[i] * yb[i]; }
            const double ggt = sgg - ua, gyt = gy_raw - ub, yyt = yy - vb;
            const double beta = gyt / ggt;
            const double ssr = yyt - beta * gyt;
            const double df = dn - (double)K;
            const double sigma2 = ssr / df;
            const double se_raw = sqrt(sigma2 / ggt);
            const double sy = v[0];                         // intercept column: sum of y over the called rows
            const double tss = yy - sy * sy / dn;
            coef = beta;
            se = se_raw;
            r2 = 1.0 - ssr / tss;
            pval = t_two_sided_p(beta / se_raw, df);
        }
    }
    p.pval[l] = pval;
    p.coef[l] = coef;
    p.se[l] = se;
    p.r2[l] = r2;
    p.std_g[l] = sdg;
}


// ---- associaTR --beagle-dosages -----------------------------------------------------------------
// load_trs' dosage branch (lafg.py:175-214): per tested sample and haplotype h the dosage of every LENGTH CLASS c
// (alleles whose lengths agree after rounding to two decimals) is d_hc = sum of AP_h over the class' alleles, the
// reference allele taking max(0, 1 - sum(AP_h)) (float32, like numpy); the regressor is g = sum_c len_c (d_1c + d_2c)
// (associaTR.py:266-270).  One warp per locus, lanes stride the sample axis, everything accumulates in FP64.
constexpr int kDosClassBatch = 16;

struct DosageAssocParams {
    AssocParams a;
    const float* ap1;
    const float* ap2;
    const int32_t* cls;          // [nA] class representative (index within the locus) of every allele
    const int32_t* order;        // [nA] per locus: the class representatives in ascending rounded length, then -1
    const double* len_round;     // [nA] python round(length, 2)
    const double* len_around;    // [nA] np.around(length, 2): what the best-guess comparison uses
    double* class_stats;         // [nA][4]
    double* length_stats;        // [L][5]
};

__device__ float np_sum_f32_assoc(const float* a, int n) {
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
    return __fadd_rn(np_sum_f32_assoc(a, n2), np_sum_f32_assoc(a + n2, n - n2));
}

template <int KP>
__global__ void __launch_bounds__(128) assoc_dosage_kernel(DosageAssocParams q) {
    const AssocParams& p = q.a;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int K = p.K, nacc = K + 3;
    for (int64_t l = warp; l < p.L; l += nwarps) {
        const int a0 = p.locus_off[l];
        const int A = p.locus_off[l + 1] - a0;
        const int nalt = A - 1;
        const int16_t* row = (const int16_t*)((const char*)p.gt + (size_t)l * p.pitch);
        const size_t ap_base = (size_t)p.S * (size_t)(a0 - l);
        const int32_t* cls = q.cls + a0;
        const int32_t* order = q.order + a0;
        const double* lr = q.len_round + a0;
        const double* la = q.len_around + a0;
        int ncls = 0;
        while (ncls < A && order[ncls] >= 0) ncls++;
        const double goff = 2.0 * lr[0];                        // conditioning shift (the intercept absorbs it)
        for (int b0 = 0; b0 < ncls || b0 == 0; b0 += kDosClassBatch) {
            const bool first = (b0 == 0);
            double acc[KP + 3], cs[4][kDosClassBatch], ls[5];
#pragma unroll
            for (int k = 0; k < KP + 3; k++) acc[k] = 0.0;
#pragma unroll
            for (int k = 0; k < 5; k++) ls[k] = 0.0;
            for (int i = 0; i < kDosClassBatch; i++) cs[0][i] = cs[1][i] = cs[2][i] = cs[3][i] = 0.0;
            for (int64_t sb = 0; sb < p.S; sb += 32) {
                const int64_t s = sb + lane;
                if (s >= p.S) continue;
                const int r = p.row_of_sample[s];
                if (r < 0) continue;
                // (a block of single-haplotype records — in practice records whose every call is a lone '.' — reads a pad)
                const int h1 = row[s * (p.P + 1)], h2 = (p.P >= 2) ? row[s * (p.P + 1) + 1] : -2;
                const bool ok1 = (h1 >= 0 && h1 < A) || h1 == -2, ok2 = (h2 >= 0 && h2 < A) || h2 == -2;
                if (!(ok1 && ok2)) continue;
                const float* r1 = q.ap1 + ap_base + (size_t)s * nalt;
                const float* r2 = q.ap2 + ap_base + (size_t)s * nalt;
                // np.maximum(0, 1 - sum) propagates NaN (a missing AP entry)
                const float t1 = __fsub_rn(1.0f, np_sum_f32_assoc(r1, nalt)), t2 = __fsub_rn(1.0f, np_sum_f32_assoc(r2, nalt));
                const float ref1 = (t1 != t1) ? t1 : fmaxf(0.0f, t1);
                const float ref2 = (t2 != t2) ? t2 : fmaxf(0.0f, t2);
                const double x1 = (h1 >= 0) ? p.allele_len[a0 + h1] : -2.0, x2 = (h2 >= 0) ? p.allele_len[a0 + h2] : -2.0;
                const double xr1 = (h1 >= 0) ? la[h1] : -2.0, xr2 = (h2 >= 0) ? la[h2] : -2.0;
                double y1 = 0.0, y2 = 0.0, g = 0.0;
                for (int ci = 0; ci < ncls; ci++) {
                    const int c = order[ci];
                    double d1 = 0.0, d2 = 0.0;
                    for (int a = c; a < A; a++) {
                        if (cls[a] != c) continue;
                        d1 += (a == 0) ? (double)ref1 : (double)r1[a - 1];
                        d2 += (a == 0) ? (double)ref2 : (double)r2[a - 1];
                    }
                    const double lenc = lr[c];
                    y1 += lenc * d1;
                    y2 += lenc * d2;
                    g += lenc * (d1 + d2);
                    const int j = ci - b0;
                    if (j >= 0 && j < kDosClassBatch) {
                        const double e1 = (xr1 == lenc) ? 1.0 : 0.0, e2 = (xr2 == lenc) ? 1.0 : 0.0;
                        cs[0][j] += d1 + d2;
                        cs[1][j] += d1 * d1 + d2 * d2;
                        cs[2][j] += e1 + e2;
                        cs[3][j] += e1 * d1 + e2 * d2;
                    }
                }
                if (first) {
                    ls[0] += x1 + x2;
                    ls[1] += x1 * x1 + x2 * x2;
                    ls[2] += y1 + y2;
                    ls[3] += y1 * y1 + y2 * y2;
                    ls[4] += x1 * y1 + x2 * y2;
                    const double gp = g - goff;
                    const double yv = p.outcome[r];
                    acc[0] += 1.0;
                    acc[1] += gp;
                    acc[2] += gp * gp;
                    acc[3] += gp * yv;
#pragma unroll
                    for (int k = 0; k < KP - 1; k++)
                        if (k < K - 1) acc[4 + k] += gp * p.covars[(int64_t)r * K + k + 1];
                }
            }
            if (first) {
#pragma unroll
                for (int k = 0; k < KP + 3; k++) {
                    if (k < nacc) {
                        const double v = warp_sum_d(acc[k]);
                        if (lane == 0) p.mom[l * nacc + k] = v;
                    }
                }
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    const double v = warp_sum_d(ls[k]);
                    if (lane == 0) q.length_stats[l * 5 + k] = v;
                }
            }
            for (int j = 0; j < kDosClassBatch && b0 + j < ncls; j++) {
                const int c = order[b0 + j];
                for (int k = 0; k < 4; k++) {
                    const double v = warp_sum_d(cs[k][j]);
                    if (lane == 0) q.class_stats[(size_t)(a0 + c) * 4 + k] = v;
                }
            }
        }
    }
}

__global__ void fill_rows_kernel(int32_t* row_of_sample, int64_t S) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < S) row_of_sample[s] = -1;
}
__global__ void scatter_rows_kernel(const int32_t* __restrict__ sample_index, int64_t n, int32_t* __restrict__ row_of_sample,
                                    uint8_t* __restrict__ mask) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        row_of_sample[sample_index[i]] = (int32_t)i;
        mask[sample_index[i]] = 1;
    }
}

}  // namespace

extern "C" {

int trt_assoc_set_design(trt_ctx* ctx, const double* covars, const double* outcome, const int32_t* sample_index,
                         int64_t n_design, int K) {
    if (!ctx) return TRT_EINVAL;
    if (!covars || !outcome || !sample_index || n_design < 0 || K < 2 || K > kMaxK)
        return trt_set_error(ctx, TRT_EINVAL, "trt_assoc_set_design: need 2 <= K <= %d design columns", kMaxK);
    TRT_CUDA(cudaSetDevice(ctx->device));
    TRT_TRY(trt_ensure(ctx, ctx->covars, (size_t)n_design * K * 8 + 16));
    TRT_TRY(trt_ensure(ctx, ctx->outcome, (size_t)n_design * 8 + 16));
    TRT_TRY(trt_ensure(ctx, ctx->sample_index, (size_t)n_design * 4 + 16));
    TRT_TRY(trt_ensure(ctx, ctx->assoc_tot, (size_t)tri_entries(K) * 8 + 16));
    if (n_design) {
        TRT_CUDA(cudaMemcpyAsync(ctx->covars.p, covars, (size_t)n_design * K * 8, cudaMemcpyHostToDevice, ctx->stream));
        TRT_CUDA(cudaMemcpyAsync(ctx->outcome.p, outcome, (size_t)n_design * 8, cudaMemcpyHostToDevice, ctx->stream));
        TRT_CUDA(cudaMemcpyAsync(ctx->sample_index.p, sample_index, (size_t)n_design * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    const int ne = tri_entries(K);
    design_totals_kernel<<<ne, 256, 0, ctx->stream>>>((const double*)ctx->covars.p, (const double*)ctx->outcome.p,
                                                                 n_design, K, (double*)ctx->assoc_tot.p);
    TRT_KERNEL_CHECK();
    ctx->n_design = n_design;
    ctx->K = K;
    if (K <= kAssocFastMaxK) TRT_TRY(trt_assoc_mma_design(ctx));
    TRT_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->have_design = true;
    ctx->design_checked_S = -1;
    return TRT_OK;
}

int trt_assoc_ols(trt_ctx* ctx, double non_major_cutoff, trt_assoc_out* out) {
    if (!ctx || !ctx->block_open || !ctx->have_gt || !ctx->harmonized)
        return trt_set_error(ctx, TRT_ESTATE, "trt_assoc_ols: needs a block with GT and trt_harmonize");
    if (!ctx->have_design) return trt_set_error(ctx, TRT_ESTATE, "trt_assoc_ols: call trt_assoc_set_design first");
    if (!out) return trt_set_error(ctx, TRT_EINVAL, "trt_assoc_ols: out is NULL");
    TRT_CUDA(cudaSetDevice(ctx->device));
    const int64_t L = ctx->L, S = ctx->S, nA = ctx->nA, n = ctx->n_design;
    const int K = ctx->K, ne = tri_entries(K), nacc = K + 3;
    // sample -> design row map and the design-membership mask of this block's sample axis
    TRT_TRY(trt_ensure(ctx, ctx->design_row_of_sample, (size_t)S * 4 + 16));
    TRT_TRY(trt_ensure(ctx, ctx->group_masks, (size_t)S + 16));
    TRT_TRY(trt_ensure(ctx, ctx->ac, (size_t)nA * 4 + 16));
    TRT_TRY(trt_ensure(ctx, ctx->ac_part, (size_t)nA * 4 + 16));
    TRT_TRY(trt_ensure(ctx, ctx->lc, (size_t)L * TRT_LC_N * 8 + 16));
    TRT_TRY(trt_ensure(ctx, ctx->assoc_acc, ((size_t)L * (nacc + ne)) * 8 + 16));
    TRT_TRY(trt_ensure(ctx, ctx->assoc_out, (size_t)L * (5 * 8 + 8 + 4) + (size_t)nA * 4 + 64));
    if (ctx->design_checked_S != S) {
        // validate the sample indices against this block's sample axis (once per design and sample count)
        std::vector<int32_t> idx((size_t)n);
        if (n) TRT_CUDA(cudaMemcpyAsync(idx.data(), ctx->sample_index.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        TRT_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int64_t i = 0; i < n; i++)
            if (idx[i] < 0 || idx[i] >= S) return trt_set_error(ctx, TRT_EINVAL, "trt_assoc_ols: design row %lld maps to sample %d outside [0,%lld)", (long long)i, idx[i], (long long)S);
        ctx->design_checked_S = S;
    }
    trt_timer_begin(ctx);
    TRT_CUDA(cudaMemsetAsync(ctx->group_masks.p, 0, (size_t)S + 16, ctx->stream));
    if (S) fill_rows_kernel<<<(unsigned)((S + 255) / 256), 256, 0, ctx->stream>>>((int32_t*)ctx->design_row_of_sample.p, S);
    if (n) scatter_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
        (const int32_t*)ctx->sample_index.p, n, (int32_t*)ctx->design_row_of_sample.p, (uint8_t*)ctx->group_masks.p);
    ctx->launches += 2;
    // allele counts over the design samples (and their partial-call share) via the scan
    TRT_TRY(trt_prepare_ranks(ctx));
    TRT_CUDA(cudaMemsetAsync(ctx->ac.p, 0, (size_t)nA * 4 + 16, ctx->stream));
    TRT_CUDA(cudaMemsetAsync(ctx->ac_part.p, 0, (size_t)nA * 4 + 16, ctx->stream));
    TRT_CUDA(cudaMemsetAsync(ctx->lc.p, 0, (size_t)L * TRT_LC_N * 8 + 16, ctx->stream));
    TRT_CUDA(cudaEventRecord(ctx->ev_s0, ctx->stream));
    if (L > 0) {
        ctx->want_ac_part = true;
        const bool all_samples = (n == S);
        int rc = trt_run_scan(ctx, all_samples ? nullptr : (const uint8_t*)ctx->group_masks.p, 1);
        ctx->want_ac_part = false;
        if (rc != TRT_OK) return rc;
        AssocParams ap;
        ap.gt = ctx->d_gt_active;
        ap.pitch = ctx->gt_active_pitch;
        ap.L = L; ap.S = S; ap.P = ctx->P;
        ap.locus_off = (const int32_t*)ctx->locus_off.p;
        ap.allele_len = (const double*)ctx->allele_len.p;
        ap.row_of_sample = (const int32_t*)ctx->design_row_of_sample.p;
        ap.covars = (const double*)ctx->covars.p;
        ap.outcome = (const double*)ctx->outcome.p;
        ap.K = K;
        ap.mom = (double*)ctx->assoc_acc.p;
        ap.dd = ap.mom + (size_t)L * nacc;
        // ---- fast path (thread per locus, TMA ring) for loci with <= kAssocFastMaxAlleles alleles; the rest
        //      (or everything, for ploidy != 2 / K > 16 / tiny sample counts) through the generic kernels ----
        const bool fast_ok = ctx->P == 2 && K <= kAssocFastMaxK && S >= kAssocFastMinSamples && !getenv("TRT_ASSOC_GENERIC") &&
                             (ctx->gt_active_pitch % 16) == 0 && ((uintptr_t)ctx->d_gt_active % 16) == 0;
        std::vector<int32_t> generic_list;
        ap.list = nullptr;
        ap.n_list = 0;
        bool need_generic = true;
        if (fast_ok) {
            for (int64_t l = 0; l < L; l++)
                if (ctx->h_locus_off[l + 1] - ctx->h_locus_off[l] > kAssocFastMaxAlleles) generic_list.push_back((int32_t)l);
            need_generic = !generic_list.empty();
            if (need_generic) {
                TRT_TRY(trt_ensure(ctx, ctx->assoc_fast_tiles, generic_list.size() * 4 + 16));
                TRT_CUDA(cudaMemcpyAsync(ctx->assoc_fast_tiles.p, generic_list.data(), generic_list.size() * 4, cudaMemcpyHostToDevice,
                                         ctx->stream));
                TRT_CUDA(cudaStreamSynchronize(ctx->stream));   // generic_list is a local
                ap.list = (const int32_t*)ctx->assoc_fast_tiles.p;
                ap.n_list = (int64_t)generic_list.size();
            }
            // integer-form loci on the tensor cores (trt_assoc_mma.cu); what does not fit that form, through the FP64 tiles
            if (trt_assoc_mma_supported(ctx)) {
                int n_fp64 = 0;
                TRT_TRY(trt_assoc_mma(ctx, ap.row_of_sample, ap.mom, ap.dd, &n_fp64));
                if (n_fp64 > 0) TRT_TRY(trt_assoc_fast(ctx, ap.row_of_sample, ap.mom, ap.dd, (const uint8_t*)ctx->assoc_flags.p));
            } else {
                TRT_TRY(trt_assoc_fast(ctx, ap.row_of_sample, ap.mom, ap.dd, nullptr));
            }
        }
        const int64_t n_gen = ap.list ? ap.n_list : L;
        const int64_t ntiles = (n_gen + kTileLoci - 1) / kTileLoci;
        const size_t smem = (size_t)K * 256 * 8;
        const unsigned mgrid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ntiles, (int64_t)ctx->sm_count * 4));
        // a handful of listed loci would otherwise sit on a handful of CTAs for the whole sample axis
        ap.nseg = 1;
        ap.seg_len = (S + 255) & ~int64_t(255);
        ap.mom_part = ap.dd_part = nullptr;
        if (ap.list && need_generic && ntiles < (int64_t)ctx->sm_count * 2 && S >= 4096) {
            int nseg = (int)std::min<int64_t>(((int64_t)ctx->sm_count * 4 + ntiles - 1) / ntiles, S / 2048);
            nseg = std::max(1, std::min(nseg, 4096));
            const int64_t seg_len = (((S + nseg - 1) / nseg) + 255) & ~int64_t(255);
            nseg = (int)((S + seg_len - 1) / seg_len);
            if (nseg > 1) {
                TRT_TRY(trt_ensure(ctx, ctx->assoc_tile_fast, (size_t)nseg * n_gen * (nacc + ne) * 8 + 64));
                ap.nseg = nseg;
                ap.seg_len = seg_len;
                ap.mom_part = (double*)ctx->assoc_tile_fast.p;
                ap.dd_part = ap.mom_part + (size_t)nseg * n_gen * nacc;
            }
        }
        const int64_t wblocks = std::max<int64_t>(1, std::min<int64_t>((n_gen * ap.nseg + 7) / 8, (int64_t)ctx->sm_count * 8));
#define LAUNCH_MOMENTS(KP)                                                                                              \
    do {                                                                                                                \
        if (smem > 48 * 1024)                                                                                           \
            TRT_CUDA(cudaFuncSetAttribute(assoc_moments_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        assoc_moments_kernel<KP><<<dim3(mgrid, (unsigned)ap.nseg), kMomThreads, smem, ctx->stream>>>(ap);               \
        TRT_KERNEL_CHECK();                                                                                             \
        assoc_downdate_kernel<KP><<<(unsigned)wblocks, 256, 0, ctx->stream>>>(ap);                                      \
        TRT_KERNEL_CHECK();                                                                                             \
        if (ap.nseg > 1) {                                                                                              \
            const int64_t nred = n_gen * (nacc + ne);                                                                   \
            assoc_generic_reduce_kernel<<<(unsigned)((nred + 255) / 256), 256, 0, ctx->stream>>>(ap, nacc, ne);         \
            TRT_KERNEL_CHECK();                                                                                         \
        }                                                                                                               \
    } while (0)
        if (!need_generic) {
        } else if (K <= 8) LAUNCH_MOMENTS(8);
        else if (K <= 16) LAUNCH_MOMENTS(16);
        else LAUNCH_MOMENTS(32);
#undef LAUNCH_MOMENTS
    }
    TRT_CUDA(cudaEventRecord(ctx->ev_s1, ctx->stream));
    char* ob = (char*)ctx->assoc_out.p;
    double* o_p = (double*)ob;
    double* o_coef = o_p + L;
    double* o_se = o_coef + L;
    double* o_r2 = o_se + L;
    double* o_sd = o_r2 + L;
    long long* o_n = (long long*)(o_sd + L);
    int32_t* o_code = (int32_t*)(o_n + L);
    int32_t* o_ac = o_code + ((L + 3) & ~int64_t(3));
    if (L > 0) {
        SolveParams sp;
        sp.L = L; sp.K = K; sp.P = ctx->P;
        sp.locus_off = (const int32_t*)ctx->locus_off.p;
        sp.allele_len = (const double*)ctx->allele_len.p;
        sp.len_class = (const int32_t*)ctx->len_class.p;
        sp.len_order = (const int32_t*)ctx->len_order.p;
        sp.ac = (const int32_t*)ctx->ac.p;
        sp.ac_part = (const int32_t*)ctx->ac_part.p;
        sp.tot = (const double*)ctx->assoc_tot.p;
        sp.mom = (const double*)ctx->assoc_acc.p;
        sp.dd = sp.mom + (size_t)L * nacc;
        sp.cutoff = non_major_cutoff;
        sp.dosage_mode = 0;
        sp.filter_code = o_code; sp.n_tested = o_n; sp.pval = o_p; sp.coef = o_coef; sp.se = o_se; sp.r2 = o_r2;
        sp.std_g = o_sd; sp.ac_len = o_ac;
        if (K <= 8) assoc_solve_kernel<8><<<(unsigned)((L + 63) / 64), 64, 0, ctx->stream>>>(sp);
        else if (K <= 16) assoc_solve_kernel<16><<<(unsigned)((L + 63) / 64), 64, 0, ctx->stream>>>(sp);
        else assoc_solve_kernel<32><<<(unsigned)((L + 63) / 64), 64, 0, ctx->stream>>>(sp);
        TRT_KERNEL_CHECK();
    }
    trt_timer_end(ctx);
    {
        float ms = 0.f;
        TRT_CUDA(cudaEventElapsedTime(&ms, ctx->ev_s0, ctx->ev_s1));
        ctx->last_scan_ms = ms;
    }
#define D2H(dst, src, bytes) \
    if ((dst) && (bytes)) TRT_CUDA(cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyDeviceToHost, ctx->stream))
    D2H(out->p, o_p, (size_t)L * 8);
    D2H(out->coef, o_coef, (size_t)L * 8);
    D2H(out->se, o_se, (size_t)L * 8);
    D2H(out->r2, o_r2, (size_t)L * 8);
    D2H(out->std_g, o_sd, (size_t)L * 8);
    D2H(out->n_tested, o_n, (size_t)L * 8);
    D2H(out->filter_code, o_code, (size_t)L * 4);
    D2H(out->ac_len, o_ac, (size_t)nA * 4);
#undef D2H
    TRT_CUDA(cudaStreamSynchronize(ctx->stream));
    return TRT_OK;
}


int trt_assoc_dosage_ols(trt_ctx* ctx, const int32_t* cls, const double* len_round, const double* len_around,
                         trt_assoc_dosage_out* out) {
    if (!ctx || !ctx->block_open || !ctx->have_gt || !ctx->harmonized)
        return trt_set_error(ctx, TRT_ESTATE, "trt_assoc_dosage_ols: needs a block with GT and trt_harmonize");
    if (!ctx->have_design) return trt_set_error(ctx, TRT_ESTATE, "trt_assoc_dosage_ols: call trt_assoc_set_design first");
    if (!ctx->have_ap) return trt_set_error(ctx, TRT_ESTATE, "trt_assoc_dosage_ols: call trt_block_set_ap first");
    if (ctx->P > 2) return trt_set_error(ctx, TRT_EINVAL, "trt_assoc_dosage_ols: Beagle AP1/AP2 dosages are diploid (block ploidy %d)", ctx->P);
    if (!out || !cls || !len_round || !len_around) return trt_set_error(ctx, TRT_EINVAL, "trt_assoc_dosage_ols: NULL argument");
    TRT_CUDA(cudaSetDevice(ctx->device));
    const int64_t L = ctx->L, S = ctx->S, nA = ctx->nA, n = ctx->n_design;
    const int K = ctx->K, ne = tri_entries(K), nacc = K + 3;
    // class order per locus: representatives in ascending rounded length (np.unique of the rounded lengths), then -1
    std::vector<int32_t> order((size_t)nA, -1);
    for (int64_t l = 0; l < L; l++) {
        const int a0 = ctx->h_locus_off[l], A = ctx->h_locus_off[l + 1] - a0;
        std::vector<int> reps;
        for (int a = 0; a < A; a++) {
            if (cls[a0 + a] < 0 || cls[a0 + a] > a) return trt_set_error(ctx, TRT_EINVAL, "trt_assoc_dosage_ols: cls[%d] of locus %lld is not a class representative", a, (long long)l);
            if (cls[a0 + a] == a) reps.push_back(a);
        }
        std::sort(reps.begin(), reps.end(), [&](int x, int y) { return len_round[a0 + x] < len_round[a0 + y]; });
        for (size_t i = 0; i < reps.size(); i++) order[(size_t)a0 + i] = reps[i];
    }
    const size_t meta_bytes = (size_t)nA * (4 + 4 + 8 + 8) + 64;
    TRT_TRY(trt_ensure(ctx, ctx->dos_meta, meta_bytes));
    char* mb = (char*)ctx->dos_meta.p;
    double* d_lr = (double*)mb;
    double* d_la = d_lr + nA;
    int32_t* d_cls = (int32_t*)(d_la + nA);
    int32_t* d_order = d_cls + nA;
    if (nA) {
        TRT_CUDA(cudaMemcpyAsync(d_lr, len_round, (size_t)nA * 8, cudaMemcpyHostToDevice, ctx->stream));
        TRT_CUDA(cudaMemcpyAsync(d_la, len_around, (size_t)nA * 8, cudaMemcpyHostToDevice, ctx->stream));
        TRT_CUDA(cudaMemcpyAsync(d_cls, cls, (size_t)nA * 4, cudaMemcpyHostToDevice, ctx->stream));
        TRT_CUDA(cudaMemcpyAsync(d_order, order.data(), (size_t)nA * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    TRT_CUDA(cudaStreamSynchronize(ctx->stream));          // the caller's arrays and `order` are free again
    TRT_TRY(trt_ensure(ctx, ctx->design_row_of_sample, (size_t)S * 4 + 16));
    TRT_TRY(trt_ensure(ctx, ctx->group_masks, (size_t)S + 16));
    TRT_TRY(trt_ensure(ctx, ctx->assoc_acc, ((size_t)L * (nacc + ne)) * 8 + 16));
    TRT_TRY(trt_ensure(ctx, ctx->assoc_out, (size_t)L * (5 * 8 + 8 + 4) + (size_t)nA * 4 + 64));
    TRT_TRY(trt_ensure(ctx, ctx->dos_out, ((size_t)nA * 4 + (size_t)L * 5) * 8 + 64));
    trt_timer_begin(ctx);
    TRT_CUDA(cudaMemsetAsync(ctx->group_masks.p, 0, (size_t)S + 16, ctx->stream));
    TRT_CUDA(cudaMemsetAsync(ctx->dos_out.p, 0, ((size_t)nA * 4 + (size_t)L * 5) * 8 + 64, ctx->stream));
    if (S) fill_rows_kernel<<<(unsigned)((S + 255) / 256), 256, 0, ctx->stream>>>((int32_t*)ctx->design_row_of_sample.p, S);
    if (n) scatter_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
        (const int32_t*)ctx->sample_index.p, n, (int32_t*)ctx->design_row_of_sample.p, (uint8_t*)ctx->group_masks.p);
    ctx->launches += 2;
    char* ob = (char*)ctx->assoc_out.p;
    double* o_p = (double*)ob;
    double* o_coef = o_p + L;
    double* o_se = o_coef + L;
    double* o_r2 = o_se + L;
    double* o_sd = o_r2 + L;
    long long* o_n = (long long*)(o_sd + L);
    int32_t* o_code = (int32_t*)(o_n + L);
    double* d_cs = (double*)ctx->dos_out.p;
    double* d_ls = d_cs + (size_t)nA * 4;
    TRT_CUDA(cudaEventRecord(ctx->ev_s0, ctx->stream));
    if (L > 0) {
        DosageAssocParams q;
        AssocParams& ap = q.a;
        ap.gt = ctx->d_gt_active;
        ap.pitch = ctx->gt_active_pitch;
        ap.L = L; ap.S = S; ap.P = ctx->P;
        ap.locus_off = (const int32_t*)ctx->locus_off.p;
        ap.allele_len = (const double*)ctx->allele_len.p;
        ap.row_of_sample = (const int32_t*)ctx->design_row_of_sample.p;
        ap.covars = (const double*)ctx->covars.p;
        ap.outcome = (const double*)ctx->outcome.p;
        ap.K = K;
        ap.mom = (double*)ctx->assoc_acc.p;
        ap.dd = ap.mom + (size_t)L * nacc;
        ap.list = nullptr;
        ap.n_list = 0;
        ap.nseg = 1;
        ap.seg_len = (S + 255) & ~int64_t(255);
        ap.mom_part = ap.dd_part = nullptr;
        q.ap1 = (const float*)ctx->ap1.p;
        q.ap2 = (const float*)ctx->ap2.p;
        q.cls = d_cls;
        q.order = d_order;
        q.len_round = d_lr;
        q.len_around = d_la;
        q.class_stats = d_cs;
        q.length_stats = d_ls;
        const int64_t wblocks = std::max<int64_t>(1, std::min<int64_t>((L + 3) / 4, (int64_t)ctx->sm_count * 16));
        const int64_t dblocks = std::max<int64_t>(1, std::min<int64_t>((L + 7) / 8, (int64_t)ctx->sm_count * 8));
#define LAUNCH_DOSAGE(KP)                                                                  \
    do {                                                                                   \
        assoc_dosage_kernel<KP><<<(unsigned)wblocks, 128, 0, ctx->stream>>>(q);            \
        TRT_KERNEL_CHECK();                                                                \
        assoc_downdate_kernel<KP><<<(unsigned)dblocks, 256, 0, ctx->stream>>>(ap);         \
        TRT_KERNEL_CHECK();                                                                \
    } while (0)
        if (K <= 8) LAUNCH_DOSAGE(8);
        else if (K <= 16) LAUNCH_DOSAGE(16);
        else LAUNCH_DOSAGE(32);
#undef LAUNCH_DOSAGE
        SolveParams sp;
        sp.L = L; sp.K = K; sp.P = ctx->P;
        sp.locus_off = (const int32_t*)ctx->locus_off.p;
        sp.allele_len = (const double*)ctx->allele_len.p;
        sp.len_class = (const int32_t*)ctx->len_class.p;
        sp.len_order = (const int32_t*)ctx->len_order.p;
        sp.ac = nullptr;
        sp.ac_part = nullptr;
        sp.tot = (const double*)ctx->assoc_tot.p;
        sp.mom = (const double*)ctx->assoc_acc.p;
        sp.dd = sp.mom + (size_t)L * nacc;
        sp.cutoff = 0.0;
        sp.dosage_mode = 1;
        sp.filter_code = o_code; sp.n_tested = o_n; sp.pval = o_p; sp.coef = o_coef; sp.se = o_se; sp.r2 = o_r2;
        sp.std_g = o_sd; sp.ac_len = nullptr;
        if (K <= 8) assoc_solve_kernel<8><<<(unsigned)((L + 63) / 64), 64, 0, ctx->stream>>>(sp);
        else if (K <= 16) assoc_solve_kernel<16><<<(unsigned)((L + 63) / 64), 64, 0, ctx->stream>>>(sp);
        else assoc_solve_kernel<32><<<(unsigned)((L + 63) / 64), 64, 0, ctx->stream>>>(sp);
        TRT_KERNEL_CHECK();
    }
    TRT_CUDA(cudaEventRecord(ctx->ev_s1, ctx->stream));
    trt_timer_end(ctx);
    {
        float ms = 0.f;
        TRT_CUDA(cudaEventElapsedTime(&ms, ctx->ev_s0, ctx->ev_s1));
        ctx->last_scan_ms = ms;
    }
#define D2H(dst, src, bytes) \
    if ((dst) && (bytes)) TRT_CUDA(cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyDeviceToHost, ctx->stream))
    D2H(out->p, o_p, (size_t)L * 8);
    D2H(out->coef, o_coef, (size_t)L * 8);
    D2H(out->se, o_se, (size_t)L * 8);
    D2H(out->r2, o_r2, (size_t)L * 8);
    D2H(out->std_g, o_sd, (size_t)L * 8);
    D2H(out->n_tested, o_n, (size_t)L * 8);
    D2H(out->ncovars_code, o_code, (size_t)L * 4);
    D2H(out->class_stats, d_cs, (size_t)nA * 4 * 8);
    D2H(out->length_stats, d_ls, (size_t)L * 5 * 8);
#undef D2H
    TRT_CUDA(cudaStreamSynchronize(ctx->stream));
    return TRT_OK;
}

}  // extern "C"
