// Shared declarations of the GT scan (trt_scan.cu) used by the statistics / filter / association units.
#pragma once
#include "trt_internal.cuh"

struct ScanParams {
    const int16_t* gt;
    size_t pitch;
    int64_t L, S;
    int P;
    const int32_t* locus_off;
    const int32_t* len_class;
    const int32_t* seq_class;
    const int32_t* len_rank;
    const int32_t* seq_rank;
    const int32_t* hflags;
    const uint8_t* mask;   // [S] bytes or null: the ONE group the wide / generic tiers count in this launch
    const uint8_t* gbits;  // pair tiers: one byte per sample (zero padded to whole chunks), bit g = "in sample group g"
    int group0;            // pair tiers: first group of this launch (bit and output index)
    size_t ac_stride;      // pair tiers: elements between the groups' ac[] / ac_part[] ...
    size_t lc_stride;      // ... and lc[] blocks
    int32_t* ac;           // [nA] (pair tiers: group 0's)
    int32_t* ac_part;      // optional [nA], zero-initialised: alleles carried by PARTIALLY called samples (a/.)
    long long* lc;         // [L][TRT_LC_N]
    int fast_enabled;      // the TMA tiers are in use (diploid, enough samples)
    int stream_only;       // calibration: consumers only drain the TMA ring (results are meaningless)
    const int32_t* list;   // loci of the tier this launch handles (built on the host per block)
    int n_list;
};

enum { TIER_PAIRS_A = 0, TIER_PAIRS_B = 1, TIER_WIDE = 2, TIER_GENERIC = 3, TIER_COUNT = 4 };
constexpr int kSquareRows = 100;            // ordered digit pairs while (A+3)^2 <= 100 rows (A <= 7)
constexpr int kPairsMainMaxAlleles = 9;     // tier A: A <= 9 (unordered pairs need 78 rows) -> table <= 101 KB
constexpr int kPairsMaxAlleles = 13;        // tier B: unordered pairs, 136 rows x 1 KB (3 ring stages still fit)
constexpr int kWideMaxAlleles = 96;         // 96 alleles x 512 threads x 2 B = 96 KB
constexpr int kMinFastSamples = 2048;

__host__ __device__ __forceinline__ int scan_tier(int A) {
    return A <= kPairsMainMaxAlleles ? TIER_PAIRS_A
           : (A <= kPairsMaxAlleles ? TIER_PAIRS_B : (A <= kWideMaxAlleles ? TIER_WIDE : TIER_GENERIC));
}
__host__ __device__ __forceinline__ bool pairs_square(int A) { return (A + 3) * (A + 3) <= kSquareRows; }
__host__ __device__ __forceinline__ int pairs_rows(int A) {
    const int D = A + 3;
    return pairs_square(A) ? D * D : D * (D + 1) / 2;
}

// the GT scan of every sample group of the block: d_masks = [G][S] device bytes (0/1), or null with G = 1 for all
// samples; results into ctx->ac [G][nA] (+ ctx->ac_part when ctx->want_ac_part) and ctx->lc [G][L][TRT_LC_N]
int trt_run_scan(trt_ctx* ctx, const uint8_t* d_masks, int G);
int trt_prepare_ranks(trt_ctx* ctx);
int trt_run_epilogue(trt_ctx* ctx, int use_length, double nalleles_thresh, int G);
