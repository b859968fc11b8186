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
    const uint8_t* mask;   // [S] bytes or null
    int32_t* ac;           // [nA]
    long long* lc;         // [L][TRT_LC_N]
    int fast_enabled;      // the TMA tiers are in use (diploid, enough samples)
};

enum { TIER_PAIRS32 = 0, TIER_PAIRS16 = 1, TIER_WIDE = 2, TIER_GENERIC = 3 };
constexpr int kPairsMaxAllelesU32 = 9;    // (9+3)^2 bins x 256 threads x 4 B = 144 KB
constexpr int kPairsMaxAllelesU16 = 14;   // (14+3)^2 bins x 256 threads x 2 B = 144.5 KB
constexpr int kWideMaxAlleles = 96;       // 96 alleles x 512 threads x 2 B = 96 KB
constexpr int kMinFastSamples = 2048;

__host__ __device__ __forceinline__ int scan_tier(int A) {
    return A <= kPairsMaxAllelesU32 ? TIER_PAIRS32 : (A <= kPairsMaxAllelesU16 ? TIER_PAIRS16 : (A <= kWideMaxAlleles ? TIER_WIDE : TIER_GENERIC));
}

int trt_run_scan(trt_ctx* ctx, const uint8_t* d_mask, int g, int G);
int trt_prepare_ranks(trt_ctx* ctx);
int trt_run_epilogue(trt_ctx* ctx, int use_length, double nalleles_thresh, int G);
