// associaTR fast path (trt_assoc_tile.cu): thread-per-locus moments over TMA-staged locus x sample tiles.
#pragma once
#include <cuda.h>

#include <vector>

#include "trt_internal.cuh"

#ifndef TRT_ASSOC_NL
#define TRT_ASSOC_NL 2
#endif
constexpr int kAssocLociPerThread = TRT_ASSOC_NL;            // loci per consumer thread (z-row reuse factor)
constexpr int kAssocTileLoci = 256 * kAssocLociPerThread;    // loci per tile
constexpr int kAssocChunk = (TRT_ASSOC_NL == 1) ? 40 : 24;   // samples per ring stage: 240 B / 144 B rows (odd multiples of 16 B)
constexpr int kAssocFastMaxAlleles = 14;     // thread-private length table: 16 digits x tile loci x 8 B
constexpr int kAssocFastMaxK = 16;           // design columns with a dedicated instantiation
constexpr int kAssocFastMinSamples = 256;

constexpr int kAssocMmaTileLoci = 192;       // loci per tile of the tensor path (trt_assoc_mma.cu): 12 warps x 16

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn trt_tmap_encode_fn();

int trt_assoc_fast_zw(int K);
// FP64 tile path for the loci with flags[l] == 2 (flags == nullptr: every locus with <= kAssocFastMaxAlleles alleles)
int trt_assoc_fast(trt_ctx* ctx, const int32_t* d_row_of_sample, double* mom, double* dd, const uint8_t* flags);
// down-dates of the loci with flags[l] == want (flags == nullptr: <= kAssocFastMaxAlleles alleles) from masks laid out
// [n_tiles][nwin][tile_loci][32] with `bits` samples per word; builds the z-rows table first
int trt_assoc_downdate(trt_ctx* ctx, const int32_t* d_row_of_sample, const uint32_t* masks, int n_tiles, int nwords, int nwin,
                       int tile_loci, int bits, const uint8_t* flags, int want, double* dd);
// tensor path (trt_assoc_mma.cu)
bool trt_assoc_mma_supported(const trt_ctx* ctx);
int trt_assoc_mma_design(trt_ctx* ctx);
int trt_assoc_mma(trt_ctx* ctx, const int32_t* d_row_of_sample, double* mom, double* dd, int* n_fp64);
