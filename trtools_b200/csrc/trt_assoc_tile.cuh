// associaTR fast path (trt_assoc_tile.cu): thread-per-locus moments over TMA-staged 256-locus x 40-sample tiles.
#pragma once
#include <vector>

#include "trt_internal.cuh"

constexpr int kAssocTileLoci = 256;          // loci per tile (= consumer threads of the tile kernel)
constexpr int kAssocChunk = 40;              // samples per ring stage: 240 B rows (odd multiple of 16 B)
constexpr int kAssocFastMaxAlleles = 14;     // thread-private length table: 16 digits x 256 threads x 8 B = 32 KB
constexpr int kAssocFastMaxK = 16;           // design columns with a dedicated instantiation
constexpr int kAssocFastMinSamples = 256;

int trt_assoc_fast_zw(int K);
int trt_assoc_fast(trt_ctx* ctx, const int32_t* d_row_of_sample, const std::vector<int32_t>& fast_tiles,
                   const uint8_t* d_tile_fast, double* mom, double* dd);
