// associaTR fast path (trt_assoc_tile.cu): thread-per-locus moments over TMA-staged locus x sample tiles.
#pragma once
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

int trt_assoc_fast_zw(int K);
int trt_assoc_fast(trt_ctx* ctx, const int32_t* d_row_of_sample, double* mom, double* dd);
