// Throughput probe for the arithmetic the associaTR moment kernel could run on (sm_100a): vector DFMA, mma.sync f64
// (DMMA) and mma.sync s8 (IMMA).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mma_probe tools/mma_probe.cu
// Prints one line per probe: operations per second with every SM busy (8 warps x 4 CTAs per SM).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kIters = 4096;

__global__ void __launch_bounds__(256) dfma_kernel(double* out, double a, double b) {
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < kIters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) dmma884_kernel(double* out, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < kIters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) dmma16816_kernel(double* out, double a, double b) {
    double c[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < kIters; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%4,%4,%4,%4,%4,%4,%4}, {%5,%5,%5,%5}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) imma_kernel(int* out, unsigned a, unsigned b) {
    int c[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 0; c[i][3] = 1; }
    for (int it = 0; it < kIters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%4,%4,%4}, {%5,%5}, {%0,%1,%2,%3};"
                         : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3]) : "r"(a), "r"(b));
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; i++) launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / 5.0;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    void* buf;
    cudaMalloc(&buf, (size_t)sms * 8 * 256 * 8);
    for (int ctas_per_sm : {1, 2, 4, 8}) {
        const int grid = sms * ctas_per_sm;
        const double warps = (double)grid * 8;
        double ms = time_ms([&] { dfma_kernel<<<grid, 256>>>((double*)buf, 1.0000001, 1e-9); });
        printf("ctas/sm %d  DFMA            %8.2f TFLOP/s\n", ctas_per_sm, warps * 32 * 8 * kIters * 2 / ms / 1e9);
        ms = time_ms([&] { dmma884_kernel<<<grid, 256>>>((double*)buf, 1.0000001, 1e-9); });
        printf("ctas/sm %d  DMMA m8n8k4     %8.2f TFLOP/s\n", ctas_per_sm, warps * 8 * (8.0 * 8 * 4) * kIters * 2 / ms / 1e9);
        ms = time_ms([&] { dmma16816_kernel<<<grid, 256>>>((double*)buf, 1.0000001, 1e-9); });
        printf("ctas/sm %d  DMMA m16n8k16   %8.2f TFLOP/s\n", ctas_per_sm, warps * 4 * (16.0 * 8 * 16) * kIters * 2 / ms / 1e9);
        ms = time_ms([&] { imma_kernel<<<grid, 256>>>((int*)buf, 0x01020304u, 0x01010101u); });
        printf("ctas/sm %d  IMMA m16n8k32   %8.2f TOP/s\n", ctas_per_sm, warps * 8 * (16.0 * 8 * 32) * kIters * 2 / ms / 1e9);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
