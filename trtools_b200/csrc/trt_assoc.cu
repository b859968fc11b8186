// K4 — associaTR per-locus OLS (placeholder until the kernels land).
#include "trt_internal.cuh"

extern "C" {
int trt_assoc_set_design(trt_ctx* ctx, const double*, const double*, const int32_t*, int64_t, int) {
    return trt_set_error(ctx, TRT_ESTATE, "trt_assoc_set_design: not built in this library revision");
}
int trt_assoc_ols(trt_ctx* ctx, double, trt_assoc_out*) {
    return trt_set_error(ctx, TRT_ESTATE, "trt_assoc_ols: not built in this library revision");
}
}
