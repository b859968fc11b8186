// K3 — dumpSTR call-level and locus-level filters (placeholder until the kernels land).
#include "trt_internal.cuh"

extern "C" {
int trt_call_filters(trt_ctx* ctx, const trt_call_filter_spec*, int, int, trt_call_filter_out*) {
    return trt_set_error(ctx, TRT_ESTATE, "trt_call_filters: not built in this library revision");
}
int trt_locus_filters(trt_ctx* ctx, const trt_locus_filter_spec*, int, int, trt_locus_filter_out*) {
    return trt_set_error(ctx, TRT_ESTATE, "trt_locus_filters: not built in this library revision");
}
}
