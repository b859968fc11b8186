// Multi-GPU plumbing: loci shard by contiguous ranges (no data-path collective); NCCL only gathers
// the fixed-width per-locus result rows and sums the dumpSTR per-sample counters.
#include <nccl.h>

#include "trt_internal.cuh"

#define TRT_NCCL(call)                                                                              \
    do {                                                                                            \
        ncclResult_t r__ = (call);                                                                  \
        if (r__ != ncclSuccess)                                                                     \
            return trt_set_error(ctx, TRT_ENCCL, "%s failed: %s", #call, ncclGetErrorString(r__));  \
    } while (0)

extern "C" {

int trt_dist_unique_id(void* out_128_bytes) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return TRT_ENCCL;
    memcpy(out_128_bytes, &id, sizeof(id));
    return TRT_OK;
}

int trt_dist_init(trt_ctx* ctx, int rank, int world, const void* unique_id_128_bytes) {
    if (!ctx || world < 1 || rank < 0 || rank >= world) return trt_set_error(ctx, TRT_EINVAL, "trt_dist_init: bad rank/world");
    TRT_CUDA(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, unique_id_128_bytes, sizeof(id));
    ncclComm_t comm;
    TRT_NCCL(ncclCommInitRank(&comm, world, id, rank));
    ctx->nccl_comm = comm;
    ctx->rank = rank;
    ctx->world = world;
    return TRT_OK;
}

static int need_comm(trt_ctx* ctx) {
    if (!ctx || !ctx->nccl_comm) return trt_set_error(ctx, TRT_ESTATE, "trt_dist_*: call trt_dist_init first");
    return TRT_OK;
}

int trt_dist_allgather_f64(trt_ctx* ctx, const double* send_host, int64_t count, double* recv_host) {
    TRT_TRY(need_comm(ctx));
    TRT_TRY(trt_ensure(ctx, ctx->dist_send, (size_t)count * 8 + 16));
    TRT_TRY(trt_ensure(ctx, ctx->dist_recv, (size_t)count * 8 * ctx->world + 16));
    TRT_CUDA(cudaMemcpyAsync(ctx->dist_send.p, send_host, (size_t)count * 8, cudaMemcpyHostToDevice, ctx->stream));
    TRT_NCCL(ncclAllGather(ctx->dist_send.p, ctx->dist_recv.p, (size_t)count, ncclDouble, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    TRT_CUDA(cudaMemcpyAsync(recv_host, ctx->dist_recv.p, (size_t)count * 8 * ctx->world, cudaMemcpyDeviceToHost, ctx->stream));
    TRT_CUDA(cudaStreamSynchronize(ctx->stream));
    return TRT_OK;
}

static int allreduce(trt_ctx* ctx, void* inout_host, int64_t count, ncclDataType_t dt) {
    TRT_TRY(need_comm(ctx));
    TRT_TRY(trt_ensure(ctx, ctx->dist_send, (size_t)count * 8 + 16));
    TRT_CUDA(cudaMemcpyAsync(ctx->dist_send.p, inout_host, (size_t)count * 8, cudaMemcpyHostToDevice, ctx->stream));
    TRT_NCCL(ncclAllReduce(ctx->dist_send.p, ctx->dist_send.p, (size_t)count, dt, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    TRT_CUDA(cudaMemcpyAsync(inout_host, ctx->dist_send.p, (size_t)count * 8, cudaMemcpyDeviceToHost, ctx->stream));
    TRT_CUDA(cudaStreamSynchronize(ctx->stream));
    return TRT_OK;
}

int trt_dist_allreduce_sum_i64(trt_ctx* ctx, int64_t* inout_host, int64_t count) { return allreduce(ctx, inout_host, count, ncclInt64); }
int trt_dist_allreduce_sum_f64(trt_ctx* ctx, double* inout_host, int64_t count) { return allreduce(ctx, inout_host, count, ncclDouble); }

int trt_dist_barrier(trt_ctx* ctx) {
    int64_t one = 1;
    return trt_dist_allreduce_sum_i64(ctx, &one, 1);
}

}  // extern "C"
