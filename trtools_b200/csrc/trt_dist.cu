// Multi-GPU plumbing: loci shard by contiguous ranges (no data-path collective); NCCL only gathers
// the fixed-width per-locus result rows and sums the dumpSTR per-sample counters.
#include <nccl.h>
#include <stdlib.h>

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

static int join_side_stream(trt_ctx* ctx);
static int need_comm(trt_ctx* ctx) {
    if (!ctx || !ctx->nccl_comm) return trt_set_error(ctx, TRT_ESTATE, "trt_dist_*: call trt_dist_init first");
    return TRT_OK;
}

int trt_dist_allgather_f64(trt_ctx* ctx, const double* send_host, int64_t count, double* recv_host) {
    TRT_TRY(need_comm(ctx));
    TRT_TRY(trt_ensure(ctx, ctx->dist_send, (size_t)count * 8 + 16));
    TRT_TRY(trt_ensure(ctx, ctx->dist_recv, (size_t)count * 8 * ctx->world + 16));
    TRT_TRY(join_side_stream(ctx));
    TRT_CUDA(cudaMemcpyAsync(ctx->dist_send.p, send_host, (size_t)count * 8, cudaMemcpyHostToDevice, ctx->stream));
    TRT_NCCL(ncclAllGather(ctx->dist_send.p, ctx->dist_recv.p, (size_t)count, ncclDouble, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    TRT_CUDA(cudaMemcpyAsync(recv_host, ctx->dist_recv.p, (size_t)count * 8 * ctx->world, cudaMemcpyDeviceToHost, ctx->stream));
    TRT_CUDA(cudaStreamSynchronize(ctx->stream));
    return TRT_OK;
}

static int allreduce(trt_ctx* ctx, void* inout_host, int64_t count, ncclDataType_t dt) {
    TRT_TRY(need_comm(ctx));
    TRT_TRY(trt_ensure(ctx, ctx->dist_send, (size_t)count * 8 + 16));
    TRT_TRY(join_side_stream(ctx));
    TRT_CUDA(cudaMemcpyAsync(ctx->dist_send.p, inout_host, (size_t)count * 8, cudaMemcpyHostToDevice, ctx->stream));
    TRT_NCCL(ncclAllReduce(ctx->dist_send.p, ctx->dist_send.p, (size_t)count, dt, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    TRT_CUDA(cudaMemcpyAsync(inout_host, ctx->dist_send.p, (size_t)count * 8, cudaMemcpyDeviceToHost, ctx->stream));
    TRT_CUDA(cudaStreamSynchronize(ctx->stream));
    return TRT_OK;
}

int trt_dist_allreduce_sum_i64(trt_ctx* ctx, int64_t* inout_host, int64_t count) { return allreduce(ctx, inout_host, count, ncclInt64); }
int trt_dist_allreduce_sum_f64(trt_ctx* ctx, double* inout_host, int64_t count) { return allreduce(ctx, inout_host, count, ncclDouble); }

int trt_dist_allreduce_max_f64(trt_ctx* ctx, double* inout_host, int64_t count) {
    TRT_TRY(need_comm(ctx));
    TRT_TRY(trt_ensure(ctx, ctx->dist_send, (size_t)count * 8 + 16));
    TRT_TRY(join_side_stream(ctx));
    TRT_CUDA(cudaMemcpyAsync(ctx->dist_send.p, inout_host, (size_t)count * 8, cudaMemcpyHostToDevice, ctx->stream));
    TRT_NCCL(ncclAllReduce(ctx->dist_send.p, ctx->dist_send.p, (size_t)count, ncclDouble, ncclMax, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    TRT_CUDA(cudaMemcpyAsync(inout_host, ctx->dist_send.p, (size_t)count * 8, cudaMemcpyDeviceToHost, ctx->stream));
    TRT_CUDA(cudaStreamSynchronize(ctx->stream));
    return TRT_OK;
}

int trt_dist_barrier(trt_ctx* ctx) {
    int64_t one = 1;
    return trt_dist_allreduce_sum_i64(ctx, &one, 1);
}

// Side stream for the gathers: the rows of a step travel (NCCL) and reach the host (D2H) while the next step's kernels
// run on the context stream.  The context stream only pays a device-to-device copy of its rows into a staging buffer.
static int ensure_copy_stream(trt_ctx* ctx) {
    if (!ctx->copy_stream) {
        TRT_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        TRT_CUDA(cudaEventCreateWithFlags(&ctx->ev_gathered, cudaEventDisableTiming));
        for (int i = 0; i < 5; i++) {
            TRT_CUDA(cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming));
            for (int b = 0; b < 2; b++) {
                TRT_CUDA(cudaEventCreateWithFlags(&ctx->ev_staged[i][b], cudaEventDisableTiming));
                TRT_CUDA(cudaEventCreateWithFlags(&ctx->ev_sent[i][b], cudaEventDisableTiming));
            }
        }
        TRT_CUDA(cudaEventCreateWithFlags(&ctx->ev_after_scan, cudaEventDisableTiming));
    }
    return TRT_OK;
}

// NCCL operations of one communicator must be issued in one order: before a collective goes onto the context stream,
// that stream waits for everything the side stream still has in flight
static int flush_pending(trt_ctx* ctx, cudaEvent_t after);
static int join_side_stream(trt_ctx* ctx) {
    TRT_TRY(flush_pending(ctx, nullptr));
    if (ctx->copy_stream) {
        TRT_CUDA(cudaEventRecord(ctx->ev_gathered, ctx->copy_stream));
        TRT_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_gathered, 0));
    }
    return TRT_OK;
}

// rank -> dst gather of `nbytes` device bytes per rank (ragged: nbytes_per_rank).  Context stream: rows -> staging buffer
// of the slot.  Side stream: grouped ncclSend / ncclRecv from the staging buffers; on dst the gathered bytes land in the
// slot's receive buffer in rank order and, when host_out is given, are copied to the host.  Nothing blocks the host
// unless `async` is 0.
//
// TRT_DIST_DEFER=1 (experiment, off by default): an asynchronous gather's exchange is not issued at once but by the next
// trt_dist_flush_after_scan (called by trt_locus_stats once its GT scan is queued), gated on an event behind that scan,
// so that the NCCL kernel never shares the SMs with the persistent scan.  Measured at N = 8 (profiles/README.md): 6.27 ms
// per step against 5.92 ms with the exchange issued at once — every rank then waits for the root's late exchange — so
// the default issues it immediately.
static int issue_exchange(trt_ctx* ctx, const trt_pending_gather& g, cudaEvent_t after) {
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    DevBuf& recv = ctx->dist_recv_r[g.slot];
    DevBuf& stage = ctx->dist_stage_r[g.slot][g.buf];
    int64_t total = 0;
    for (int r = 0; r < ctx->world; r++) total += g.per_rank[r];
    TRT_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_staged[g.slot][g.buf], 0));
    if (after) TRT_CUDA(cudaStreamWaitEvent(ctx->copy_stream, after, 0));
    // (stream order already puts this behind the previous gather's host copy out of the same receive buffer)
    TRT_NCCL(ncclGroupStart());
    if (ctx->rank == g.dst) {
        int64_t off = 0;
        for (int r = 0; r < ctx->world; r++) {
            if (r != g.dst && g.per_rank[r] > 0)
                TRT_NCCL(ncclRecv((char*)recv.p + off, (size_t)g.per_rank[r], ncclChar, r, comm, ctx->copy_stream));
            off += g.per_rank[r];
        }
    } else if (g.nbytes > 0) {
        TRT_NCCL(ncclSend(stage.p, (size_t)g.nbytes, ncclChar, g.dst, comm, ctx->copy_stream));
    }
    TRT_NCCL(ncclGroupEnd());
    if (ctx->rank == g.dst) {
        int64_t off = 0;
        for (int r = 0; r < g.dst; r++) off += g.per_rank[r];
        if (g.nbytes > 0)
            TRT_CUDA(cudaMemcpyAsync((char*)recv.p + off, stage.p, (size_t)g.nbytes, cudaMemcpyDeviceToDevice, ctx->copy_stream));
    }
    TRT_CUDA(cudaEventRecord(ctx->ev_sent[g.slot][g.buf], ctx->copy_stream));
    if (ctx->rank == g.dst) {
        if (g.host_out && total > 0)
            TRT_CUDA(cudaMemcpyAsync(g.host_out, recv.p, (size_t)total, cudaMemcpyDeviceToHost, ctx->copy_stream));
        TRT_CUDA(cudaEventRecord(ctx->ev_copied[g.slot], ctx->copy_stream));
    }
    return TRT_OK;
}

static int flush_pending(trt_ctx* ctx, cudaEvent_t after) {
    if (ctx->dist_pending.empty()) return TRT_OK;
    std::vector<trt_pending_gather> todo;
    todo.swap(ctx->dist_pending);
    for (const trt_pending_gather& g : todo) TRT_TRY(issue_exchange(ctx, g, after));
    return TRT_OK;
}

static int gather_device(trt_ctx* ctx, int slot, const void* send_dev, int64_t nbytes, const int64_t* nbytes_per_rank, int dst,
                         void* host_out, int async) {
    TRT_TRY(need_comm(ctx));
    if (dst < 0 || dst >= ctx->world || nbytes < 0 || !nbytes_per_rank || nbytes_per_rank[ctx->rank] != nbytes)
        return trt_set_error(ctx, TRT_EINVAL, "trt_dist_gather: bad dst / byte counts");
    TRT_CUDA(cudaSetDevice(ctx->device));
    TRT_TRY(ensure_copy_stream(ctx));
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    DevBuf& recv = ctx->dist_recv_r[slot];
    const int buf = ctx->dist_seq[slot] & 1;      // staging buffers alternate: staging never waits for the exchange just issued
    DevBuf& stage = ctx->dist_stage_r[slot][buf];
    int64_t total = 0;
    for (int r = 0; r < ctx->world; r++) total += nbytes_per_rank[r];
    // TRT_DIST_MAIN_STREAM=1: the exchange itself stays on the context stream (only the host copy overlaps the next step)
    static const bool side = !getenv("TRT_DIST_MAIN_STREAM");
    static const bool defer = side && getenv("TRT_DIST_DEFER") != nullptr;
    // a deferred gather of this slot still owns the staging buffer: let it go first
    for (const trt_pending_gather& g : ctx->dist_pending)
        if (g.slot == slot || !side) {
            TRT_TRY(flush_pending(ctx, nullptr));
            break;
        }
    if ((size_t)nbytes + 16 > stage.cap || (ctx->rank == dst && (size_t)total + 16 > recv.cap)) {
        TRT_TRY(flush_pending(ctx, nullptr));
        TRT_CUDA(cudaStreamSynchronize(ctx->copy_stream));          // growing a buffer the side stream may still use
        TRT_TRY(trt_ensure(ctx, stage, (size_t)nbytes + 16));
        if (ctx->rank == dst) TRT_TRY(trt_ensure(ctx, recv, (size_t)total + 16));
    }
    if (side) {
        // context stream: the previous send from this staging buffer must be over, then stage the rows
        TRT_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_sent[slot][buf], 0));
        if (nbytes > 0) TRT_CUDA(cudaMemcpyAsync(stage.p, send_dev, (size_t)nbytes, cudaMemcpyDeviceToDevice, ctx->stream));
        TRT_CUDA(cudaEventRecord(ctx->ev_staged[slot][buf], ctx->stream));
        ctx->dist_seq[slot]++;
        trt_pending_gather g;
        g.slot = slot;
        g.buf = buf;
        g.nbytes = nbytes;
        g.per_rank.assign(nbytes_per_rank, nbytes_per_rank + ctx->world);
        g.dst = dst;
        g.host_out = host_out;
        ctx->dist_pending.push_back(g);
        if (!(async && defer)) TRT_TRY(flush_pending(ctx, nullptr));
    } else {
        if (ctx->rank == dst)     // the previous gather's host copy must have left the receive buffer before it is overwritten
            TRT_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[slot], 0));
        TRT_NCCL(ncclGroupStart());
        if (ctx->rank == dst) {
            int64_t off = 0;
            for (int r = 0; r < ctx->world; r++) {
                if (r != dst && nbytes_per_rank[r] > 0)
                    TRT_NCCL(ncclRecv((char*)recv.p + off, (size_t)nbytes_per_rank[r], ncclChar, r, comm, ctx->stream));
                off += nbytes_per_rank[r];
            }
        } else if (nbytes > 0) {
            TRT_NCCL(ncclSend(send_dev, (size_t)nbytes, ncclChar, dst, comm, ctx->stream));
        }
        TRT_NCCL(ncclGroupEnd());
        if (ctx->rank == dst) {
            int64_t off = 0;
            for (int r = 0; r < dst; r++) off += nbytes_per_rank[r];
            if (nbytes > 0)
                TRT_CUDA(cudaMemcpyAsync((char*)recv.p + off, send_dev, (size_t)nbytes, cudaMemcpyDeviceToDevice, ctx->stream));
            TRT_CUDA(cudaEventRecord(ctx->ev_gathered, ctx->stream));
            TRT_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_gathered, 0));
            if (host_out && total > 0)
                TRT_CUDA(cudaMemcpyAsync(host_out, recv.p, (size_t)total, cudaMemcpyDeviceToHost, ctx->copy_stream));
            TRT_CUDA(cudaEventRecord(ctx->ev_copied[slot], ctx->copy_stream));
        }
    }
    if (!async) {
        TRT_CUDA(cudaStreamSynchronize(ctx->stream));
        TRT_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    }
    return TRT_OK;
}

}  // extern "C"

// called by trt_locus_stats right after its scan kernels are queued: deferred exchanges start behind that scan
int trt_dist_flush_after_scan(trt_ctx* ctx) {
    if (!ctx || ctx->dist_pending.empty()) return TRT_OK;
    TRT_CUDA(cudaEventRecord(ctx->ev_after_scan, ctx->stream));
    return flush_pending(ctx, ctx->ev_after_scan);
}

extern "C" {

int trt_dist_gather_region(trt_ctx* ctx, int region, int64_t offset_bytes, int64_t nbytes, const int64_t* nbytes_per_rank,
                           int dst, void* host_out, int async) {
    if (!ctx) return TRT_EINVAL;
    const DevBuf* b = nullptr;
    switch (region) {
        case TRT_REGION_STATS: b = &ctx->stat_f64; break;
        case TRT_REGION_ALLELE_COUNTS: b = &ctx->ac; break;
        case TRT_REGION_ASSOC: b = &ctx->assoc_out; break;
        case TRT_REGION_LOCUS_FILTERS: b = &ctx->misc; break;
        default: return trt_set_error(ctx, TRT_EINVAL, "trt_dist_gather_region: unknown region %d", region);
    }
    if (offset_bytes < 0 || nbytes < 0 || (size_t)(offset_bytes + nbytes) > b->cap)
        return trt_set_error(ctx, TRT_EINVAL, "trt_dist_gather_region: [%lld, +%lld) is outside region %d (%zu bytes)",
                             (long long)offset_bytes, (long long)nbytes, region, b->cap);
    return gather_device(ctx, region, (const char*)b->p + offset_bytes, nbytes, nbytes_per_rank, dst, host_out, async);
}

int trt_dist_gather_host(trt_ctx* ctx, const void* send_host, int64_t nbytes, const int64_t* nbytes_per_rank, int dst,
                         void* recv_host) {
    TRT_TRY(need_comm(ctx));
    if (nbytes < 0 || (nbytes > 0 && !send_host)) return trt_set_error(ctx, TRT_EINVAL, "trt_dist_gather_host: bad send buffer");
    TRT_CUDA(cudaSetDevice(ctx->device));
    TRT_TRY(trt_ensure(ctx, ctx->dist_send, (size_t)nbytes + 16));
    if (nbytes) TRT_CUDA(cudaMemcpyAsync(ctx->dist_send.p, send_host, (size_t)nbytes, cudaMemcpyHostToDevice, ctx->stream));
    return gather_device(ctx, 4, ctx->dist_send.p, nbytes, nbytes_per_rank, dst, recv_host, 0);
}

int trt_dist_wait(trt_ctx* ctx) {
    if (!ctx) return TRT_EINVAL;
    TRT_TRY(flush_pending(ctx, nullptr));
    TRT_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->copy_stream) TRT_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    return TRT_OK;
}

int trt_dist_finalize(trt_ctx* ctx) {
    if (!ctx) return TRT_EINVAL;
    if (ctx->nccl_comm) {
        flush_pending(ctx, nullptr);
        cudaStreamSynchronize(ctx->stream);
        if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
        ncclCommDestroy((ncclComm_t)ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
        ctx->world = 1;
        ctx->rank = 0;
    }
    return TRT_OK;
}

}  // extern "C"
