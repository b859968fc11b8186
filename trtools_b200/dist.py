"""
Multi-GPU plumbing: loci shard by contiguous ranges, one process per GPU (launched by ``torchrun`` /
``python -m torch.distributed.run``, which only supplies RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT).
There is no data-path collective: NCCL only gathers the fixed-width per-locus result rows on rank 0 and
sums dumpSTR's per-sample accumulators.

Two communicators with the same interface:

* :class:`NcclComm` — the product path: the library's own ``trt_dist_*`` entry points (csrc/trt_dist.cu) on the
  context's stream.  Result tables are gathered from DEVICE buffers (no host bounce, no torch): every rank's kernels
  leave their rows in HBM, ``ncclSend/ncclRecv`` moves them to rank 0 and one copy on a side stream brings the gathered
  table to pinned host memory while the next block's kernels already run.  The 128-byte NCCL id travels over a plain
  TCP socket (rank 0 listens on MASTER_PORT + 1 + TRT_RDZV_OFFSET).
* :class:`GlooComm` — ``torch.distributed`` with the gloo backend, CPU only: lets the world-size-2 tests of the host
  logic (sharding, gather order, NaN-poison-preserving sums) run on a box without GPUs.  It computes nothing.
"""
import ctypes as C
import os
import socket
import struct
import time
from typing import List, Optional, Tuple

import numpy as np

REGION_STATS, REGION_ALLELE_COUNTS, REGION_ASSOC, REGION_LOCUS_FILTERS = range(4)


def env_rank_world() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Pin this process to the CPUs NVML reports as local to the GPU, so that the pinned staging blocks allocated
    afterwards (first touch) and the copy threads live on the GPU's NUMA node.  With several ranks streaming host
    blocks at once this keeps each rank's H2D traffic off the inter-socket link.  Returns the number of CPUs bound
    to, or None when NVML / affinity control is unavailable (nothing is changed then)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        n_words = (os.cpu_count() + 63) // 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return None


def locus_shard(n_loci: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous locus range [lo, hi) of ``rank``: concatenating the ranks' outputs restores VCF order."""
    base, rem = divmod(n_loci, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def block_owner(block_index: int, world: int) -> int:
    """The CLIs deal locus BLOCKS round-robin (a file's record count is not known up front); rank 0 restores the
    file order when it merges the ranks' rows (:func:`merge_round_robin`)."""
    return block_index % world


def merge_round_robin(per_rank: List[List[bytes]]) -> List[bytes]:
    """per_rank[r] = the output chunks of the blocks rank r owned, in its own order (block r, r + world, ...).
    Returns all chunks in block order."""
    world = len(per_rank)
    out = []
    n = max((len(x) for x in per_rank), default=0)
    for i in range(n):
        for r in range(world):
            if i < len(per_rank[r]):
                out.append(per_rank[r][i])
    return out


class BlockSharder:
    """How the CLIs (statSTR / dumpSTR / associaTR ``main``) use several GPUs: every rank reads the file, blocks of
    records are dealt round-robin, each rank runs the kernels of its own blocks only and keeps the text it would have
    written; ``finish`` gathers the chunks on rank 0 (NCCL, ragged byte gather) and returns them in file order."""

    def __init__(self, comm):
        self.comm = comm
        self.rank = 0 if comm is None else comm.rank
        self.world = 1 if comm is None else comm.world
        self._next = 0
        self._chunks: List[bytes] = []

    def mine(self) -> bool:
        """Call once per block, in file order, on every rank."""
        own = block_owner(self._next, self.world) == self.rank
        self._next += 1
        return own

    def add(self, text) -> None:
        self._chunks.append(text.encode("utf-8") if isinstance(text, str) else bytes(text))

    def finish(self) -> Optional[List[bytes]]:
        if self.comm is None:
            return list(self._chunks)
        payload = b"".join(struct.pack("<q", len(c)) + c for c in self._chunks)
        parts = self.comm.gather_bytes(payload, 0)
        if parts is None:
            return None
        per_rank = []
        for p in parts:
            lst, off = [], 0
            while off < len(p):
                n = struct.unpack_from("<q", p, off)[0]
                lst.append(p[off + 8:off + 8 + n])
                off += 8 + n
            per_rank.append(lst)
        return merge_round_robin(per_rank)


def cli_comm(ctx):
    """Communicator for a CLI run under torchrun (WORLD_SIZE > 1), else None."""
    return init(ctx) if env_rank_world()[1] > 1 else None


# ---------------------------------------------------------------------------------------------------------------
# rendezvous of the 128-byte NCCL unique id over TCP
# ---------------------------------------------------------------------------------------------------------------
def _rdzv_addr():
    host = os.environ.get("MASTER_ADDR", "127.0.0.1")
    port = int(os.environ.get("MASTER_PORT", "29500")) + 1 + int(os.environ.get("TRT_RDZV_OFFSET", "0"))
    return host, port


def _recv_exact(conn, n):
    buf = b""
    while len(buf) < n:
        chunk = conn.recv(n - len(buf))
        if not chunk:
            raise ConnectionError("rendezvous peer closed the connection")
        buf += chunk
    return buf


def exchange_unique_id(rank: int, world: int, make_id, timeout: float = 120.0) -> bytes:
    """rank 0 creates the id and serves it to the world - 1 other ranks; they retry until rank 0 listens."""
    host, port = _rdzv_addr()
    if rank == 0:
        uid = make_id()
        srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        srv.bind((host if host in ("127.0.0.1", "localhost") else "", port))
        srv.listen(world)
        srv.settimeout(timeout)
        served = set()
        try:
            while len(served) < world - 1:
                conn, _ = srv.accept()
                with conn:
                    peer = struct.unpack("<i", _recv_exact(conn, 4))[0]
                    conn.sendall(uid)
                    served.add(peer)
        finally:
            srv.close()
        return uid
    deadline = time.time() + timeout
    while True:
        try:
            with socket.create_connection((host, port), timeout=5.0) as conn:
                conn.sendall(struct.pack("<i", rank))
                return _recv_exact(conn, 128)
        except (ConnectionError, OSError):
            if time.time() > deadline:
                raise
            time.sleep(0.05)


# ---------------------------------------------------------------------------------------------------------------
# communicators
# ---------------------------------------------------------------------------------------------------------------
class NcclComm:
    """NCCL communicator owned by a ``trt_ctx`` (one per GPU / process)."""

    backend = "nccl"

    def __init__(self, ctx, rank: int, world: int):
        self.ctx, self.rank, self.world = ctx, rank, world
        lib = ctx.lib

        def make_id():
            buf = C.create_string_buffer(128)
            ctx.check(lib.trt_dist_unique_id(buf))
            return buf.raw

        uid = exchange_unique_id(rank, world, make_id)
        ctx.check(lib.trt_dist_init(ctx.h, rank, world, C.create_string_buffer(uid, 128)))

    # -- small host-value collectives ---------------------------------------------------------------------------
    def barrier(self):
        self.ctx.check(self.ctx.lib.trt_dist_barrier(self.ctx.h))

    def allreduce_sum(self, arr: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(arr).copy()
        if a.dtype == np.int64:
            fn = self.ctx.lib.trt_dist_allreduce_sum_i64
        elif a.dtype == np.float64:
            fn = self.ctx.lib.trt_dist_allreduce_sum_f64
        else:
            raise TypeError("allreduce_sum takes int64 or float64 arrays")
        self.ctx.check(fn(self.ctx.h, a.ctypes.data_as(C.c_void_p), a.size))
        return a

    def max(self, value: float) -> float:
        a = np.array([value], dtype=np.float64)
        self.ctx.check(self.ctx.lib.trt_dist_allreduce_max_f64(self.ctx.h, a.ctypes.data_as(C.c_void_p), 1))
        return float(a[0])

    def allgather_i64(self, values) -> np.ndarray:
        """[world, len(values)] int64 (ranks agree on shapes / byte counts with this before a ragged gather)."""
        v = np.ascontiguousarray(values, dtype=np.float64)          # exact for |x| < 2^53
        out = np.empty((self.world, v.size), np.float64)
        self.ctx.check(self.ctx.lib.trt_dist_allgather_f64(self.ctx.h, v.ctypes.data_as(C.c_void_p), v.size,
                                                          out.ctypes.data_as(C.c_void_p)))
        return out.astype(np.int64)

    # -- result tables --------------------------------------------------------------------------------------------
    def gather_region(self, region: int, offset: int, nbytes: int, nbytes_per_rank, dst: int = 0,
                      host_out: Optional[np.ndarray] = None, wait: bool = True):
        """Device-to-device gather of a result region (see include/trtools_b200.h); ``host_out``: uint8 array of
        sum(nbytes_per_rank) bytes on dst (pinned for full copy speed)."""
        counts = np.ascontiguousarray(nbytes_per_rank, dtype=np.int64)
        ptr = None if host_out is None else host_out.ctypes.data_as(C.c_void_p)
        self.ctx.check(self.ctx.lib.trt_dist_gather_region(self.ctx.h, int(region), int(offset), int(nbytes),
                                                           counts.ctypes.data_as(C.c_void_p), int(dst), ptr, 0 if wait else 1))

    def wait(self):
        self.ctx.check(self.ctx.lib.trt_dist_wait(self.ctx.h))

    def gather_bytes(self, payload: bytes, dst: int = 0) -> Optional[List[bytes]]:
        """Ragged gather of host byte strings (the CLIs' formatted rows) -> list per rank on dst."""
        n = len(payload)
        counts = self.allgather_i64([n])[:, 0].copy()
        send = np.frombuffer(payload, dtype=np.uint8) if n else np.zeros(1, np.uint8)
        recv = np.empty(max(int(counts.sum()), 1), np.uint8) if self.rank == dst else None
        self.ctx.check(self.ctx.lib.trt_dist_gather_host(
            self.ctx.h, send.ctypes.data_as(C.c_void_p), n, counts.ctypes.data_as(C.c_void_p), int(dst),
            None if recv is None else recv.ctypes.data_as(C.c_void_p)))
        if self.rank != dst:
            return None
        out, off = [], 0
        for c in counts:
            out.append(recv[off:off + int(c)].tobytes())
            off += int(c)
        return out

    def gather_table(self, local: np.ndarray, dst: int = 0) -> Optional[np.ndarray]:
        local = np.ascontiguousarray(local, dtype=np.float64)
        cols = local.shape[1] if local.ndim == 2 else 1
        parts = self.gather_bytes(local.tobytes(), dst)
        if parts is None:
            return None
        return np.concatenate([np.frombuffer(p, dtype=np.float64).reshape(-1, cols) for p in parts], axis=0)

    def close(self):
        self.ctx.check(self.ctx.lib.trt_dist_finalize(self.ctx.h))


class GlooComm:
    """torch.distributed (gloo) on CPU — test double of :class:`NcclComm` for the host logic."""

    backend = "gloo"

    def __init__(self):
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group("gloo")
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def barrier(self):
        self.dist.barrier()

    def allreduce_sum(self, arr: np.ndarray) -> np.ndarray:
        import torch
        t = torch.from_numpy(np.ascontiguousarray(arr).copy())
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.numpy()

    def max(self, value: float) -> float:
        import torch
        t = torch.tensor([value], dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def allgather_i64(self, values) -> np.ndarray:
        import torch
        v = torch.tensor(list(values), dtype=torch.int64)
        out = [torch.zeros_like(v) for _ in range(self.world)]
        self.dist.all_gather(out, v)
        return np.stack([o.numpy() for o in out])

    def gather_bytes(self, payload: bytes, dst: int = 0) -> Optional[List[bytes]]:
        import torch
        counts = self.allgather_i64([len(payload)])[:, 0]
        pad = int(max(counts.max(), 1))
        buf = torch.zeros(pad, dtype=torch.uint8)
        if payload:
            buf[:len(payload)] = torch.frombuffer(bytearray(payload), dtype=torch.uint8)
        out = [torch.zeros_like(buf) for _ in range(self.world)] if self.rank == dst else None
        self.dist.gather(buf, out, dst=dst)
        if self.rank != dst:
            return None
        return [o[:int(c)].numpy().tobytes() for o, c in zip(out, counts)]

    def gather_table(self, local: np.ndarray, dst: int = 0) -> Optional[np.ndarray]:
        local = np.ascontiguousarray(local, dtype=np.float64)
        cols = local.shape[1] if local.ndim == 2 else 1
        parts = self.gather_bytes(local.tobytes(), dst)
        if parts is None:
            return None
        return np.concatenate([np.frombuffer(p, dtype=np.float64).reshape(-1, cols) for p in parts], axis=0)

    def wait(self):
        pass

    def close(self):
        self.dist.barrier()
        self.dist.destroy_process_group()


def init(ctx=None, backend: Optional[str] = None):
    """Communicator of this process from the torchrun environment, or None when world == 1.  ``ctx`` (a
    ``_lib.Context`` on this rank's GPU) selects the NCCL path; without it (CPU tests) gloo is used."""
    rank, world, _ = env_rank_world()
    if world == 1:
        return None
    if backend is None:
        backend = "nccl" if ctx is not None else "gloo"
    if backend == "nccl":
        if ctx is None:
            raise ValueError("the NCCL communicator needs a trtools_b200 context")
        return NcclComm(ctx, rank, world)
    return GlooComm()


# ---- helpers that accept ``None`` (single process) ---------------------------------------------------------------
def gather_table(comm, local: np.ndarray, dst: int = 0) -> Optional[np.ndarray]:
    """Gather per-locus rows float64 [n_local, C] of every rank on ``dst`` in rank order (row counts may differ)."""
    if comm is None:
        return np.ascontiguousarray(local, dtype=np.float64)
    return comm.gather_table(local, dst)


def allreduce_sum(comm, arr: np.ndarray) -> np.ndarray:
    """Sum an int64 / float64 array over ranks (dumpSTR per-sample accumulators; NaN poison propagates)."""
    return arr if comm is None else comm.allreduce_sum(arr)


def max_over_ranks(comm, value: float) -> float:
    return value if comm is None else comm.max(value)


def gather_bytes(comm, payload: bytes, dst: int = 0) -> Optional[List[bytes]]:
    return [payload] if comm is None else comm.gather_bytes(payload, dst)
