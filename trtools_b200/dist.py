"""
Multi-GPU plumbing: loci shard by contiguous ranges (one process per GPU, ``torch.distributed``); there is no
data-path collective.  NCCL (or gloo on CPU boxes, for the tests) is used only to gather the fixed-width per-locus
result tables on rank 0 and to sum dumpSTR's per-sample accumulators.
"""
import os
from typing import Optional, Tuple

import numpy as np


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


def init(backend: Optional[str] = None):
    """Initialise torch.distributed from the torchrun environment; returns the module (or None when world == 1)."""
    rank, world, local_rank = env_rank_world()
    if world == 1:
        return None
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return dist


def _device(dist):
    import torch
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def gather_table(dist, local: np.ndarray, dst: int = 0) -> Optional[np.ndarray]:
    """Gather per-locus rows float64 [n_local, C] of every rank on ``dst`` in rank order (row counts may differ)."""
    local = np.ascontiguousarray(local, dtype=np.float64)
    if dist is None:
        return local
    import torch
    dev = _device(dist)
    world, rank = dist.get_world_size(), dist.get_rank()
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(x.item()) for x in sizes]
    cols = local.shape[1] if local.ndim == 2 else 1
    pad = max(sizes) if sizes else 0
    buf = torch.zeros((pad, cols), dtype=torch.float64, device=dev)
    if local.shape[0]:
        buf[:local.shape[0]] = torch.from_numpy(local.reshape(local.shape[0], cols)).to(dev)
    out = [torch.zeros_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, out, dst=dst)
    if rank != dst:
        return None
    return np.concatenate([o[:k].cpu().numpy() for o, k in zip(out, sizes)], axis=0)


class GatherPlan:
    """Repeated gather of a fixed-shape per-locus table: float64 [n_cols, n_rows] per rank -> [world, n_cols, n_rows] on
    ``dst``.  Buffers (device send/receive tensors, pinned host result) are allocated once; a gather is then n_cols
    host->device copies, ONE NCCL gather and one device->host copy — the only collective of the whole path."""

    def __init__(self, dist, n_cols: int, n_rows: int, dst: int = 0):
        import torch
        self.dist, self.dst, self.n_cols, self.n_rows = dist, dst, n_cols, n_rows
        self.world = 1 if dist is None else dist.get_world_size()
        self.rank = 0 if dist is None else dist.get_rank()
        if dist is None:
            self.out = np.empty((1, n_cols, n_rows))
            return
        dev = _device(dist)
        # every rank must bring the same shape (bench: weak scaling, fixed loci per rank)
        shape = torch.tensor([n_cols, n_rows], dtype=torch.int64, device=dev)
        lo, hi = shape.clone(), shape.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        if not (torch.equal(lo, shape) and torch.equal(hi, shape)):
            raise ValueError("GatherPlan needs the same table shape on every rank; use gather_table for ragged tables")
        self.send = torch.empty((n_cols, n_rows), dtype=torch.float64, device=dev)
        self.recv = torch.empty((self.world, n_cols, n_rows), dtype=torch.float64, device=dev) if self.rank == dst else None
        pin = dev.type == "cuda"
        self.host = torch.empty((self.world, n_cols, n_rows), dtype=torch.float64, pin_memory=pin) if self.rank == dst else None
        self.out = None if self.host is None else self.host.numpy()

    def gather(self, columns, wait: bool = True):
        """columns: n_cols float64 arrays of length n_rows (host).  Returns [world, n_cols, n_rows] on dst, else None.
        With ``wait=False`` the NCCL gather and the device->host copy of the result stay in flight (call ``wait()``
        before reading ``out``): the next step's kernels overlap them.  The host columns may be reused on return."""
        import torch
        if self.dist is None:
            for j, c in enumerate(columns):
                self.out[0, j] = c
            return self.out
        for j, c in enumerate(columns):
            self.send[j].copy_(torch.from_numpy(np.ascontiguousarray(c, dtype=np.float64)), non_blocking=True)
        if self.send.is_cuda:
            ev = torch.cuda.Event()
            ev.record()
            ev.synchronize()                                   # the host columns have been read
        self.dist.gather(self.send, list(self.recv.unbind(0)) if self.rank == self.dst else None, dst=self.dst)
        if self.rank == self.dst:
            self.host.copy_(self.recv, non_blocking=True)
        if wait:
            return self.wait()
        return None

    def wait(self):
        """Block until the last gather (and its copy to the pinned host table) has finished."""
        import torch
        if self.dist is not None and self.send.is_cuda:
            torch.cuda.current_stream().synchronize()
        return self.out if (self.dist is None or self.rank == self.dst) else None


def allreduce_sum(dist, arr: np.ndarray) -> np.ndarray:
    """Sum an int64 / float64 array over ranks (dumpSTR per-sample accumulators; NaN poison propagates)."""
    if dist is None:
        return arr
    import torch
    t = torch.from_numpy(np.ascontiguousarray(arr)).to(_device(dist))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def max_over_ranks(dist, value: float) -> float:
    if dist is None:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=_device(dist))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
