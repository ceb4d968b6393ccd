"""Multi-GPU plumbing for the correlation path: one process per GPU, torch.distributed.

The path itself needs NO collective: every op is per frame pair (reference core/corr.py:121 is a batched
matmul, the pooling / grid_sample calls fold the batch into the leading dim, both alt_cuda_corr kernels use
blockIdx.x = batch), so frame pairs are sharded across ranks and each rank runs CorrBlock on its shard.
The only exchange in the reference's multi-GPU use is the end-of-backward gradient reduction that
nn.DataParallel performs (train.py:172); ``allreduce_grads`` is its one-process-per-GPU replacement
(NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialises the default process group from torchrun's environment.  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kwargs["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kwargs)
    return rank, world, local


def shard_range(n_pairs, world, rank):
    """Contiguous shard [begin, end) of n_pairs frame pairs for `rank`: sizes differ by at most one, earlier
    ranks take the remainder (every pair is owned by exactly one rank; ranks beyond n_pairs get nothing)."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, rem = divmod(n_pairs, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def max_over_ranks(values, device=None):
    """Element-wise maximum of a list of floats over all ranks (timings are reported as the slowest rank)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def allreduce_grads(params, bucket_bytes=32 << 20, average=True):
    """Sums (or averages) the .grad of `params` over all ranks in flat buckets of ~bucket_bytes.
    RAFT-full has 5 257 536 fp32 parameters = 21 MB: one bucket; on NVSwitch the cost is launch
    latency, so buckets are sized for few launches rather than for link count.  Returns the launch count."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not dist.is_initialized() or dist.get_world_size() == 1:
        return 0
    world = dist.get_world_size()
    state = {"launches": 0, "bucket": [], "size": 0}

    def flush():
        bucket = state["bucket"]
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if average:
            flat.div_(world)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        state["launches"] += 1
        state["bucket"], state["size"] = [], 0

    for g in grads:
        state["bucket"].append(g)
        state["size"] += g.numel() * g.element_size()
        if state["size"] >= bucket_bytes:
            flush()
    flush()
    return state["launches"]
