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


_binding = {"desc": "not bound"}


def _gpu_cpu_affinity(local):
    """CPUs the driver reports as local to GPU `local` (NVML), or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(local).uuid)
            handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:  # noqa: BLE001 -- older torch / masked devices: fall back to the index
            handle = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        return cpus or None
    except Exception:  # noqa: BLE001
        return None


def bind_to_gpu_cpus(local, local_world=None):
    """Pins the calling process to its own share of the host CPUs next to GPU `local` (one process per GPU): the CPUs
    NVML reports as local to the GPU (its NUMA node), intersected with what the process may use, and split evenly
    between the ranks of the node so that the launch threads and the pinned staging buffers of different ranks do not
    migrate onto each other.  Call before allocating pinned memory.  Returns the CPU set (also on hosts without NVML,
    where the allowed set is split instead)."""
    if local_world is None:
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
    allowed = sorted(os.sched_getaffinity(0))
    near = _gpu_cpu_affinity(local)
    pool = [c for c in allowed if near is None or c in near] or allowed
    if local_world > 1 and len(pool) >= local_world:
        # ranks that share a pool (same NUMA node) take disjoint slices of it
        per = len(pool) // local_world
        mine = pool[(local % local_world) * per:(local % local_world + 1) * per]
    else:
        mine = pool
    try:
        os.sched_setaffinity(0, mine)
        torch.set_num_threads(max(1, min(len(mine), torch.get_num_threads())))
        _binding["desc"] = (f"{len(mine)} of {len(allowed)} host CPUs ({mine[0]}-{mine[-1]}), "
                            f"{'NVML affinity of the GPU' if near else 'even split of the allowed set'}")
    except OSError as e:
        _binding["desc"] = f"not bound ({e})"
    return mine


def binding_description():
    return _binding["desc"]


def _dense_grads(params):
    """Gradient tensors of every parameter that requires grad, in the given (rank-independent) order; a parameter
    that received no gradient on this rank contributes zeros so that all ranks reduce identical layouts."""
    out = []
    for p in params:
        if not p.requires_grad:
            continue
        if p.grad is None:
            p.grad = torch.zeros_like(p)
        out.append(p.grad)
    return out


def allreduce_grads(params, bucket_bytes=32 << 20, average=True):
    """Sums (or averages) the .grad of `params` over all ranks in flat buckets of ~bucket_bytes, after backward has
    finished (see GradBucketReducer for the overlapped form).  Buckets cover EVERY parameter that requires grad, in
    the given order, with zeros standing in for a missing .grad, and never mix dtypes -- so the layout is identical
    on all ranks whatever each rank's graph touched.  RAFT-full has 5 257 536 fp32 parameters = 21 MB:
    one bucket.  Returns the launch count."""
    params = list(params)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 0
    grads = _dense_grads(params)
    world = dist.get_world_size()
    launches = 0
    bucket, size = [], 0

    def flush():
        nonlocal bucket, size, launches
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
        launches += 1
        bucket, size = [], 0

    for g in grads:
        if bucket and bucket[0].dtype != g.dtype:
            flush()
        bucket.append(g)
        size += g.numel() * g.element_size()
        if size >= bucket_bytes:
            flush()
    flush()
    return launches


class GradBucketReducer:
    """Gradient all-reduce overlapped with backward: the one-process-per-GPU replacement of what nn.DataParallel does
    at the end of backward (reference train.py:172,212).

    Parameters are grouped into fixed buckets of ~bucket_bytes in REVERSE registration order (the order backward
    produces gradients in), one dtype per bucket, identical on every rank.  A post-accumulate-grad hook on every
    parameter counts the bucket down; when a bucket is complete -- and every bucket before it has been launched, so
    that all ranks issue their collectives in the same order even when a rank's graph skips a layer -- its gradients
    are flattened and an asynchronous all_reduce is launched at once (NCCL: on the communicator's own stream, so it
    runs under the rest of backward; gloo in the CPU tests), while autograd keeps producing the earlier layers'
    gradients.  ``finish()`` -- called after ``loss.backward()`` -- launches what is left in order (zeros stand in for
    a parameter that got no gradient this step), waits for the handles and scatters the averaged values back into
    ``.grad``.
    """

    def __init__(self, params, bucket_bytes=4 << 20, average=True):
        self.params = [p for p in params if p.requires_grad]
        self.average = average
        self.enabled = dist.is_initialized() and dist.get_world_size() > 1
        self.buckets = []          # lists of parameters
        cur, size = [], 0
        for p in reversed(self.params):
            if cur and (cur[0].dtype != p.dtype or size >= bucket_bytes):
                self.buckets.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += p.numel() * p.element_size()
        if cur:
            self.buckets.append(cur)
        self._bucket_of = {id(p): i for i, b in enumerate(self.buckets) for p in b}
        self._hooks = []
        if self.enabled:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self.launch_order = []     # bucket indices in the order their all-reduce was launched (last step)
        self._reset()

    def _reset(self):
        self._pending = [len(b) for b in self.buckets]
        self._ready = [set() for _ in self.buckets]
        self._inflight = {}
        self._next = 0  # buckets [0, _next) have been launched

    def _on_grad(self, p):
        i = self._bucket_of[id(p)]
        if id(p) in self._ready[i]:
            return  # a second accumulation into the same .grad within one step (not expected in RAFT)
        self._ready[i].add(id(p))
        self._pending[i] -= 1
        while self._next < len(self.buckets) and self._pending[self._next] == 0:
            self._launch(self._next)

    def _launch(self, i):
        grads = []
        for p in self.buckets[i]:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            grads.append(p.grad)
        flat = torch.cat([g.reshape(-1) for g in grads])
        handle = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True)
        self._inflight[i] = (flat, handle, grads)
        self.launch_order.append(i)
        self._next = i + 1

    def finish(self):
        """Call after backward.  Returns the number of all-reduce launches of this step."""
        if not self.enabled:
            return 0
        for i in range(self._next, len(self.buckets)):  # a parameter of bucket _next received no gradient on this rank
            self._launch(i)
        world = dist.get_world_size()
        for i, (flat, handle, grads) in self._inflight.items():
            handle.wait()
            if self.average:
                flat.div_(world)
            off = 0
            for g in grads:
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()
        n = len(self._inflight)
        self._reset()
        return n

    def begin_step(self):
        self.launch_order = []

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
