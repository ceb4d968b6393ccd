"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: batch sharding, max-over-ranks timing and the
gradient all-reduce that replaces nn.DataParallel's reduction (reference train.py:172)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from raft_optical_flow_b200 import parallel


def test_shard_range_partitions_every_pair_once():
    for n in (0, 1, 7, 8, 16, 33):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(8, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world)
    tmax = parallel.max_over_ranks([1.0 + rank, 5.0 - rank])  # slowest rank wins
    # gradient all-reduce in two buckets; rank r holds grads filled with (r + 1) * (i + 1)
    params = [torch.nn.Parameter(torch.zeros(3, 5)), torch.nn.Parameter(torch.zeros(7)),
              torch.nn.Parameter(torch.zeros(2, 2)), torch.nn.Parameter(torch.zeros(4))]
    for i, p in enumerate(params[:3]):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    if rank == 0:  # only one rank has a gradient for the last parameter (an unused branch on the other)
        params[3].grad = torch.full_like(params[3], 8.0)
    launches = parallel.allreduce_grads(params, bucket_bytes=64, average=True)
    span = parallel.shard_range(9, world, rank)
    reducer_result = _reducer_case(rank, world)
    step_result = {ov: _train_step_case(rank, world, ov) for ov in (True, False)}
    # numpy arrays travel through the queue by value (tensors would be shared by file descriptor and need this
    # process to be alive when the parent unpickles them)
    out.put((rank, tmax, launches, [p.grad.numpy().copy() if p.grad is not None else None for p in params], span,
             reducer_result, step_result))
    dist.barrier()
    dist.destroy_process_group()


class _ToyFlow(torch.nn.Module):
    """Stands in for RAFT: `iters` predictions [N, 2, H, W]."""

    def __init__(self):
        super().__init__()
        self.a = torch.nn.Conv2d(6, 8, 3, padding=1)
        self.b = torch.nn.Conv2d(8, 2, 3, padding=1)

    def forward(self, image1, image2, iters=3):
        f = self.b(torch.tanh(self.a(torch.cat([image1, image2], 1) / 255.0)))
        return [f * (i + 1) / iters for i in range(iters)]


def _toy_batch(world):
    g = torch.Generator().manual_seed(7)
    n = 2 * world
    return (255 * torch.rand(n, 3, 8, 10, generator=g), 255 * torch.rand(n, 3, 8, 10, generator=g),
            torch.randn(n, 2, 8, 10, generator=g), torch.ones(n, 8, 10))


def _train_step_case(rank, world, overlap):
    """train.TrainStep on this rank's shard of a fixed global batch; returns the parameters after two steps."""
    from raft_optical_flow_b200 import train
    torch.manual_seed(3)
    net = _ToyFlow()
    b, e = parallel.shard_range(2 * world, world, rank)
    batch = [t[b:e] for t in _toy_batch(world)]
    step = train.TrainStep(net, num_steps=50, iters=3, overlap=overlap, bucket_bytes=512)
    for _ in range(2):
        loss, _ = step(*batch)
    launches = step.allreduce_launches
    step.close()
    return [p.detach().numpy().copy() for p in net.parameters()], launches


def _reducer_case(rank, world):
    """GradBucketReducer (all-reduce launched from backward hooks, bucket by bucket) against a plain all-reduce of the
    same local gradients; the middle layer is unused on rank 1, so one bucket has to be flushed by finish()."""
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.Tanh(), torch.nn.Linear(16, 16), torch.nn.Tanh(),
                              torch.nn.Linear(16, 3))
    twin = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.Tanh(), torch.nn.Linear(16, 16), torch.nn.Tanh(),
                               torch.nn.Linear(16, 3))
    twin.load_state_dict(net.state_dict())
    x = torch.randn(5, 6, generator=torch.Generator().manual_seed(100 + rank))

    def fwd(m):
        h = m[1](m[0](x))
        if rank == 0:
            h = m[3](m[2](h))
        return m[4](h).square().sum()

    red = parallel.GradBucketReducer(net.parameters(), bucket_bytes=200, average=True)  # one bucket per layer
    results = []
    for step in range(2):  # state must reset between steps
        net.zero_grad(set_to_none=True)
        twin.zero_grad(set_to_none=True)
        red.begin_step()
        fwd(net).backward()
        launched_in_backward = list(red.launch_order)
        n = red.finish()
        fwd(twin).backward()
        for p in twin.parameters():
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            dist.all_reduce(p.grad)
            p.grad /= world
        err = max((a.grad - b.grad).abs().max().item() for a, b in zip(net.parameters(), twin.parameters()))
        results.append((n, len(red.buckets), launched_in_backward, err))
    red.remove()
    return results


def test_two_rank_gloo_allreduce_and_timing():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted((out.get(timeout=180) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process step on the whole batch: what every rank must end up with (equal shards, mean losses)
    from raft_optical_flow_b200 import train
    torch.manual_seed(3)
    whole = _ToyFlow()
    ref_step = train.TrainStep(whole, num_steps=50, iters=3)
    for _ in range(2):
        ref_step(*_toy_batch(world))
    for rank, tmax, launches, grads, span, reducer, steps in results:
        for overlap in (True, False):
            got, n_launch = steps[overlap]
            assert n_launch >= 1
            for a, b in zip(got, whole.parameters()):
                assert np.allclose(a, b.detach().numpy(), atol=2e-6), (overlap, np.abs(a - b.detach().numpy()).max())
        assert tmax == [2.0, 5.0]
        assert launches == 2  # 15*4 + 7*4 >= 64 bytes -> first bucket; the two small tensors flush at the end
        for i in range(3):
            assert np.allclose(grads[i], 1.5 * (i + 1))  # mean of 1x and 2x
        # a gradient that exists on one rank only is averaged against zeros, with identical bucket layouts on all ranks
        assert np.allclose(grads[3], 4.0)
        for n, nbuckets, in_backward, err in reducer:
            assert nbuckets >= 3 and n == nbuckets       # every bucket reduced exactly once per step
            assert len(in_backward) >= 1                 # ... and at least the last layer's already during backward
            assert in_backward == sorted(in_backward)    # buckets complete in reverse layer order
            if rank == 1:
                assert len(in_backward) < nbuckets       # the unused layer's bucket was flushed by finish()
            assert err < 1e-6
    assert [r[4] for r in results] == [(0, 5), (5, 9)]
