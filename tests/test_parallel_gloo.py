"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: batch sharding, max-over-ranks timing and the
gradient all-reduce that replaces nn.DataParallel's reduction (reference train.py:172)."""
import os
import socket

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
    launches = parallel.allreduce_grads(params, bucket_bytes=64, average=True)
    span = parallel.shard_range(9, world, rank)
    out.put((rank, tmax, launches, [p.grad.clone() if p.grad is not None else None for p in params], span))
    dist.barrier()
    dist.destroy_process_group()


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
    for rank, tmax, launches, grads, span in results:
        assert tmax == [2.0, 5.0]
        assert launches == 2  # 15*4 + 7*4 >= 64 bytes -> first bucket; the 2x2 tensor flushes at the end
        for i in range(3):
            assert torch.allclose(grads[i], torch.full_like(grads[i], 1.5 * (i + 1)))  # mean of 1x and 2x
        assert grads[3] is None
    assert [r[4] for r in results] == [(0, 5), (5, 9)]
