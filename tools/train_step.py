"""cfg5 of BASELINE.json as a whole training step: the UNMODIFIED reference RAFT-full (random init; core/raft.py from
oracle/_ref/reference_raft.tar) with this package's correlation block and fused upsampling patched in, FlyingChairs
shape 368x496, batch 12 per GPU, 12 GRU iterations, sequence loss (gamma 0.8, train.py:47-106), backward, NCCL
allreduce of the 21 MB parameter gradient (raft_optical_flow_b200.parallel.allreduce_grads -- the one-process-per-GPU
replacement of nn.DataParallel, train.py:172), gradient clipping and an AdamW step.  One process per GPU:

    python tools/train_step.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/train_step.py

Prints one JSON line from rank 0: step time (CUDA events, max over ranks) with the reference CorrBlock (torch ops) and
with ours, samples/s over all ranks, the allreduce launch count."""
import argparse, json, os, sys, tarfile, tempfile, warnings
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raft_optical_flow_b200 as rcb  # noqa: E402
from raft_optical_flow_b200 import parallel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=12)
ap.add_argument("--iters", type=int, default=12)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--size", type=int, nargs=2, default=[368, 496])
a = ap.parse_args()
rank, world, local = parallel.init_from_env("nccl")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
warnings.filterwarnings("ignore")
tmp = tempfile.mkdtemp()
with tarfile.open(os.path.join(ROOT, "oracle", "_ref", "reference_raft.tar")) as t:
    t.extractall(tmp)
sys.path.insert(0, os.path.join(tmp, "core"))
import raft as raft_mod  # noqa: E402  (reference core/raft.py)

torch.manual_seed(1234)  # train.py:294; same initial weights on every rank
model = raft_mod.RAFT(argparse.Namespace(small=False, mixed_precision=False, alternate_corr=False, dropout=0.0)).to(dev)
model.train()
opt = torch.optim.AdamW(model.parameters(), lr=4e-4, weight_decay=1e-4, eps=1e-8)  # train.py:109-121
g = torch.Generator(device="cpu").manual_seed(1234 + rank)
H, W = a.size
im1 = (255 * torch.rand(a.batch, 3, H, W, generator=g)).to(dev)
im2 = (255 * torch.rand(a.batch, 3, H, W, generator=g)).to(dev)
flow_gt = (5 * torch.randn(a.batch, 2, H, W, generator=g)).to(dev)
valid = torch.ones(a.batch, H, W, device=dev)


def sequence_loss(preds, gt, valid, gamma=0.8, max_flow=400):  # the arithmetic of train.py:47-106
    mag = torch.sum(gt ** 2, dim=1).sqrt()
    v = (valid >= 0.5) & (mag < max_flow)
    loss = 0.0
    for i, p in enumerate(preds):
        loss = loss + gamma ** (len(preds) - i - 1) * (v[:, None] * (p - gt).abs()).mean()
    return loss


def step():
    opt.zero_grad(set_to_none=True)
    loss = sequence_loss(model(im1, im2, iters=a.iters), flow_gt, valid)
    loss.backward()
    n = parallel.allreduce_grads(list(model.parameters()))
    torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)  # train.py:216
    opt.step()
    return loss, n


def timed():
    step(); step()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss, n = step()
    e1.record()
    torch.cuda.synchronize()
    return parallel.max_over_ranks([e0.elapsed_time(e1) / a.steps], device=dev)[0], loss.item(), n


ms_ref, loss_ref, _ = timed()
old = rcb.patch_raft(raft_mod)
ms_ours, loss_ours, launches = timed()
if rank == 0:
    print(json.dumps({"config": f"cfg5: RAFT-full {H}x{W}, batch {a.batch}/GPU, {a.iters} iters, fwd+bwd+allreduce+AdamW",
                      "n_gpus": world, "ms_per_step_reference_corr": round(ms_ref, 2), "ms_per_step_ours": round(ms_ours, 2),
                      "samples_per_s_ours": round(world * a.batch / (ms_ours * 1e-3), 1),
                      "samples_per_s_reference_corr": round(world * a.batch / (ms_ref * 1e-3), 1),
                      "allreduce_launches_per_step": launches, "grad_bytes": 4 * sum(p.numel() for p in model.parameters()),
                      "note": "the two arms continue one optimisation run (reference blocks first), so their losses are "
                              "not comparable; gradient parity is tests/test_gpu_e2e_raft.py"}))
if world > 1:
    torch.distributed.destroy_process_group()
