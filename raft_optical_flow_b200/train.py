"""Training-step component for RAFT with this package's correlation path (scope table 8f, f4).

The reference trains under ``nn.DataParallel`` from one process (train.py:172) with the loop body of
train.py:197-236: forward, ``sequence_loss`` (train.py:47-106), GradScaler-scaled backward, unscale, gradient clipping,
AdamW step, OneCycle schedule (train.py:109-121).  Here the same step runs as ONE PROCESS PER GPU: every rank owns its
batch shard (the correlation path needs no exchange) and the only collective is the parameter-gradient
all-reduce, launched bucket by bucket from backward hooks (``parallel.GradBucketReducer``) so that NCCL runs under the
rest of backward instead of after it.

    step = TrainStep(model, num_steps=100000, iters=12)      # model: the reference RAFT after patch_raft(...)
    loss, metrics = step(image1, image2, flow_gt, valid)     # tensors already on this rank's GPU

Nothing here touches the kernels: it is host logic over torch.distributed, tested on CPU with gloo.
"""
import torch

from . import parallel

__all__ = ["MAX_FLOW", "sequence_loss", "fetch_optimizer", "TrainStep"]

MAX_FLOW = 400.0  # train.py:41: displacements beyond this are excluded from the loss


def sequence_loss(flow_preds, flow_gt, valid, gamma=0.8, max_flow=MAX_FLOW):
    """Exponentially weighted L1 loss over the sequence of flow predictions (reference train.py:47-106).

    flow_preds: list of [N, 2, H, W] (one per GRU iteration), flow_gt [N, 2, H, W], valid [N, H, W].
    loss = sum_i gamma^(n - 1 - i) * mean(mask * |pred_i - gt|) with mask = (valid >= 0.5) & (|gt| < max_flow),
    the mean taken over ALL N*2*H*W elements (masked ones count as zero), exactly as the reference does.
    Returns (loss, metrics); the metrics -- end-point error of the last prediction over the masked pixels and the
    fractions below 1 / 3 / 5 px -- stay 0-dim tensors on the device (``metrics_to_host`` converts them) so that a
    step does not synchronise four times like the reference's ``.item()`` calls."""
    n = len(flow_preds)
    mask = (valid >= 0.5) & (flow_gt.square().sum(dim=1).sqrt() < max_flow)
    w = mask[:, None].to(flow_gt.dtype)
    loss = flow_gt.new_zeros(())
    for i, pred in enumerate(flow_preds):
        loss = loss + (gamma ** (n - 1 - i)) * (w * (pred - flow_gt).abs()).mean()
    with torch.no_grad():
        epe = (flow_preds[-1] - flow_gt).square().sum(dim=1).sqrt()[mask]
        metrics = {"epe": epe.mean(), "1px": (epe < 1).float().mean(), "3px": (epe < 3).float().mean(),
                   "5px": (epe < 5).float().mean()}
    return loss, metrics


def metrics_to_host(metrics):
    return {k: float(v) for k, v in metrics.items()}


def fetch_optimizer(model, lr=4e-4, wdecay=1e-4, epsilon=1e-8, num_steps=100000):
    """AdamW + linear one-cycle schedule with 5 % warm-up over num_steps + 100 steps (reference train.py:109-121)."""
    optimizer = torch.optim.AdamW(model.parameters(), lr=lr, weight_decay=wdecay, eps=epsilon)
    scheduler = torch.optim.lr_scheduler.OneCycleLR(optimizer, lr, num_steps + 100, pct_start=0.05,
                                                    cycle_momentum=False, anneal_strategy="linear")
    return optimizer, scheduler


class TrainStep:
    """One optimisation step (reference train.py:197-236), one process per GPU.

    overlap=True   gradient all-reduce launched from backward hooks, bucket by bucket (GradBucketReducer)
    overlap=False  one bucketed all-reduce after backward has returned (parallel.allreduce_grads)
    With a single process both are no-ops and the step equals the reference's single-GPU step."""

    def __init__(self, model, lr=4e-4, wdecay=1e-4, epsilon=1e-8, num_steps=100000, clip=1.0, gamma=0.8, iters=12,
                 mixed_precision=False, add_noise=False, overlap=True, bucket_bytes=4 << 20):
        self.model = model
        self.clip, self.gamma, self.iters, self.add_noise = clip, gamma, iters, add_noise
        self.optimizer, self.scheduler = fetch_optimizer(model, lr, wdecay, epsilon, num_steps)
        dev = next(model.parameters()).device
        self.scaler = torch.amp.GradScaler(dev.type, enabled=mixed_precision and dev.type == "cuda")
        self.overlap = overlap
        self.reducer = parallel.GradBucketReducer(model.parameters(), bucket_bytes) if overlap else None
        self.allreduce_launches = 0
        self.steps_done = 0

    def __call__(self, image1, image2, flow_gt, valid, forward=None):
        """`forward` (optional) maps (model, image1, image2, iters) to the list of predictions; the default calls the
        model like train.py:209."""
        self.optimizer.zero_grad(set_to_none=True)
        if self.add_noise:  # train.py:204-207
            stdv = float(torch.empty(()).uniform_(0.0, 5.0))
            image1 = (image1 + stdv * torch.randn_like(image1)).clamp(0.0, 255.0)
            image2 = (image2 + stdv * torch.randn_like(image2)).clamp(0.0, 255.0)
        if self.reducer is not None:
            self.reducer.begin_step()
        preds = forward(self.model, image1, image2, self.iters) if forward else self.model(image1, image2, iters=self.iters)
        loss, metrics = sequence_loss(preds, flow_gt, valid, self.gamma)
        self.scaler.scale(loss).backward()  # the reducer's hooks launch the all-reduces from inside this call
        if self.reducer is not None:
            self.allreduce_launches = self.reducer.finish()
        else:
            self.allreduce_launches = parallel.allreduce_grads(self.model.parameters())
        self.scaler.unscale_(self.optimizer)
        torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.clip)
        self.scaler.step(self.optimizer)
        self.scheduler.step()
        self.scaler.update()
        self.steps_done += 1
        return loss.detach(), metrics

    def close(self):
        if self.reducer is not None:
            self.reducer.remove()
