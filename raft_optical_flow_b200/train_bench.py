"""`bench.py --train`: BASELINE.json configs[4] (cfg5) as a whole training step.

The UNMODIFIED reference RAFT-full (random init; core/raft.py from oracle/_ref/reference_raft.tar, staged by
oracle/stage_reference.py where the reference checkout exists) with this package's correlation block and fused
upsampling patched in (``patch_raft``), FlyingChairs shape 368x496, batch 12 per GPU, 12 GRU iterations, driven by
``train.TrainStep``: forward, sequence loss, backward with the gradient all-reduce (NCCL) launched bucket by bucket from
backward hooks, clipping, AdamW + OneCycle (reference train.py:172,197-236).  One process per GPU, weak scaling.

The model, the encoders and the update block are the reference's own code (the callers of the hot path, out of
scope); what this package contributes to the step is the correlation forward + backward, the upsampling and the
reduction, and the line reports the step with the reference's own CorrBlock / upsampling beside it.
"""
import argparse
import json
import os
import statistics
import sys
import tarfile
import tempfile
import warnings


def load_reference_raft(root):
    """The reference's core/raft.py as a module (from the staged archive), or None."""
    tar_path = os.path.join(root, "oracle", "_ref", "reference_raft.tar")
    if not os.path.exists(tar_path):
        return None
    d = tempfile.mkdtemp(prefix="rcb_ref_")
    with tarfile.open(tar_path) as tar:
        tar.extractall(d)
    sys.path.insert(0, os.path.join(d, "core"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import raft as raft_mod  # noqa: E402  (reference core/raft.py)
    return raft_mod


def run(args, cfg, root, sampler_cls=None):
    import torch
    import torch.distributed as dist

    from . import _cabi, parallel, train
    from .upsample import patch_raft

    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) == 0:
            print(json.dumps({"impl": "reference", "unavailable": "the reference has no CPU training path worth timing; "
                              "its GPU step (own CorrBlock, same model) is the gpu_reference key of the repo arm"}))
        return 0
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --train needs a CUDA device")
    warnings.filterwarnings("ignore")
    B, C, Hc, Wc, r, L, iters, desc = cfg
    H, W = 8 * Hc, 8 * Wc
    rank, world, local = parallel.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _cabi.lib()
    parallel.bind_to_gpu_cpus(local)
    raft_mod = load_reference_raft(root)
    if raft_mod is None:
        if rank == 0:
            print(json.dumps({"metric": "train frame pairs/sec", "unavailable": "oracle/_ref/reference_raft.tar not staged"}))
        return 0

    torch.manual_seed(1234)  # train.py:294; the same initial weights on every rank
    model = raft_mod.RAFT(argparse.Namespace(small=False, mixed_precision=False, alternate_corr=False, dropout=0.0)).to(dev)
    model.train()
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    host = [(255 * torch.rand(B, 3, H, W, generator=g)).pin_memory(), (255 * torch.rand(B, 3, H, W, generator=g)).pin_memory(),
            (5 * torch.randn(B, 2, H, W, generator=g)).pin_memory(), torch.ones(B, H, W).pin_memory()]
    resident = [t.to(dev) for t in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step, nsteps, e2e=False):
        """median / all step times in ms over nsteps steps (events on the current stream), after two warm-up steps"""
        for _ in range(2):
            step(*resident)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(nsteps + 1)]
        barrier()
        for k in range(nsteps):
            evs[k].record()
            if e2e:  # every step uploads its batch from pinned host memory and reads the loss back
                batch = [t.to(dev, non_blocking=True) for t in host]
                loss, _ = step(*batch)
                loss.item()
            else:
                step(*resident)
        evs[nsteps].record()
        barrier()
        per = [evs[k].elapsed_time(evs[k + 1]) for k in range(nsteps)]
        total = evs[0].elapsed_time(evs[nsteps])
        return per, total

    nsteps = args.steps
    # 1. the reference's own CorrBlock / upsampling in the same model on this GPU (torch ops), overlapped reduction
    step_ref = train.TrainStep(model, num_steps=100000, iters=iters, overlap=True)
    per_ref, _ = timed(step_ref, max(3, nsteps // 4))
    step_ref.close()
    # 2. ours
    patch_raft(raft_mod)
    step_seq = train.TrainStep(model, num_steps=100000, iters=iters, overlap=False)
    per_seq, _ = timed(step_seq, max(3, nsteps // 4))
    step = train.TrainStep(model, num_steps=100000, iters=iters, overlap=True)
    sampler = sampler_cls(local) if (sampler_cls and rank == 0) else None
    if sampler:  # clocks / power cap state of the timed region of `value` (the sampler needs ~0.5 s to start: the
        sampler.start()  # two warm-up steps inside timed() cover it)
    per, total = timed(step, nsteps)
    clocks = sampler.stop() if sampler else None
    per_e2e, total_e2e = timed(step, nsteps, e2e=True)
    nb = len(step.reducer.buckets) if step.reducer else 0
    launches = step.allreduce_launches

    med, med_e2e, med_ref, med_seq = parallel.max_over_ranks(
        [statistics.median(per), statistics.median(per_e2e), statistics.median(per_ref), statistics.median(per_seq)],
        device=dev)
    mn = parallel.max_over_ranks([min(per)], device=dev)[0]
    if rank == 0:
        nparam = sum(p.numel() for p in model.parameters())
        line = {
            "metric": "train frame pairs/sec", "value": world * B / (med * 1e-3), "unit": "pairs/s", "n_gpus": world,
            "steps": nsteps, "warmup": 2, "ms_per_step": med, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.config}: {desc}, whole training step (fwd + sequence loss + bwd + gradient "
                                   "all-reduce + clip + AdamW/OneCycle) of the unmodified reference RAFT-full, random init",
                       "image": [H, W], "iters": iters, "pairs_per_gpu": B,
                       "parallelism": f"data parallel, {world} x 1 process per GPU, NCCL all-reduce of {4 * nparam} gradient "
                                      f"bytes in {nb} buckets launched from backward hooks"},
            "impl_config": {"build_mode": os.environ.get("RAFT_CORR_MODE", "f16f8"), "cuda_graph": False},
            "timing": {"ms_per_step_median": med, "ms_per_step_min": mn, "timed_steps": nsteps},
            "e2e": {"value": world * B / (med_e2e * 1e-3), "unit": "pairs/s",
                    "h2d_bytes_per_step": sum(t.numel() * 4 for t in host), "d2h_bytes_per_step": 4,
                    "note": "every step uploads its frames / flow / mask from pinned host memory and reads the loss back"},
            "allreduce": {"buckets": nb, "launches_per_step": launches, "overlapped_ms_per_step": med,
                          "after_backward_ms_per_step": med_seq},
            "gpu_reference": {"ms_per_step": med_ref, "value": world * B / (med_ref * 1e-3), "unit": "pairs/s",
                              "note": "same model, same step, the reference's own CorrBlock (torch ops) and upsample_flow"},
            # build (pack + main) + iters x (lookup + upsample) forward; iters x (lookup backward + 2 upsample backward)
            # + (L - 1) pool-backward + 4 packs + 2 GEMMs backward
            "gpu_launches": nsteps * (2 + 2 * iters + 3 * iters + (L - 1) + 6),
            "clocks": clocks,
            "roofline": None,
        }
        print(json.dumps(line))
    step.close()
    step_seq.close()
    if world > 1:
        dist.destroy_process_group()
    return 0
