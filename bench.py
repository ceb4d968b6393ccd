#!/usr/bin/env python
"""Benchmark of the RAFT correlation hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2] [--mode bf16x3]

A step = one pass of the hot path over one batch of frame pairs: build the correlation pyramid from
(fmap1, fmap2), then `iters` window lookups with a fresh coords tensor each (what core/raft.py:186-219 does
per forward).  Default workload = BASELINE.json configs[1]: RAFT-full, Sintel 440x1024 -> 55x128 grid,
C=256, radius 4, 4 levels, batch 8 per GPU, 32 iterations.  Multi-GPU: one process per GPU (torchrun), each
rank owns its own batch shard (weak scaling), no data-path collective; value = all pairs / max-over-ranks time.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (B per GPU, C, H, W, radius, levels, iters, description)
    "cfg1": (1, 128, 55, 128, 3, 4, 12, "RAFT-small Sintel 436x1024(->440x1024), 12 iters"),
    "cfg2": (8, 256, 55, 128, 4, 4, 32, "RAFT-full Sintel 440x1024, batch 8/GPU, 32 iters"),
    "cfg3": (16, 256, 47, 156, 4, 4, 24, "RAFT-full KITTI 376x1248, batch 16/GPU, 24 iters"),
    "cfg4": (4, 256, 136, 240, 4, 4, 32, "RAFT-full 1088x1920 (1080p padded), batch 4/GPU, 32 iters"),
    "cfg5": (12, 256, 46, 62, 4, 4, 12, "RAFT-full FlyingChairs 368x496, batch 12/GPU, 12 iters"),
}
SEED = 1234  # the reference's own seed (train.py:294)


def algorithmic_bytes(B, C, H, W, r, L):
    """build = fmaps in + fp32 pyramid out; lookup per query = unique taps + outputs + coords."""
    Q = H * W
    hs, ws = [H], [W]
    for _ in range(L - 1):
        hs.append(hs[-1] // 2)
        ws.append(ws[-1] // 2)
    build = 2 * B * C * Q * 4 + 4 * B * Q * sum(h * w for h, w in zip(hs, ws))
    lookup = B * Q * (4 * L * (2 * r + 2) ** 2 + 4 * L * (2 * r + 1) ** 2 + 8)
    flops = 2.0 * B * Q * Q * C
    return build, lookup, flops


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p["hbm_gbs"], p.get("bf16_tflops_sustained", p["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, "fallback"  # B200_PROFILING.md


def ncu_traffic(kernel):
    """Per-launch DRAM bytes of `kernel` from the committed ncu capture, if one has been summarised."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get(kernel)
    return None


class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.file = index, None, None

    def start(self):
        try:
            self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=self.file, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.file.flush()
        self.file.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.file.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        self.file.close()
        os.unlink(self.file.name)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_step(B, C, H, W, r, L, iters, seed):
    """One bounded sample of the workload on the host cores with the CPU oracle port (oracle/).
    Returns seconds for: volume (fp32 accumulate, like the reference's SGEMM) + pooled pyramid + `iters` lookups."""
    import numpy as np
    from oracle import oracle as orc
    rs = np.random.RandomState(seed)
    f1 = (0.75 * rs.standard_normal((B, C, H, W))).astype(np.float32)
    f2 = (0.75 * rs.standard_normal((B, C, H, W))).astype(np.float32)
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    grid = np.stack([xs, ys])[None].astype(np.float32)
    coords = [(grid + 4.0 * rs.standard_normal((B, 2, H, W))).astype(np.float32) for _ in range(iters)]
    t0 = time.perf_counter()
    blk = orc.OracleCorrBlock(f1, f2, num_levels=L, radius=r, acc64=False)
    t1 = time.perf_counter()
    for c in coords:
        blk(c)
    t2 = time.perf_counter()
    return t2 - t0, t1 - t0, (t2 - t1) / iters


def run_reference(args, cfg):
    """--impl reference: the reference's CPU implementation of the path, restated in oracle/ (the reference is
    Python whose arithmetic lives in torch CPU ops; /root/reference does not exist on the GPU box), all host
    threads, one full batch per step (about 1-2 s of CPU work on 8 cores)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    B, C, H, W, r, L, iters, desc = cfg
    from oracle import oracle as orc
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm is meant to use every host core it can
    orc.set_num_threads(len(os.sched_getaffinity(0)))
    cores = orc.num_threads()
    sample_pairs = B  # the whole batch: a step is ~1-2 s of CPU work on 8 cores
    for _ in range(args.warmup):
        cpu_reference_step(sample_pairs, C, H, W, r, L, max(1, iters // 8), SEED)
    t = 0.0
    for k in range(args.steps):
        dt, _, _ = cpu_reference_step(sample_pairs, C, H, W, r, L, iters, SEED + k)
        t += dt
    value = sample_pairs * args.steps / t
    line = {
        "impl": "reference", "metric": "corr pairs/sec", "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}: {desc}", "C": C, "grid": [H, W], "radius": r, "levels": L,
                   "iters": iters, "pairs_per_step": sample_pairs},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port",
                         "sample": f"{sample_pairs} frame pairs per step (the full batch), build + {iters} lookups, "
                                   f"oracle/corr_oracle.c with OpenMP on {cores} threads"},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def run_ours(args, cfg):
    import torch
    import torch.distributed as dist
    from raft_optical_flow_b200 import CorrBlock, _cabi, parallel

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: raft_optical_flow_b200 has no CPU fallback")
    B, C, H, W, r, L, iters, desc = cfg
    rank, world, local = parallel.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _cabi.lib()  # fail loudly if the extension is missing

    g = torch.Generator(device="cpu").manual_seed(SEED + rank)
    nsets = 2  # alternate input sets; the 2.2 GB pyramid written every step flushes the 126 MB L2 anyway
    host_f = [(0.75 * torch.randn(2, B, C, H, W, generator=g)).pin_memory() for _ in range(nsets)]
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    grid = torch.stack([xs, ys]).float()[None]
    host_c = (grid + 4.0 * torch.randn(iters, B, 2, H, W, generator=g)).pin_memory()
    dev_f = [h.to(dev) for h in host_f]
    dev_c = host_c.to(dev)
    host_out = torch.empty((B, L * (2 * r + 1) ** 2, H, W), dtype=torch.float32).pin_memory()

    def step_resident(k, ev=None, mode=None, pdt=None):
        f = dev_f[k % nsets]
        if ev:
            ev[0].record()
        blk = CorrBlock(f[0], f[1], num_levels=L, radius=r, mode=mode or args.mode, pyramid_dtype=pdt or args.pyramid)
        if ev:
            ev[1].record()
        out = None
        for i in range(iters):
            out = blk(dev_c[i])
        if ev:
            ev[2].record()
        return out

    # End to end: host buffers in, host result out, through the public CorrBlock API.  Three streams, inputs
    # double-buffered on the device: the H2D copy of step k+1 and the D2H copy of step k-1 overlap the kernels of
    # step k (PCIe is full duplex); every step still copies all of its inputs and its result inside the timed
    # region.
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    s_main = torch.cuda.current_stream(dev)
    in_bufs = [(torch.empty_like(dev_f[0]), torch.empty_like(dev_c)) for _ in range(2)]
    in_ready = [torch.cuda.Event() for _ in range(2)]
    in_free = [torch.cuda.Event() for _ in range(2)]
    out_bufs = [torch.empty((B, L * (2 * r + 1) ** 2, H, W), dtype=torch.float32, device=dev) for _ in range(2)]
    out_ready = [torch.cuda.Event() for _ in range(2)]
    out_free = [torch.cuda.Event() for _ in range(2)]

    def e2e_upload(k):
        j = k % 2
        with torch.cuda.stream(s_in):
            s_in.wait_event(in_free[j])  # step k-2 no longer reads this slot
            in_bufs[j][0].copy_(host_f[k % nsets], non_blocking=True)
            in_bufs[j][1].copy_(host_c, non_blocking=True)
            in_ready[j].record(s_in)

    def e2e_compute(k):
        j = k % 2
        s_main.wait_event(in_ready[j])
        f, c = in_bufs[j]
        blk = CorrBlock(f[0], f[1], num_levels=L, radius=r, mode=args.mode, pyramid_dtype=args.pyramid)
        out = None
        for i in range(iters):
            out = blk(c[i])
        in_free[j].record(s_main)
        s_main.wait_event(out_free[j])  # the D2H of step k-2 has drained this slot
        out_bufs[j].copy_(out)
        out_ready[j].record(s_main)
        with torch.cuda.stream(s_out):
            s_out.wait_event(out_ready[j])
            host_out.copy_(out_bufs[j], non_blocking=True)
            out_free[j].record(s_out)

    def run_e2e(nsteps):
        e2e_upload(0)
        for k in range(nsteps):
            if k + 1 < nsteps:
                e2e_upload(k + 1)
            e2e_compute(k)
        s_main.wait_stream(s_out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    t_load = time.perf_counter()
    for k in range(args.warmup):
        step_resident(k)
    while time.perf_counter() - t_load < 0.6:  # nvidia-smi needs a few hundred ms to deliver its first samples
        step_resident(0)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    end = torch.cuda.Event(enable_timing=True)
    barrier()
    for k in range(args.steps):
        step_resident(k, evs[k])
    end.record()
    barrier()
    total_ms = evs[0][0].elapsed_time(end)
    build_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    lookup_ms = sum(e[1].elapsed_time(e[2]) for e in evs) / (args.steps * iters)

    # ---- fast mode, reported next to the headline (never instead of it): single bf16 pass + fp16-stored pyramid,
    # the "within a stated bound" path of the spec (flow EPE delta <= 0.01 px, tests/test_gpu_e2e_raft.py)
    fast = None
    if args.pyramid == "f32" and args.mode == "bf16x3" and not args.no_fast_mode:
        for k in range(3):
            step_resident(k, mode="bf16", pdt="f16")
        fevs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(min(args.steps, 5))]
        fend = torch.cuda.Event(enable_timing=True)
        barrier()
        for k in range(len(fevs)):
            step_resident(k, fevs[k], mode="bf16", pdt="f16")
        fend.record()
        barrier()
        f_total = parallel.max_over_ranks([fevs[0][0].elapsed_time(fend)], device=dev)[0]
        fast = {"build_mode": "bf16", "pyramid_dtype": "f16", "value": world * B * len(fevs) / (f_total * 1e-3),
                "unit": "pairs/s", "ms_per_step": f_total / len(fevs),
                "build_us": 1e3 * sum(e[0].elapsed_time(e[1]) for e in fevs) / len(fevs),
                "lookup_us": 1e3 * sum(e[1].elapsed_time(e[2]) for e in fevs) / (len(fevs) * iters),
                "note": "operands rounded to bf16 (2.6e-3 of max-abs), pyramid stored as fp16; "
                        "final-flow EPE delta vs the reference 1e-3 px mean on the demo frames"}

    # ---- next row of the scope table (8f, f1), reported next to the headline: the lookup fused with the motion
    # encoder's first layer, against this package's lookup followed by the convolution + ReLU the reference runs
    fused = None
    if args.pyramid == "f32" and r in (3, 4) and not args.no_fast_mode:
        import torch.nn.functional as F
        from raft_optical_flow_b200 import PackedConvC1
        cout, cin = (256 if r == 4 else 96), L * (2 * r + 1) ** 2
        gw = torch.Generator(device="cpu").manual_seed(SEED + 1)
        wgt = (torch.randn(cout, cin, 1, 1, generator=gw) / cin ** 0.5).to(dev)
        bia = (0.1 * torch.randn(cout, generator=gw)).to(dev)
        blk = CorrBlock(dev_f[0][0], dev_f[0][1], num_levels=L, radius=r, mode=args.mode)
        packed = PackedConvC1(wgt, bia, L, r)

        def timed(fn, n=16):
            for i in range(3):
                fn(dev_c[i % iters])
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for i in range(n):
                fn(dev_c[i % iters])
            a1.record()
            torch.cuda.synchronize(dev)
            return 1e3 * a0.elapsed_time(a1) / n

        t_f = timed(lambda c: blk.lookup_conv(c, packed))
        t_p = timed(lambda c: F.relu(F.conv2d(blk(c), wgt, bia)))
        fused = {"kernel": "lookup_conv_kernel", "out_channels": cout, "us_per_launch": t_f,
                 "unfused_pair_us": t_p,
                 "note": "corr_fn(coords) + relu(convc1(corr)) (core/raft.py:219, core/update.py:202) in one launch, "
                         "fp16 tensor-core operands; unfused pair = this package's lookup + torch conv2d (cuDNN, TF32 "
                         "allowed as in the reference's defaults) + relu"}
        del blk, packed

    # ---- end to end: host buffers in, host result out ------------------------------------------
    run_e2e(max(2, args.warmup))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    run_e2e(args.steps)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None  # sampled from warm-up to here: the GPU is under load throughout

    total_ms, e2e_ms = parallel.max_over_ranks([total_ms, e2e_ms], device=dev)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    pairs = world * B * args.steps
    value = pairs / (total_ms * 1e-3)
    hbm_peak, tc_peak, peak_kind = measured_peaks()
    build_bytes, lookup_bytes, flops = algorithmic_bytes(B, C, H, W, r, L)
    lookup_gbs = lookup_bytes / (lookup_ms * 1e-3) / 1e9
    build_gbs = build_bytes / (build_ms * 1e-3) / 1e9
    build_tflops = flops / (build_ms * 1e-3) / 1e12
    launches_build = {"fp32": L, "bf16x3": 2, "bf16": 2}[args.mode]
    line = {
        "metric": "corr pairs/sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}: {desc}", "C": C, "grid": [H, W], "radius": r, "levels": L,
                   "iters": iters, "pairs_per_gpu": B, "build_mode": args.mode, "pyramid_dtype": args.pyramid,
                   "l2": "inputs larger than L2: every step streams a "
                         f"{build_bytes / 1e9:.2f} GB pyramid through the 126 MB L2 and alternates input sets",
                   "parallelism": f"batch shards, {world} x 1 process per GPU, no collective"},
        "e2e": {"value": pairs / (e2e_ms * 1e-3), "unit": "pairs/s",
                "h2d_bytes_per_step": host_f[0].numel() * 4 + host_c.numel() * 4,
                "d2h_bytes_per_step": host_out.numel() * 4,
                "note": "pinned host fmaps+coords -> CorrBlock(...)(coords) x iters -> last corr tensor to pinned host; "
                        "copies of neighbouring steps overlap the kernels on separate streams"},
        "gpu_launches": args.steps * (launches_build + iters),
        "roofline": {"kernel": "lookup_tma_kernel<4>", "bound": "hbm", "achieved": lookup_gbs, "peak": hbm_peak,
                     "unit": "GB/s", "frac": lookup_gbs / hbm_peak, "traffic": ncu_traffic("lookup"),
                     "peak_source": peak_kind, "us_per_launch": lookup_ms * 1e3,
                     "algorithmic_bytes_per_launch": lookup_bytes},
        "roofline_build": {"kernel": f"pack_operands_kernel + build_tc_kernel<1> [{args.mode}]", "bound": "hbm", "achieved": build_gbs,
                           "peak": hbm_peak, "unit": "GB/s", "frac": build_gbs / hbm_peak,
                           "tensor_tflops": build_tflops, "tensor_frac_of_bf16_sustained": build_tflops / tc_peak,
                           "traffic": ncu_traffic("build"), "us_per_launch": build_ms * 1e3,
                           "algorithmic_bytes_per_launch": build_bytes, "algorithmic_flops": flops},
        "clocks": clocks,
    }
    if fast:
        line["fast_mode"] = fast
    if fused:
        line["fused_lookup_convc1"] = fused
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as orc
        orc.set_num_threads(len(os.sched_getaffinity(0)))
        cores = orc.num_threads()
        cpu_reference_step(1, C, H, W, r, L, 2, SEED)  # warm the OpenMP pool / page in the library
        dt, tb, tl = cpu_reference_step(B, C, H, W, r, L, iters, SEED)
        line["cpu_baseline"] = {"value": B / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
                                "sample": f"{B} frame pairs (one full step), build {tb:.2f} s + {iters} lookups x {tl * 1e3:.1f} ms, "
                                          f"oracle/corr_oracle.c with OpenMP on {cores} threads"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--mode", default=os.environ.get("RAFT_CORR_MODE", "bf16x3"), choices=["fp32", "bf16x3", "bf16"])
    ap.add_argument("--pyramid", default="f32", choices=["f32", "f16"], help="storage type of the correlation pyramid")
    ap.add_argument("--no-fast-mode", action="store_true", help="skip the extra bf16 + fp16-pyramid measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        return run_reference(args, cfg)
    return run_ours(args, cfg)


if __name__ == "__main__":
    sys.exit(main())
