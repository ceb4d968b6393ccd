#!/usr/bin/env python
"""Benchmark of the RAFT correlation hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2] [--mode f16f8]
                    [--alternate] [--train]

A step = one pass of the hot path over one batch of frame pairs: build the correlation pyramid from
(fmap1, fmap2), then `iters` window lookups with a fresh coords tensor each (what core/raft.py:186-219 does
per forward).  Default workload = BASELINE.json configs[1]: RAFT-full, Sintel 440x1024 -> 55x128 grid,
C=256, radius 4, 4 levels, batch 8 per GPU, 32 iterations.  Multi-GPU: one process per GPU (torchrun), each
rank owns its own batch shard (weak scaling), no data-path collective; value = all pairs / max-over-ranks time.

    --alternate   the on-the-fly path (AlternateCorrBlock -> alt_cuda_corr, core/corr.py:130-198) instead of the
                  all-pairs volume; BASELINE.json configs[3] (cfg4, 1088x1920) compares the two.
    --train       BASELINE.json configs[4] (cfg5): a whole training step of the unmodified reference RAFT-full with this
                  package patched in -- forward, sequence loss, backward with the gradient all-reduce (NCCL) launched
                  bucket by bucket from backward hooks, clipping, AdamW/OneCycle (train.py:172,197-236).

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import importlib.util
import json
import os
import statistics
import subprocess
import sys
import tarfile
import tempfile
import time
import warnings

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
MIN_TIMED_S = 1.0  # the K-step block is repeated until at least this much device time has been measured
REF_TAR = os.path.join(ROOT, "oracle", "_ref", "reference_raft.tar")
REF_EXT = os.path.join(ROOT, "oracle", "_ref", "alt_cuda_corr.so")


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


def config_dict(name, cfg, path):
    """The workload, identical in both arms (`--impl ours` / `--impl reference`); arm-specific settings (build mode,
    pyramid storage type) are reported outside of it."""
    B, C, H, W, r, L, iters, desc = cfg
    build_bytes = algorithmic_bytes(B, C, H, W, r, L)[0]
    return {"workload": f"{name}: {desc}", "path": path, "C": C, "grid": [H, W], "radius": r, "levels": L,
            "iters": iters, "pairs_per_gpu": B,
            "l2": "inputs larger than L2: every step streams a "
                  f"{build_bytes / 1e9:.2f} GB pyramid through the 126 MB L2 and alternates input sets",
            "parallelism": "batch shards, one process per GPU, no data-path collective"}


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
        sm, mx, pw, reasons = [], [], [], set()
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
            try:
                pw.append(float(parts[2]))
            except ValueError:
                pass
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        self.file.close()
        os.unlink(self.file.name)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "sm_mhz_min": min(sm), "power_w_max": max(pw) if pw else None}


# ---- the reference itself (unmodified core/corr.py, staged by oracle/stage_reference.py; it travels to the GPU box
# as oracle/_ref/reference_raft.tar because /root/reference does not exist there) --------------------------------
_ref_cache = {}


def reference_corr_module():
    """The reference's core/corr.py imported from the staged archive, or None when it was never staged."""
    if "mod" in _ref_cache:
        return _ref_cache["mod"]
    mod = None
    if os.path.exists(REF_TAR):
        d = tempfile.mkdtemp(prefix="rcb_ref_")
        with tarfile.open(REF_TAR) as tar:
            tar.extractall(d)
        core = os.path.join(d, "core")
        sys.path.insert(0, core)
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                spec = importlib.util.spec_from_file_location("rcb_reference_corr", os.path.join(core, "corr.py"))
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)  # `import alt_cuda_corr` inside is optional there (try/except)
        except Exception:  # noqa: BLE001 -- a reference that does not import is reported as unavailable
            mod = None
        _ref_cache["dir"] = d
    _ref_cache["mod"] = mod
    return mod


def synthetic_numpy(B, C, H, W, iters, seed):
    import numpy as np
    rs = np.random.RandomState(seed)
    f1 = (0.75 * rs.standard_normal((B, C, H, W))).astype(np.float32)
    f2 = (0.75 * rs.standard_normal((B, C, H, W))).astype(np.float32)
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    grid = np.stack([xs, ys])[None].astype(np.float32)
    coords = [(grid + 4.0 * rs.standard_normal((B, 2, H, W))).astype(np.float32) for _ in range(iters)]
    return f1, f2, coords


def cpu_port_step(B, C, H, W, r, L, iters, seed):
    """One bounded sample of the workload on the host cores with the CPU oracle port (oracle/corr_oracle.c).
    Returns seconds for: volume (fp32 accumulate, like the reference's SGEMM) + pooled pyramid + `iters` lookups."""
    from oracle import oracle as orc
    f1, f2, coords = synthetic_numpy(B, C, H, W, iters, seed)
    t0 = time.perf_counter()
    blk = orc.OracleCorrBlock(f1, f2, num_levels=L, radius=r, acc64=False)
    t1 = time.perf_counter()
    for c in coords:
        blk(c)
    t2 = time.perf_counter()
    return t2 - t0, t1 - t0, (t2 - t1) / max(iters, 1)


def cpu_reference_step(ref_corr, B, C, H, W, r, L, iters, seed):
    """The same sample through the UNMODIFIED reference CorrBlock (core/corr.py:12-127: torch.matmul, avg_pool2d,
    grid_sample) on the host cores."""
    import torch
    f1, f2, coords = synthetic_numpy(B, C, H, W, iters, seed)
    f1, f2 = torch.from_numpy(f1), torch.from_numpy(f2)
    coords = [torch.from_numpy(c) for c in coords]
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t0 = time.perf_counter()
        blk = ref_corr.CorrBlock(f1, f2, num_levels=L, radius=r)
        t1 = time.perf_counter()
        for c in coords:
            blk(c)
        t2 = time.perf_counter()
    return t2 - t0, t1 - t0, (t2 - t1) / max(iters, 1)


def cpu_sample_pairs(B, H, W):
    """Pairs per CPU step: the whole batch for cfg2-sized work, fewer for the larger grids (work ~ pairs * Q^2)."""
    budget = 8 * (55 * 128) ** 2
    return max(1, min(B, budget // (H * W) ** 2))


def host_threads():
    return len(os.sched_getaffinity(0))


def run_reference(args, cfg):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores -- the unmodified
    core/corr.py CorrBlock through torch's CPU ops with every host thread (kind "reference"); the C port of oracle/ is
    timed beside it (and stands in, kind "port", only if the staged reference archive is missing)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    B, C, H, W, r, L, iters, desc = cfg
    cores = host_threads()
    pairs = cpu_sample_pairs(B, H, W)
    from oracle import oracle as orc
    orc.set_num_threads(cores)  # torchrun exports OMP_NUM_THREADS=1; this arm is meant to use every host core
    ref_corr = reference_corr_module()
    if ref_corr is not None:
        import torch
        torch.set_num_threads(cores)
        kind, what = "reference", "unmodified reference core/corr.py CorrBlock (torch CPU ops)"

        def step(k, n_iters):
            return cpu_reference_step(ref_corr, pairs, C, H, W, r, L, n_iters, SEED + k)
    else:
        kind, what = "port", "oracle/corr_oracle.c with OpenMP (reference archive not staged)"

        def step(k, n_iters):
            return cpu_port_step(pairs, C, H, W, r, L, n_iters, SEED + k)
    for k in range(args.warmup):
        step(k, max(1, iters // 8))
    times, tb, tl = [], 0.0, 0.0
    for k in range(args.steps):
        dt, b_, l_ = step(k, iters)
        times.append(dt)
        tb += b_
        tl += l_
    t = sum(times)
    value = pairs * args.steps / t
    sample = (f"{pairs} frame pairs per step, build {tb / args.steps:.2f} s + {iters} lookups x "
              f"{1e3 * tl / args.steps:.1f} ms, {what} on {cores} threads")
    line = {
        "impl": "reference", "metric": "corr pairs/sec", "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args.config, cfg, "alternate" if args.alternate else "all-pairs"),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "timing": {"ms_per_step_median": 1e3 * statistics.median(times), "ms_per_step_min": 1e3 * min(times)},
    }
    if kind == "reference":  # the C port beside it, one step
        cpu_port_step(1, C, H, W, r, L, 2, SEED)
        dt, pb, pl = cpu_port_step(pairs, C, H, W, r, L, iters, SEED)
        line["cpu_port"] = {"value": pairs / dt, "unit": "pairs/s", "cores": orc.num_threads(), "kind": "port",
                            "sample": f"{pairs} frame pairs, build {pb:.2f} s + {iters} lookups x {pl * 1e3:.1f} ms, "
                                      "oracle/corr_oracle.c with OpenMP"}
    print(json.dumps(line))
    return 0


def timed_blocks(step_fn, steps, barrier, max_over_ranks, n_events=3):
    """Runs blocks of `steps` steps -- each block bracketed by barrier + synchronize, every step with its own CUDA
    events -- until MIN_TIMED_S of device time has been measured.  Returns per-step records (lists of event tuples
    resolved to milliseconds) and the number of blocks; the block count is agreed on by all ranks."""
    import torch
    per_step, blocks, total = [], 0, 0.0
    k = 0
    while True:
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(n_events)] for _ in range(steps)]
        end = torch.cuda.Event(enable_timing=True)
        barrier()
        for i in range(steps):
            step_fn(k, evs[i])
            k += 1
        end.record()
        barrier()
        for i in range(steps):
            nxt = evs[i + 1][0] if i + 1 < steps else end
            per_step.append([evs[i][0].elapsed_time(nxt)] +
                            [evs[i][j].elapsed_time(evs[i][j + 1]) for j in range(n_events - 1)])
        total += evs[0][0].elapsed_time(end)
        blocks += 1
        # all ranks take the same decision (the slowest rank's clock)
        if max_over_ranks([total])[0] * 1e-3 >= MIN_TIMED_S or blocks >= 200:
            return per_step, blocks


def run_ours(args, cfg):
    import torch
    import torch.distributed as dist
    from raft_optical_flow_b200 import AlternateCorrBlock, CorrBlock, _cabi, parallel

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: raft_optical_flow_b200 has no CPU fallback")
    B, C, H, W, r, L, iters, desc = cfg
    rank, world, local = parallel.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _cabi.lib()  # fail loudly if the extension is missing
    parallel.bind_to_gpu_cpus(local)  # pinned buffers and the launching thread next to this rank's GPU

    g = torch.Generator(device="cpu").manual_seed(SEED + rank)
    nsets = 2  # alternate input sets; the 2.2 GB pyramid written every step flushes the 126 MB L2 anyway
    host_f = [(0.75 * torch.randn(2, B, C, H, W, generator=g)).pin_memory() for _ in range(nsets)]
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    grid = torch.stack([xs, ys]).float()[None]
    host_c = (grid + 4.0 * torch.randn(iters, B, 2, H, W, generator=g)).pin_memory()
    dev_f = [h.to(dev) for h in host_f]
    dev_c = host_c.to(dev)
    out_ch = L * (2 * r + 1) ** 2
    host_out = torch.empty((B, out_ch, H, W), dtype=torch.float32).pin_memory()
    host_us = {"n": 0, "t": 0.0}

    def make_block(f, mode=None, pdt=None):
        if args.alternate:
            return AlternateCorrBlock(f[0], f[1], num_levels=L, radius=r)
        return CorrBlock(f[0], f[1], num_levels=L, radius=r, mode=mode or args.mode, pyramid_dtype=pdt or args.pyramid)

    def step_resident(k, ev=None, mode=None, pdt=None):
        f = dev_f[k % nsets]
        if ev:
            ev[0].record()
        blk = make_block(f, mode, pdt)
        if ev:
            ev[1].record()
        out = None
        t0 = time.perf_counter()
        for i in range(iters):
            out = blk(dev_c[i])
        host_us["t"] += time.perf_counter() - t0
        host_us["n"] += iters
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
    out_bufs = [torch.empty((B, out_ch, H, W), dtype=torch.float32, device=dev) for _ in range(2)]
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
        blk = make_block(f)
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

    def mx(vals):
        return parallel.max_over_ranks(vals, device=dev)

    # ---- device-resident timing -------------------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    t_load = time.perf_counter()
    for k in range(args.warmup):
        step_resident(k)
    while time.perf_counter() - t_load < 0.6:  # nvidia-smi needs a few hundred ms to deliver its first samples
        step_resident(0)
    host_us["n"], host_us["t"] = 0, 0.0
    per_step, blocks = timed_blocks(lambda k, ev: step_resident(k, ev), args.steps, barrier, mx)
    step_ms = sorted(p[0] for p in per_step)
    med_ms, min_ms = statistics.median(step_ms), step_ms[0]
    build_ms = statistics.median(p[1] for p in per_step)
    lookup_ms = statistics.median(p[2] for p in per_step) / iters
    host_lookup_us = 1e6 * host_us["t"] / max(host_us["n"], 1)
    # clocks of the warm-up + timed region of `value` only: the GPU is under this load throughout (the legs below --
    # PCIe-bound e2e, CPU baselines -- leave it partly idle and would pull the median back to the maximum clock)
    clocks = sampler.stop() if sampler else None

    # ---- fast mode, reported next to the headline (never instead of it): single bf16 pass + fp16-stored pyramid,
    # the "within a stated bound" path of the spec (flow EPE delta <= 0.01 px mean, tests/test_gpu_e2e_raft.py)
    fast = None
    if not args.alternate and args.pyramid == "f32" and args.mode != "bf16" and not args.no_extras:
        for k in range(3):
            step_resident(k, mode="bf16", pdt="f16")
        fevs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(min(args.steps, 5))]
        fend = torch.cuda.Event(enable_timing=True)
        barrier()
        for k in range(len(fevs)):
            step_resident(k, fevs[k], mode="bf16", pdt="f16")
        fend.record()
        barrier()
        f_total = mx([fevs[0][0].elapsed_time(fend)])[0]
        fast = {"build_mode": "bf16", "pyramid_dtype": "f16", "value": world * B * len(fevs) / (f_total * 1e-3),
                "unit": "pairs/s", "ms_per_step": f_total / len(fevs),
                "build_us": 1e3 * sum(e[0].elapsed_time(e[1]) for e in fevs) / len(fevs),
                "lookup_us": 1e3 * sum(e[1].elapsed_time(e[2]) for e in fevs) / (len(fevs) * iters),
                "note": "operands rounded to bf16 (2.6e-3 of max-abs), pyramid stored as fp16; "
                        "final-flow EPE delta vs the reference 1e-3 px mean on the demo frames"}

    # ---- next row of the scope table (8f, f1), reported next to the headline: the lookup fused with the motion
    # encoder's first layer, against this package's lookup followed by the convolution + ReLU the reference runs
    fused = None
    if not args.alternate and args.pyramid == "f32" and r in (3, 4) and not args.no_extras:
        import torch.nn.functional as F
        from raft_optical_flow_b200 import PackedConvC1
        cout, cin = (256 if r == 4 else 96), out_ch
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

    # ---- batch-1 launch floor: the same step replayed as ONE CUDA graph (build + all lookups), the form
    # patch_raft(cuda_graph=True) uses around the GRU loop; reported for every config, decisive for cfg1
    graph = None
    if not args.alternate and not args.no_extras:
        graph = graph_step_timing(torch, CorrBlock, dev, dev_f[0], dev_c, L, r, iters, args, mx, world, B)

    # ---- the other path of cfg4 ("alternate_corr on-the-fly path vs all-pairs volume"), a few steps
    other_path = None
    if args.alternate and not args.no_extras:
        saved = args.alternate
        args.alternate = False
        try:
            for k in range(2):
                step_resident(k)
            oevs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(3)]
            oend = torch.cuda.Event(enable_timing=True)
            barrier()
            for k in range(3):
                step_resident(k, oevs[k])
            oend.record()
            barrier()
            o_total = mx([oevs[0][0].elapsed_time(oend)])[0]
            other_path = {"path": "all-pairs", "build_mode": args.mode, "value": world * B * 3 / (o_total * 1e-3),
                          "unit": "pairs/s", "ms_per_step": o_total / 3,
                          "build_us": 1e3 * sum(e[0].elapsed_time(e[1]) for e in oevs) / 3,
                          "lookup_us": 1e3 * sum(e[1].elapsed_time(e[2]) for e in oevs) / (3 * iters),
                          "pyramid_gb": algorithmic_bytes(B, C, H, W, r, L)[0] / 1e9}
        finally:
            args.alternate = saved
        torch.cuda.empty_cache()

    # ---- the reference's own code on this GPU, same inputs (SURVEY 8d "beat this on the same box") -----------
    gpu_ref = None
    if rank == 0 and not args.no_extras:
        gpu_ref = gpu_reference_timings(torch, dev, dev_f[0], dev_c, L, r, iters, B)

    # ---- end to end: host buffers in, host result out ------------------------------------------
    e2e_sampler = ClockSampler(local) if rank == 0 else None
    if e2e_sampler:
        e2e_sampler.start()
    run_e2e(max(2, args.warmup))
    e2e_steps = args.steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    run_e2e(e2e_steps)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    if mx([e2e_ms])[0] < 500.0:  # short region: measure a longer one instead (pipeline fill/drain amortised)
        e2e_steps = args.steps * max(2, int(1000.0 / max(e2e_ms, 1.0)))
        barrier()
        e0.record()
        run_e2e(e2e_steps)
        e1.record()
        barrier()
        e2e_ms = e0.elapsed_time(e1)
    e2e_clocks = e2e_sampler.stop() if e2e_sampler else None

    # ---- the same step after an idle second: the timed region above runs at the board's power cap (sw_power_cap, SM
    # clock ~1.6 of 1.965 GHz, tools/power_timeline.py), where both kernels take 5-12 % longer than they do alone
    burst = None
    if not args.alternate and not args.no_extras:
        barrier()
        time.sleep(1.0)
        bev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(3)]
        for k in range(3):
            step_resident(k, bev[k])
        barrier()
        b_build = mx([min(e[0].elapsed_time(e[1]) for e in bev)])[0]
        b_lookup = mx([min(e[1].elapsed_time(e[2]) for e in bev)])[0] / iters
        burst = {"build_us": b_build * 1e3, "lookup_us": b_lookup * 1e3,
                 "note": "3 steps after 1 s idle, minimum: the kernels before the power cap pulls the SM clock down"}

    # ---- the same boundary one level up: frames in, flow out through the unmodified reference model ------
    frames_e2e = None
    if not args.alternate and not args.no_extras:
        del in_bufs, out_bufs
        torch.cuda.empty_cache()
        frames_e2e = e2e_frames_timing(torch, dev, cfg, world, mx, barrier, args.steps)

    # every rank's own figures (GPUs of one box differ by a few percent; the headline is the slowest rank)
    per_rank = None
    if world > 1:
        mine = torch.tensor([med_ms, build_ms * 1e3, lookup_ms * 1e3, host_lookup_us], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [{"rank": i, "ms_per_step": round(float(t[0]), 4), "build_us": round(float(t[1]), 1),
                     "lookup_us": round(float(t[2]), 2), "host_us_per_lookup_call": round(float(t[3]), 1)}
                    for i, t in enumerate(allr)]
    med_ms, min_ms, e2e_ms = mx([med_ms, min_ms, e2e_ms])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    value = world * B / (med_ms * 1e-3)
    hbm_peak, tc_peak, peak_kind = measured_peaks()
    build_bytes, lookup_bytes, flops = algorithmic_bytes(B, C, H, W, r, L)
    h2d = host_f[0].numel() * 4 + host_c.numel() * 4
    d2h = host_out.numel() * 4
    e2e_s = e2e_ms * 1e-3
    line = {
        "metric": "corr pairs/sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": med_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args.config, cfg, "alternate" if args.alternate else "all-pairs"),
        "impl_config": {"build_mode": None if args.alternate else args.mode,
                        "pyramid_dtype": None if args.alternate else args.pyramid},
        "timing": {"ms_per_step_median": med_ms, "ms_per_step_min": min_ms, "timed_steps": len(per_step),
                   "blocks_of_steps": blocks, "build_us_median": build_ms * 1e3, "lookup_us_median": lookup_ms * 1e3,
                   "host_us_per_lookup_call": host_lookup_us,
                   "note": f"blocks of {args.steps} steps repeated until >= {MIN_TIMED_S} s of device time; value = "
                           "pairs / median step time (max over ranks)"},
        "e2e": {"value": world * B * e2e_steps / e2e_s, "unit": "pairs/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "h2d_gbs_per_rank": h2d * e2e_steps / e2e_s / 1e9, "d2h_gbs_per_rank": d2h * e2e_steps / e2e_s / 1e9,
                "cpu_binding": parallel.binding_description(), "clocks": e2e_clocks,
                "note": "pinned host fmaps+coords -> CorrBlock(...)(coords) x iters -> last corr tensor to pinned host; "
                        "copies of neighbouring steps overlap the kernels on separate streams"},
        "clocks": clocks,
    }
    if args.alternate:
        alt_flops = 2.0 * B * H * W * L * (2 * r + 2) ** 2 * C  # B*Q*L*(2r+2)^2*2C per call
        fma_peak = 2.0 * 128 * 148 * (clocks["sm_max_mhz"] if clocks else 1965.0) * 1e6 / 1e12
        tf = alt_flops / (lookup_ms * 1e-3) / 1e12
        line["gpu_launches"] = len(per_step) * (iters + 3)
        line["roofline"] = {"kernel": "altcorr_fwd_kernel (all levels, one launch per call)", "bound": "fp32_fma",
                            "achieved": tf, "peak": fma_peak, "unit": "TFLOP/s", "frac": tf / fma_peak, "traffic": None,
                            "peak_source": "148 SMs x 128 FMA/clk x 2 x max SM clock (no tensor-core form: windows are "
                                           "per query and data dependent)",
                            "us_per_launch": lookup_ms * 1e3, "algorithmic_flops_per_launch": alt_flops,
                            "prepare_us": build_ms * 1e3}
        if other_path:
            line["all_pairs_path"] = other_path
    else:
        lookup_gbs = lookup_bytes / (lookup_ms * 1e-3) / 1e9
        build_gbs = build_bytes / (build_ms * 1e-3) / 1e9
        build_tflops = flops / (build_ms * 1e-3) / 1e12
        launches_build = {"fp32": L, "bf16x3": 2, "bf16": 2, "f16f8": 2}[args.mode]
        executed = {"fp32": 1.0, "bf16x3": 3.0, "bf16": 1.0, "f16f8": 2.0}[args.mode]
        line["gpu_launches"] = len(per_step) * (launches_build + iters)
        line["roofline"] = {"kernel": "lookup_tma_kernel<4>", "bound": "hbm", "achieved": lookup_gbs, "peak": hbm_peak,
                            "unit": "GB/s", "frac": lookup_gbs / hbm_peak, "traffic": ncu_traffic("lookup"),
                            "peak_source": peak_kind, "us_per_launch": lookup_ms * 1e3,
                            "algorithmic_bytes_per_launch": lookup_bytes}
        rd = ncu_traffic("lookup_dram_read")
        if rd and args.config == "cfg2" and args.pyramid == "f32":
            # back to back, every launch reads what ncu saw it read and (sooner or later) writes its whole output
            steady = rd + 4 * B * out_ch * H * W
            line["roofline"]["steady_state"] = {
                "traffic": steady, "gbs": steady / (lookup_ms * 1e-3) / 1e9,
                "frac": steady / (lookup_ms * 1e-3) / 1e9 / hbm_peak,
                "note": "DRAM bytes per launch in a loop of launches = ncu dram read of one launch + the whole output "
                        "(a single-launch capture defers part of the write-back past the launch); the distance to "
                        "`frac` is the 64-byte atom over-fetch of the 10x10 windows (10.0 atoms instead of 6.25)"}
        if burst:
            line["roofline"]["burst"] = {"us_per_launch": burst["lookup_us"], "frac": lookup_bytes / burst["lookup_us"] / 1e3 / hbm_peak}
        line["roofline_build"] = {"kernel": f"pack + build_tc_kernel [{args.mode}]", "bound": "hbm",
                                  "achieved": build_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": build_gbs / hbm_peak,
                                  "tensor_tflops": build_tflops, "tensor_frac_of_bf16_sustained": build_tflops / tc_peak,
                                  "tensor_pass_equivalents": executed,
                                  "tensor_frac_executed": executed * build_tflops / tc_peak,
                                  "traffic": ncu_traffic("build"), "us_per_launch": build_ms * 1e3,
                                  "algorithmic_bytes_per_launch": build_bytes, "algorithmic_flops": flops}
        if burst:
            line["roofline_build"]["burst"] = {"us_per_launch": burst["build_us"],
                                               "frac": build_bytes / burst["build_us"] / 1e3 / hbm_peak}
            line["timing"]["burst"] = burst
    if per_rank:
        line["per_rank"] = per_rank
    if fast:
        line["fast_mode"] = fast
    if fused:
        line["fused_lookup_convc1"] = fused
    if graph:
        line["cuda_graph_step"] = graph
    if gpu_ref:
        line["gpu_reference"] = gpu_ref
    if frames_e2e:
        line["e2e_frames"] = frames_e2e
    if world == 1 and not args.no_cpu_baseline:
        line.update(cpu_baseline_legs(cfg))
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def graph_step_timing(torch, CorrBlock, dev, f, dev_c, L, r, iters, args, mx, world, B):
    """build + `iters` lookups captured once into a CUDA graph and replayed: no per-launch host work at all."""
    try:
        side = torch.cuda.Stream(dev)
        outs = []
        with torch.cuda.stream(side):
            for _ in range(2):  # warm the allocator on the capture stream
                blk = CorrBlock(f[0], f[1], num_levels=L, radius=r, mode=args.mode, pyramid_dtype=args.pyramid)
                out = blk(dev_c[0])
        side.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            blk = CorrBlock(f[0], f[1], num_levels=L, radius=r, mode=args.mode, pyramid_dtype=args.pyramid)
            for i in range(iters):
                outs.append(blk(dev_c[i]))
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize(dev)
        # timed like `value`: blocks of replays until >= MIN_TIMED_S, median block -- a quarter of a second would end
        # before the power cap has pulled the clock down and flatter the graph by 3-4 %
        n, blocks, total = 20, [], 0.0
        while total < MIN_TIMED_S * 1e3 and len(blocks) < 200:
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(n):
                g.replay()
            a1.record()
            torch.cuda.synchronize(dev)
            blocks.append(a0.elapsed_time(a1) / n)
            total = mx([total + a0.elapsed_time(a1)])[0]
        ms = mx([statistics.median(blocks)])[0]
        del g, outs, blk, out
        torch.cuda.empty_cache()
        return {"ms_per_step": ms, "value": world * B / (ms * 1e-3), "unit": "pairs/s", "replays": n * len(blocks),
                "note": "one graph = pack + build + all lookups of a step (every lookup keeps its own output tensor); "
                        "median of blocks of 20 replays over >= 1 s, as for `value`"}
    except Exception as e:  # noqa: BLE001 -- an optional figure must not take the headline down
        return {"error": f"{type(e).__name__}: {e}"[:200]}


def gpu_reference_timings(torch, dev, f, dev_c, L, r, iters, B):
    """The reference's own implementation on the same B200 and the same inputs: CorrBlock through torch's CUDA ops
    (core/corr.py:25-94: cuBLAS SGEMM, avg_pool2d, grid_sample) and AlternateCorrBlock through the reference's own
    alt_cuda_corr kernels compiled for sm_100 (oracle/_ref/alt_cuda_corr.so)."""
    ref_corr = reference_corr_module()
    if ref_corr is None:
        return {"unavailable": "oracle/_ref/reference_raft.tar not staged"}
    res = {}
    try:
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)

            def timed(fn, reps):
                fn()
                torch.cuda.synchronize(dev)
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                for _ in range(reps):
                    fn()
                a1.record()
                torch.cuda.synchronize(dev)
                return a0.elapsed_time(a1) / reps

            state = {}

            def build():
                state["blk"] = ref_corr.CorrBlock(f[0], f[1], num_levels=L, radius=r)

            t_build = timed(build, 3)
            t_look = timed(lambda: [state["blk"](dev_c[i]) for i in range(min(iters, 8))], 2) / min(iters, 8)
            state.clear()
            torch.cuda.empty_cache()
            step_ms = t_build + iters * t_look
            res["corrblock_torch_ops"] = {"build_us": 1e3 * t_build, "lookup_us": 1e3 * t_look, "ms_per_step": step_ms,
                                          "value": B / (step_ms * 1e-3), "unit": "pairs/s",
                                          "matmul_tf32": bool(tf32[0])}
            if os.path.exists(REF_EXT):
                spec = importlib.util.spec_from_file_location("alt_cuda_corr", REF_EXT)
                ext = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(ext)
                ref_corr.alt_cuda_corr = ext  # what `import alt_cuda_corr` (core/corr.py:6) would have bound
                alt = ref_corr.AlternateCorrBlock(f[0], f[1], num_levels=L, radius=r)
                t_call = timed(lambda: alt(dev_c[0]), 2)
                res["alternate_reference_extension"] = {"us_per_call": 1e3 * t_call,
                                                        "value": B / (iters * t_call * 1e-3), "unit": "pairs/s"}
                del alt
            else:
                res["alternate_reference_extension"] = {"unavailable": "oracle/_ref/alt_cuda_corr.so not built"}
    except Exception as e:  # noqa: BLE001
        res["error"] = f"{type(e).__name__}: {e}"[:200]
    torch.cuda.empty_cache()
    return res


def e2e_frames_timing(torch, dev, cfg, world, mx, barrier, steps):
    """The faithful drop-in boundary (demo.py:59-65, core/raft.py:164-182): uint8 frames in pinned host memory ->
    H2D -> the UNMODIFIED reference RAFT (random init) with this package patched in (patch_raft) -> full-resolution
    flow -> D2H, every step; and the same model with the reference's own CorrBlock / upsampling on the same GPU."""
    import argparse as ap
    from raft_optical_flow_b200 import patch_raft, train_bench
    B, C, H, W, r, L, iters, desc = cfg
    try:
        raft_mod = train_bench.load_reference_raft(ROOT)
        if raft_mod is None:
            return {"unavailable": "oracle/_ref/reference_raft.tar not staged"}
        small = C == 128
        torch.manual_seed(SEED)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            model = raft_mod.RAFT(ap.Namespace(small=small, mixed_precision=False, alternate_corr=False, dropout=0.0))
        model = model.to(dev).eval()
        g = torch.Generator(device="cpu").manual_seed(SEED)
        frames = torch.randint(0, 256, (2, B, 3, 8 * H, 8 * W), dtype=torch.uint8, generator=g).pin_memory()
        host_flow = torch.empty((B, 2, 8 * H, 8 * W), dtype=torch.float32).pin_memory()

        def step():
            d = frames.to(dev, non_blocking=True).float()
            with torch.no_grad(), warnings.catch_warnings():
                warnings.simplefilter("ignore")
                _, up = model(d[0], d[1], iters=iters, test_mode=True)
            host_flow.copy_(up, non_blocking=True)

        def timed(n):
            for _ in range(2):
                step()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            a0.record()
            for _ in range(n):
                step()
            a1.record()
            barrier()
            return mx([a0.elapsed_time(a1) / n])[0]

        n = max(3, min(steps, 8))
        ms_ref = timed(n)
        old = patch_raft(raft_mod, fuse_motion_encoder=True)
        try:
            ms_ours = timed(n)
        finally:
            raft_mod._rcb_undo_fused()
            raft_mod.CorrBlock, raft_mod.AlternateCorrBlock, raft_mod.RAFT.upsample_flow = old
        # the same with the whole no-grad forward replayed as one CUDA graph (decisive at batch 1: launch latency)
        ms_graph = None
        try:
            old = patch_raft(raft_mod, fuse_motion_encoder=True, cuda_graph=True)
            try:
                ms_graph = timed(n)
            finally:
                raft_mod._rcb_undo_graph()
                raft_mod._rcb_undo_fused()
                raft_mod.CorrBlock, raft_mod.AlternateCorrBlock, raft_mod.RAFT.upsample_flow = old
        except Exception as e:  # noqa: BLE001
            ms_graph = f"{type(e).__name__}: {e}"[:160]
        del model
        torch.cuda.empty_cache()
        return {"value": world * B / (ms_ours * 1e-3), "unit": "frame pairs/s", "ms_per_step": ms_ours, "steps": n,
                "h2d_bytes_per_step": frames.numel(), "d2h_bytes_per_step": host_flow.numel() * 4,
                "model": "RAFT-small" if small else "RAFT-full", "iters": iters,
                "cuda_graph": ({"ms_per_step": ms_graph, "value": world * B / (ms_graph * 1e-3)}
                               if isinstance(ms_graph, float) else {"error": ms_graph}),
                "reference_model_same_gpu": {"value": world * B / (ms_ref * 1e-3), "ms_per_step": ms_ref,
                                             "note": "the reference's own CorrBlock (torch ops) and upsample_flow"},
                "note": "uint8 frames (pinned host) -> reference RAFT with patch_raft(fuse_motion_encoder=True) -> "
                        "flow [B,2,8H,8W] -> pinned host; encoders / update block are the reference's own cuDNN code"}
    except Exception as e:  # noqa: BLE001 -- an optional figure must not take the headline down
        return {"error": f"{type(e).__name__}: {e}"[:200]}


def cpu_baseline_legs(cfg):
    """cpu_baseline of the repo arm (N = 1 only): one bounded step of the unmodified reference CorrBlock on the host
    cores (kind "reference"), the C port of oracle/ beside it."""
    B, C, H, W, r, L, iters, desc = cfg
    from oracle import oracle as orc
    cores = host_threads()
    orc.set_num_threads(cores)
    pairs = cpu_sample_pairs(B, H, W)
    out = {}
    cpu_port_step(1, C, H, W, r, L, 2, SEED)  # warm the OpenMP pool / page in the library
    dt, tb, tl = cpu_port_step(pairs, C, H, W, r, L, iters, SEED)
    port = {"value": pairs / dt, "unit": "pairs/s", "cores": orc.num_threads(), "kind": "port",
            "sample": f"{pairs} frame pairs (one step), build {tb:.2f} s + {iters} lookups x {tl * 1e3:.1f} ms, "
                      f"oracle/corr_oracle.c with OpenMP on {orc.num_threads()} threads"}
    ref_corr = reference_corr_module()
    if ref_corr is None:
        out["cpu_baseline"] = port
        return out
    import torch
    torch.set_num_threads(cores)
    cpu_reference_step(ref_corr, 1, C, H, W, r, L, 2, SEED)
    dt, tb, tl = cpu_reference_step(ref_corr, pairs, C, H, W, r, L, iters, SEED)
    out["cpu_baseline"] = {"value": pairs / dt, "unit": "pairs/s", "cores": cores, "kind": "reference",
                           "sample": f"{pairs} frame pairs (one step), build {tb:.2f} s + {iters} lookups x "
                                     f"{tl * 1e3:.1f} ms, unmodified reference core/corr.py CorrBlock (torch CPU ops) on "
                                     f"{cores} threads"}
    out["cpu_port"] = port
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS))
    ap.add_argument("--mode", default=os.environ.get("RAFT_CORR_MODE", "f16f8"),
                    choices=["fp32", "bf16x3", "bf16", "f16f8"])
    ap.add_argument("--pyramid", default="f32", choices=["f32", "f16"], help="storage type of the correlation pyramid")
    ap.add_argument("--alternate", action="store_true", help="time the on-the-fly path (AlternateCorrBlock)")
    ap.add_argument("--train", action="store_true", help="cfg5: whole training step with overlapped gradient all-reduce")
    ap.add_argument("--no-extras", "--no-fast-mode", dest="no_extras", action="store_true",
                    help="headline only: skip the fast-mode / fused / graph / same-GPU reference figures")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.config is None:
        args.config = "cfg5" if args.train else "cfg4" if args.alternate else "cfg2"
    cfg = CONFIGS[args.config]
    if args.train:
        from raft_optical_flow_b200 import train_bench
        return train_bench.run(args, cfg, ROOT, sampler_cls=ClockSampler)
    if args.impl == "reference":
        return run_reference(args, cfg)
    return run_ours(args, cfg)


if __name__ == "__main__":
    sys.exit(main())
