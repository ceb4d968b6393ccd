"""Where does the eager bench step lose time against the CUDA-graph replay of the same launches?  Variants of one cfg2
step (CorrBlock constructor + iters lookups), events on the launching stream: [start, after build, after lookups].
    python tools/step_anatomy.py [--steps 40]"""
import argparse, os, statistics, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CONFIGS, SEED  # noqa: E402
from raft_optical_flow_b200 import CorrBlock, _cabi  # noqa: E402
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg2")
ap.add_argument("--steps", type=int, default=40)
a = ap.parse_args()
B, C, H, W, r, L, iters, _ = CONFIGS[a.config]
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(SEED)
fs = [(0.75 * torch.randn(2, B, C, H, W, generator=g)).to(dev) for _ in range(2)]
ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
grid = torch.stack([xs, ys]).float()[None]
dev_c = (grid + 4.0 * torch.randn(iters, B, 2, H, W, generator=g)).to(dev)
rd = 2 * r + 1
lib = _cabi.lib()
outs = [torch.empty((B, L * rd * rd, H, W), device=dev) for _ in range(iters)]

def step_api(k, ev):
    f = fs[k % 2]
    ev[0].record()
    blk = CorrBlock(f[0], f[1], num_levels=L, radius=r)
    ev[1].record()
    out = None
    for i in range(iters):
        out = blk(dev_c[i])
    ev[2].record()
    return out

def step_raw(k, ev, nouts=iters):
    f = fs[k % 2]
    ev[0].record()
    blk = CorrBlock(f[0], f[1], num_levels=L, radius=r)
    ev[1].record()
    s = torch.cuda.current_stream().cuda_stream
    p = blk._state.plan.ptr
    for i in range(iters):
        lib.rcb_corr_lookup_planned(p, dev_c[i].data_ptr(), outs[i % nouts].data_ptr(), s)
    ev[2].record()
    return blk

def step_keep(k, ev):  # API, but every output of the step stays alive until the step ends
    f = fs[k % 2]
    ev[0].record()
    blk = CorrBlock(f[0], f[1], num_levels=L, radius=r)
    ev[1].record()
    keep = [blk(dev_c[i]) for i in range(iters)]
    ev[2].record()
    return keep

def step_noev_between(k, ev):  # API, no event between the build and the lookups
    f = fs[k % 2]
    ev[0].record()
    blk = CorrBlock(f[0], f[1], num_levels=L, radius=r)
    out = None
    for i in range(iters):
        out = blk(dev_c[i])
    ev[1].record()
    ev[2].record()
    return out

def run(name, fn):
    for k in range(4):
        fn(k, [torch.cuda.Event(enable_timing=True) for _ in range(3)])
    torch.cuda.synchronize()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(a.steps)]
    end = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    for k in range(a.steps):
        fn(k, evs[k])
    host = (time.perf_counter() - t0) / a.steps * 1e3
    end.record()
    torch.cuda.synchronize()
    tot = [evs[k][0].elapsed_time(evs[k + 1][0] if k + 1 < a.steps else end) for k in range(a.steps)]
    bld = [e[0].elapsed_time(e[1]) for e in evs]
    lk = [e[1].elapsed_time(e[2]) for e in evs]
    print(f"{name:28s} step {statistics.median(tot) * 1e3:7.1f} us  build {statistics.median(bld) * 1e3:6.1f}  "
          f"lookups {statistics.median(lk) * 1e3:7.1f} ({statistics.median(lk) * 1e3 / iters:.2f} each)  host {host * 1e3:6.0f} us/step")

with torch.no_grad():
    run("api", step_api)
    run("raw C ABI, own outputs", step_raw)
    run("raw C ABI, 2 outputs", lambda k, ev: step_raw(k, ev, 2))
    run("api, outputs kept", step_keep)
    run("api, no event after build", step_noev_between)
    run("api", step_api)
