"""Times rcb_corr_lookup_planned on a prebuilt pyramid (CUDA events on the launching stream, a different coords
tensor per launch, no allocation inside the loop).  --dump / --compare save and diff the output of one launch
across builds; RCB_LOOKUP_DEBUG=1 / 2 time the gather alone / the resampling + stores alone.
    python tools/time_lookup.py [--config cfg2] [--reps 64] [--smooth] [--dump out.pt | --compare out.pt]"""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CONFIGS, SEED, algorithmic_bytes  # noqa: E402
from raft_optical_flow_b200 import CorrBlock, _cabi  # noqa: E402
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg2")
ap.add_argument("--reps", type=int, default=64)
ap.add_argument("--sigma", type=float, default=4.0)
ap.add_argument("--smooth", action="store_true", help="smooth flow field (neighbouring queries share window phase)")
ap.add_argument("--outs", type=int, default=1, help="rotate over this many output buffers (1: every launch overwrites the same 73 MB)")
ap.add_argument("--via-block", action="store_true", help="call CorrBlock.__call__ (allocates its output) instead of the raw C ABI")
ap.add_argument("--alloc", action="store_true", help="raw C ABI, but torch.empty a fresh output for every launch")
ap.add_argument("--no-grad", action="store_true")
ap.add_argument("--graph", action="store_true", help="replay the launches of the timed loop as one CUDA graph")
ap.add_argument("--dump", default=None)
ap.add_argument("--compare", default=None)
a = ap.parse_args()
B, C, H, W, r, L, iters, _ = CONFIGS[a.config]
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(SEED)
f1 = (0.75 * torch.randn(B, C, H, W, generator=g)).to(dev)
f2 = (0.75 * torch.randn(B, C, H, W, generator=g)).to(dev)
blk = CorrBlock(f1, f2, num_levels=L, radius=r)
ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
grid = torch.stack([xs, ys]).float()[None]
if a.smooth:  # slowly varying flow: a different constant + gentle ramp per set
    ramp = torch.stack([xs.float() / W, ys.float() / H])[None]
    coords = [(grid + torch.tensor([3.3 + i, -2.7 - 0.5 * i]).view(1, 2, 1, 1) + 2.0 * ramp).expand(B, 2, H, W).to(dev).contiguous()
              for i in range(8)]
else:
    coords = [(grid + a.sigma * torch.randn(B, 2, H, W, generator=g)).to(dev).contiguous() for _ in range(8)]
st = blk._state
rd = 2 * r + 1
outs = [torch.empty((B, L * rd * rd, H, W), device=dev) for _ in range(a.outs)]
out = outs[0]
lib = _cabi.lib()
def call(c, o=None):
    if a.via_block:
        return blk(c)
    if a.alloc:
        o = torch.empty((B, L * rd * rd, H, W), device=dev)
        _cabi.check(lib.rcb_corr_lookup_planned(st.plan.ptr, c.data_ptr(), o.data_ptr(), torch.cuda.current_stream().cuda_stream), "lookup")
        return o
    o = out if o is None else o
    _cabi.check(lib.rcb_corr_lookup_planned(st.plan.ptr, c.data_ptr(), o.data_ptr(), torch.cuda.current_stream().cuda_stream), "lookup")
def loop():
    keep = None
    for i in range(a.reps):
        keep = call(coords[i % 8], outs[i % a.outs])
for i in range(8):
    call(coords[i])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if a.graph:
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        loop()
        side.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=side):
            loop()
    gr.replay()
    torch.cuda.synchronize()
    e0.record()
    gr.replay()
    e1.record()
else:
    import time
    torch.set_grad_enabled(not a.no_grad)
    e0.record()
    t0 = time.perf_counter()
    loop()
    host_us = (time.perf_counter() - t0) / a.reps * 1e6
    e1.record()
    print(f"host {host_us:.1f} us per call")
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / a.reps * 1e3
_, lb, _ = algorithmic_bytes(B, C, H, W, r, L)
print(f"{a.config} outs={a.outs} via_block={a.via_block} graph={a.graph} debug={os.environ.get('RCB_LOOKUP_DEBUG', '0')} {'smooth' if a.smooth else f'sigma={a.sigma}'} lookup {us:.1f} us/launch  "
      f"{lb / us / 1e3:.0f} GB/s algorithmic = {lb / us / 1e3 / 6545.3:.3f} of HBM peak")
call(coords[0])
torch.cuda.synchronize()
if a.dump:
    torch.save(out.cpu(), a.dump)
if a.compare:
    ref = torch.load(a.compare)
    d = (out.cpu() - ref).abs().max().item()
    print(f"max abs diff vs {a.compare}: {d:.3e} (ref max {ref.abs().max().item():.3f})")
