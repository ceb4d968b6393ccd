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
out = torch.empty((B, L * rd * rd, H, W), device=dev)
lib = _cabi.lib()
s = torch.cuda.current_stream().cuda_stream
def call(c):
    _cabi.check(lib.rcb_corr_lookup_planned(st.plan.ptr, c.data_ptr(), out.data_ptr(), s), "lookup")
for i in range(8):
    call(coords[i])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(a.reps):
    call(coords[i % 8])
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / a.reps * 1e3
_, lb, _ = algorithmic_bytes(B, C, H, W, r, L)
print(f"{a.config} debug={os.environ.get('RCB_LOOKUP_DEBUG', '0')} {'smooth' if a.smooth else f'sigma={a.sigma}'} lookup {us:.1f} us/launch  "
      f"{lb / us / 1e3:.0f} GB/s algorithmic = {lb / us / 1e3 / 6545.3:.3f} of HBM peak")
call(coords[0])
torch.cuda.synchronize()
if a.dump:
    torch.save(out.cpu(), a.dump)
if a.compare:
    ref = torch.load(a.compare)
    d = (out.cpu() - ref).abs().max().item()
    print(f"max abs diff vs {a.compare}: {d:.3e} (ref max {ref.abs().max().item():.3f})")
