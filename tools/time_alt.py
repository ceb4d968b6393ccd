"""Times one fused on-the-fly call of AlternateCorrBlock (all levels, one launch).  python tools/time_alt.py [--config cfg2]"""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CONFIGS, SEED  # noqa: E402
from raft_optical_flow_b200 import AlternateCorrBlock  # noqa: E402
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg2")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--smooth", action="store_true")
a = ap.parse_args()
B, C, H, W, r, L, iters, _ = CONFIGS[a.config]
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(SEED)
f1 = (0.75 * torch.randn(B, C, H, W, generator=g)).to(dev)
f2 = (0.75 * torch.randn(B, C, H, W, generator=g)).to(dev)
ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
grid = torch.stack([xs, ys]).float()[None]
if a.smooth:
    c = (grid + torch.tensor([3.3, -2.7]).view(1, 2, 1, 1)).expand(B, 2, H, W).contiguous().to(dev)
else:
    c = (grid + 4.0 * torch.randn(B, 2, H, W, generator=g)).to(dev)
alt = AlternateCorrBlock(f1, f2, num_levels=L, radius=r)
alt(c); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    alt(c)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / a.reps * 1e3
fl = B * H * W * L * (2 * r + 2) ** 2 * 2 * C
print(f"{a.config} {'smooth' if a.smooth else 'noisy'}: alt call {us:.0f} us, {fl / us / 1e6:.2f} TFLOP/s fp32")
