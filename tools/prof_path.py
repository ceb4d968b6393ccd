"""Small driver for ncu captures: one build + a few lookups (and optionally the alternate path) at a bench shape.
    python tools/prof_path.py [--config cfg2] [--mode fp32] [--lookups 4] [--alt]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CONFIGS, SEED  # noqa: E402
from raft_optical_flow_b200 import AlternateCorrBlock, CorrBlock  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg2")
ap.add_argument("--mode", default="f16f8")
ap.add_argument("--lookups", type=int, default=4)
ap.add_argument("--builds", type=int, default=2)
ap.add_argument("--alt", action="store_true")
a = ap.parse_args()
B, C, H, W, r, L, iters, _ = CONFIGS[a.config]
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(SEED)
f1 = (0.75 * torch.randn(B, C, H, W, generator=g)).to(dev)
f2 = (0.75 * torch.randn(B, C, H, W, generator=g)).to(dev)
ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
grid = torch.stack([xs, ys]).float()[None]
coords = [(grid + 4.0 * torch.randn(B, 2, H, W, generator=g)).to(dev) for _ in range(a.lookups)]
for _ in range(a.builds):
    blk = CorrBlock(f1, f2, num_levels=L, radius=r, mode=a.mode)
for c in coords:
    out = blk(c)
if a.alt:
    alt = AlternateCorrBlock(f1, f2, num_levels=L, radius=r)
    for c in coords[:2]:
        out = alt(c)
torch.cuda.synchronize()
print("ok", float(out.abs().max()))
