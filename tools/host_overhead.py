"""Host-side cost of one CorrBlock(...)(coords) call through the Python mirror (wall clock over many calls on a tiny
problem whose kernels are shorter than the enqueue path).   python tools/host_overhead.py"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raft_optical_flow_b200 import CorrBlock
dev = torch.device("cuda:0")
f1 = torch.randn(1, 128, 16, 24, device=dev); f2 = torch.randn(1, 128, 16, 24, device=dev)
ys, xs = torch.meshgrid(torch.arange(16), torch.arange(24), indexing="ij")
c = torch.stack([xs, ys]).float()[None].to(dev)
blk = CorrBlock(f1, f2, radius=3)
for _ in range(100): blk(c)
torch.cuda.synchronize()
n = 5000
t0 = time.perf_counter()
with torch.no_grad():
    for _ in range(n): blk(c)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"lookup call: {1e6 * (t1 - t0) / n:.1f} us host enqueue per call ({1e6 * (t2 - t0) / n:.1f} us incl. drain)")
t0 = time.perf_counter()
with torch.no_grad():
    for _ in range(500): CorrBlock(f1, f2, radius=3)
torch.cuda.synchronize()
print(f"constructor: {1e6 * (time.perf_counter() - t0) / 500:.1f} us per CorrBlock(...)")
