"""Times the fused convex upsampling against the reference's torch formulation (core/raft.py:112-142 restated with
torch ops) on the same GPU.   python tools/time_upsample.py [--config cfg2]"""
import argparse, os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CONFIGS  # noqa: E402
from raft_optical_flow_b200 import upsample_flow  # noqa: E402
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg2")
a = ap.parse_args()
B, C, H, W, r, L, iters, _ = CONFIGS[a.config]
dev = torch.device("cuda:0")
flow = 3 * torch.randn(B, 2, H, W, device=dev)
mask = 2 * torch.randn(B, 576, H, W, device=dev)


def torch_formula(flow, mask):
    N, _, H, W = flow.shape
    m = torch.softmax(mask.view(N, 1, 9, 8, 8, H, W), dim=2)
    up = F.unfold(8 * flow, [3, 3], padding=1).view(N, 2, 9, 1, 1, H, W)
    return torch.sum(m * up, dim=2).permute(0, 1, 4, 2, 5, 3).reshape(N, 2, 8 * H, 8 * W)


def timed(fn, reps=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


with torch.no_grad():
    t_ours = timed(lambda: upsample_flow(flow, mask))
    t_ref = timed(lambda: torch_formula(flow, mask))
    err = (upsample_flow(flow, mask) - torch_formula(flow, mask)).abs().max().item()
nbytes = B * H * W * (576 * 4 + 128 * 4 + 8)
print(f"{a.config}: fused {t_ours:.1f} us ({nbytes / t_ours / 1e3:.0f} GB/s of algorithmic bytes), torch formulation {t_ref:.1f} us, "
      f"max abs diff {err:.2e}")
