"""Checks and times the fused lookup + convc1 kernel (rcb_corr_lookup_convc1) against lookup followed by the
fp32 torch convolution + ReLU (the reference's pair core/raft.py:219 + core/update.py:202).
    python tools/time_lookup_conv.py [--config cfg2] [--cout 256] [--reps 32] [--check-only]"""
import argparse, os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CONFIGS, SEED  # noqa: E402
from raft_optical_flow_b200 import CorrBlock, PackedConvC1  # noqa: E402
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg2")
ap.add_argument("--cout", type=int, default=None)
ap.add_argument("--reps", type=int, default=32)
ap.add_argument("--batch", type=int, default=None)
ap.add_argument("--check-only", action="store_true")
ap.add_argument("--fused-only", action="store_true")
a = ap.parse_args()
B, C, H, W, r, L, iters, _ = CONFIGS[a.config]
if a.batch:
    B = a.batch
cout = a.cout or (256 if r == 4 else 96)
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(SEED)
f1 = (0.75 * torch.randn(B, C, H, W, generator=g)).to(dev)
f2 = (0.75 * torch.randn(B, C, H, W, generator=g)).to(dev)
rd = 2 * r + 1
cin = L * rd * rd
weight = (torch.randn(cout, cin, 1, 1, generator=g) / cin ** 0.5).to(dev)
bias = (0.1 * torch.randn(cout, generator=g)).to(dev)
blk = CorrBlock(f1, f2, num_levels=L, radius=r)
packed = PackedConvC1(weight, bias, L, r)
ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
grid = torch.stack([xs, ys]).float()[None]
coords = [(grid + 4.0 * torch.randn(B, 2, H, W, generator=g)).to(dev).contiguous() for _ in range(8)]
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
got = blk.lookup_conv(coords[0], packed)
corr = blk(coords[0])
ref = F.relu(F.conv2d(corr, weight, bias))
torch.cuda.synchronize()
err = (got - ref).abs().max().item()
print(f"{a.config} B={B} cout={cout}: max abs err {err:.3e}, ref max {ref.abs().max().item():.3f}, "
      f"rel {err / ref.abs().max().item():.2e}; pre-activation check:", end=" ")
got2 = blk.lookup_conv(coords[1], packed, relu=False)
ref2 = F.conv2d(blk(coords[1]), weight, bias)
print(f"{((got2 - ref2).abs().max() / ref2.abs().max()).item():.2e}")
if a.check_only:
    sys.exit(0)

def timeit(fn):
    for i in range(4):
        fn(coords[i])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.reps):
        fn(coords[i % 8])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.reps * 1e3

t_fused = timeit(lambda c: blk.lookup_conv(c, packed))
if os.environ.get("RCB_LCONV_PROF") == "1":
    prof = torch.zeros(24, dtype=torch.int64, device=dev)
    os.environ["RCB_LCONV_PROF_PTR"] = hex(prof.data_ptr())
    blk.lookup_conv(coords[0], packed)
    torch.cuda.synchronize()
    del os.environ["RCB_LCONV_PROF_PTR"]
    pr = prof.double().cpu()
    nm, ne, nt = 148 * 16, 148 * 4, 148
    print("mean cycles per warp: math total %d, wait gather %d, wait s_free %d, resample %d (chunks/warp %.1f) | "
          "mma total %d, wait level_done %d, wait d_free %d, wait weights %d | epilogue total %d, wait acc_full %d, "
          "weights->tmem %d | math fence+arrive %d, issue %d, prologue %d (first gathers %d, weights %d)" % (pr[0] / nm, pr[1] / nm, pr[2] / nm, pr[3] / nm, pr[4] / nm, pr[5] / nt, pr[6] / nt,
                                pr[7] / nt, pr[8] / nt, pr[9] / ne, pr[10] / ne, pr[12] / ne, pr[13] / nm, pr[14] / nm, pr[15] / nm, pr[16] / nm, pr[17] / nm))
if a.fused_only:
    print(f"debug={os.environ.get('RCB_LCONV_DEBUG', '0')} fused {t_fused:.1f} us")
    sys.exit(0)
t_lookup = timeit(lambda c: blk(c))
t_pair32 = timeit(lambda c: F.relu(F.conv2d(blk(c), weight, bias)))
torch.backends.cudnn.allow_tf32 = True
t_pair_tf32 = timeit(lambda c: F.relu(F.conv2d(blk(c), weight, bias)))
print(f"fused {t_fused:.1f} us | lookup alone {t_lookup:.1f} us | lookup + torch conv+relu fp32 {t_pair32:.1f} us, "
      f"tf32 (torch default) {t_pair_tf32:.1f} us")
