"""On-GPU cross-check of the tcgen05 build modes against the fp32 SIMT build (level by level).
    python tools/check_tc.py            # several shapes"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raft_optical_flow_b200 import CorrBlock  # noqa: E402

dev = torch.device("cuda:0")
shapes = [(1, 64, 16, 16, 4), (2, 24, 11, 13, 3), (1, 128, 24, 40, 4), (2, 256, 55, 128, 4), (1, 256, 47, 156, 4),
          (1, 32, 8, 8, 1)]
ok = True
for direct in ("0",):
    for (B, C, H, W, L) in shapes:
        g = torch.Generator(device="cpu").manual_seed(7)
        f1 = (0.75 * torch.randn(B, C, H, W, generator=g)).to(dev)
        f2 = (0.75 * torch.randn(B, C, H, W, generator=g)).to(dev)
        ref = CorrBlock(f1, f2, num_levels=L, radius=4, mode="fp32").corr_pyramid
        for mode, tol in (("bf16x3", 2e-5), ("bf16", 1e-2)):
            try:
                got = CorrBlock(f1, f2, num_levels=L, radius=4, mode=mode).corr_pyramid
                torch.cuda.synchronize()
            except Exception as e:  # noqa: BLE001
                print(f"{mode} {(B, C, H, W, L)}: EXCEPTION {e}")
                ok = False
                raise SystemExit(1)
            errs = []
            for l in range(L):
                r, t = ref[l].contiguous(), got[l].contiguous()
                e = ((r - t).abs().max() / r.abs().max()).item()
                errs.append(e)
                if not (e < tol):
                    ok = False
                    bad = ((r - t).abs() > tol * r.abs().max()).nonzero()
                    print(f"   level {l}: {bad.shape[0]} bad of {r.numel()}, first {bad[:6].tolist()}, "
                          f"nan={torch.isnan(t).sum().item()}")
            print(f"{mode:7s} {str((B, C, H, W, L)):24s} rel err per level:",
                  " ".join(f"{e:.2e}" for e in errs), "OK" if all(e < tol for e in errs) else "FAIL")
print("ALL OK" if ok else "SOME FAILED")
sys.exit(0 if ok else 1)
