"""Times the secondary paths of a config on one GPU (CUDA events, device-resident inputs) and prints one JSON line:
  all-pairs   build + `iters` lookups (what bench.py measures for cfg2)
  alternate   AlternateCorrBlock: prepare + one fused on-the-fly call per iteration
  training    CorrBlock forward with autograd + backward of sum(out * g) over `iters` lookups (a10; cfg5)
with the algorithmic flops / bytes of SURVEY 8(d) next to each time.
    python tools/bench_paths.py --config cfg4 [--skip-train]"""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CONFIGS, SEED, algorithmic_bytes  # noqa: E402
from raft_optical_flow_b200 import AlternateCorrBlock, CorrBlock  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg4")
ap.add_argument("--skip-train", action="store_true")
ap.add_argument("--skip-allpairs", action="store_true")
a = ap.parse_args()
B, C, H, W, r, L, iters, desc = CONFIGS[a.config]
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(SEED)
f1 = (0.75 * torch.randn(B, C, H, W, generator=g)).to(dev)
f2 = (0.75 * torch.randn(B, C, H, W, generator=g)).to(dev)
ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
grid = torch.stack([xs, ys]).float()[None]
coords = [(grid + 4.0 * torch.randn(B, 2, H, W, generator=g)).to(dev) for _ in range(4)]
Q = H * W
rd = 2 * r + 1


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


res = {"config": f"{a.config}: {desc}", "B": B, "C": C, "grid": [H, W], "radius": r, "levels": L, "iters": iters}
build_bytes, lookup_bytes, flops = algorithmic_bytes(B, C, H, W, r, L)
with torch.no_grad():
    if not a.skip_allpairs:
        t_build = timed(lambda: CorrBlock(f1, f2, num_levels=L, radius=r), 3)
        blk = CorrBlock(f1, f2, num_levels=L, radius=r)
        t_look = timed(lambda: blk(coords[0]), 10)
        del blk
        torch.cuda.empty_cache()
        res["allpairs"] = {"build_us": round(t_build, 1), "lookup_us": round(t_look, 1),
                           "pairs_per_s": round(B / ((t_build + iters * t_look) * 1e-6), 1),
                           "build_GBps": round(build_bytes / t_build / 1e3, 0), "lookup_GBps": round(lookup_bytes / t_look / 1e3, 0),
                           "pyramid_GB": round(build_bytes / 1e9, 2)}
    t_prep = timed(lambda: AlternateCorrBlock(f1, f2, num_levels=L, radius=r), 3)
    alt = AlternateCorrBlock(f1, f2, num_levels=L, radius=r)
    t_alt = timed(lambda: alt(coords[1]), 5)
    alt_flops = B * Q * L * (2 * r + 2) ** 2 * 2 * C
    res["alternate"] = {"prepare_us": round(t_prep, 1), "call_us": round(t_alt, 1),
                        "pairs_per_s": round(B / ((t_prep + iters * t_alt) * 1e-6), 1),
                        "fp32_TFLOPs": round(alt_flops / t_alt / 1e6, 2), "algorithmic_flops_per_call": alt_flops}
    del alt
if not a.skip_train:
    go = torch.randn(B, L * rd * rd, H, W, device=dev)

    def train_step():
        a1, a2 = f1.clone().requires_grad_(True), f2.clone().requires_grad_(True)
        blk = CorrBlock(a1, a2, num_levels=L, radius=r)
        loss = 0
        for i in range(iters):
            loss = loss + (blk(coords[i % 4]) * go).sum()
        loss.backward()
        return a1.grad, a2.grad

    t_train = timed(train_step, 3)
    with torch.no_grad():
        t_fwd = timed(lambda: [CorrBlock(f1, f2, num_levels=L, radius=r)(coords[0]) for _ in range(1)], 3)
    res["training"] = {"fwd_bwd_us": round(t_train, 1), "pairs_per_s": round(B / (t_train * 1e-6), 1),
                       "note": f"build + {iters} lookups + backward of all of them (lookup_backward x{iters}, "
                               "pool_backward, contract_backward); includes the torch-side loss arithmetic"}
print(json.dumps(res))
