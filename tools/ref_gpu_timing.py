"""Same-box baselines on the B200: the reference CorrBlock (torch ops on the GPU: cuBLAS SGEMM + avg_pool2d +
grid_sample) and the reference's own alt_cuda_corr kernels compiled for sm_100 (oracle/_ref), next to ours.
Needs oracle/_ref/reference_raft.tar and oracle/_ref/alt_cuda_corr.so (built where the reference checkout exists).
    python tools/ref_gpu_timing.py [--config cfg2]"""
import argparse, importlib.util, json, os, sys, tarfile, tempfile, warnings
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import CONFIGS, SEED  # noqa: E402
import raft_optical_flow_b200 as rcb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg2")
a = ap.parse_args()
B, C, H, W, r, L, iters, _ = CONFIGS[a.config]
dev = torch.device("cuda:0")
warnings.filterwarnings("ignore")
tmp = tempfile.mkdtemp()
with tarfile.open(os.path.join(ROOT, "oracle", "_ref", "reference_raft.tar")) as t:
    t.extractall(tmp)
sys.path.insert(0, os.path.join(tmp, "core"))
import corr as ref_corr  # noqa: E402  (reference core/corr.py)
spec = importlib.util.spec_from_file_location("alt_cuda_corr", os.path.join(ROOT, "oracle", "_ref", "alt_cuda_corr.so"))
ref_ext = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_ext)

g = torch.Generator(device="cpu").manual_seed(SEED)
f1 = (0.75 * torch.randn(B, C, H, W, generator=g)).to(dev)
f2 = (0.75 * torch.randn(B, C, H, W, generator=g)).to(dev)
ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
coords = [(torch.stack([xs, ys]).float()[None] + 4.0 * torch.randn(B, 2, H, W, generator=g)).to(dev) for _ in range(4)]


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


res = {"config": a.config}
with torch.no_grad():
    res["ref_corrblock_build_us"] = timeit(lambda: ref_corr.CorrBlock(f1, f2, num_levels=L, radius=r))
    blk = ref_corr.CorrBlock(f1, f2, num_levels=L, radius=r)
    res["ref_corrblock_lookup_us"] = timeit(lambda: blk(coords[0]))
    del blk
    res["ours_build_us"] = timeit(lambda: rcb.CorrBlock(f1, f2, num_levels=L, radius=r))
    ours = rcb.CorrBlock(f1, f2, num_levels=L, radius=r)
    res["ours_lookup_us"] = timeit(lambda: ours(coords[0]), reps=20)
    del ours
    # on-the-fly path: reference AlternateCorrBlock on its own compiled kernels vs ours
    ref_corr.alt_cuda_corr = ref_ext
    ralt = ref_corr.AlternateCorrBlock(f1, f2, num_levels=L, radius=r)
    res["ref_alternate_call_us"] = timeit(lambda: ralt(coords[1]), reps=3)
    oalt = rcb.AlternateCorrBlock(f1, f2, num_levels=L, radius=r)
    res["ours_alternate_call_us"] = timeit(lambda: oalt(coords[1]), reps=3)
    # extension level, one level, plus backward
    f1n = f1.permute(0, 2, 3, 1).contiguous(); f2n = f2.permute(0, 2, 3, 1).contiguous()
    cn = coords[2].permute(0, 2, 3, 1).reshape(B, 1, H, W, 2).contiguous()
    res["ref_ext_forward_us"] = timeit(lambda: ref_ext.forward(f1n, f2n, cn, r), reps=3)
    res["ours_ext_forward_us"] = timeit(lambda: rcb.alt_cuda_corr.forward(f1n, f2n, cn, r), reps=3)
    cg = torch.randn(B, 1, (2 * r + 1) ** 2, H, W, device=dev)
    res["ref_ext_backward_us"] = timeit(lambda: ref_ext.backward(f1n, f2n, cn, cg, r), reps=2)
    res["ours_ext_backward_us"] = timeit(lambda: rcb.alt_cuda_corr.backward(f1n, f2n, cn, cg, r), reps=2)
res = {k: (round(v, 1) if isinstance(v, float) else v) for k, v in res.items()}
print(json.dumps(res))
