"""Whole-model inference of the UNMODIFIED reference RAFT (core/raft.py from oracle/_ref/reference_raft.tar) on one
B200, timed three ways on the same inputs:
    reference      the reference's own CorrBlock / bilinear_sampler / upsample_flow (torch ops)
    patched        raft_optical_flow_b200.patch_raft(raft): our CorrBlock + fused convex upsampling
    patched+fused  patch_raft(raft, fuse_motion_encoder=True): additionally corr lookup + relu(convc1(.)) in one kernel
RAFT-full is random-init (no checkpoint ships with the reference), RAFT-small uses the shipped raft-small.pth; the
end-point distance between the variants' flows is printed next to the times (CUDA events, test_mode, no autograd).

    python tools/raft_inference.py [--batch 8] [--iters 32] [--size 440 1024] [--small]
Prints one JSON line."""
import argparse, json, os, sys, tarfile, tempfile, warnings
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raft_optical_flow_b200 as rcb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--iters", type=int, default=32)
ap.add_argument("--size", type=int, nargs=2, default=[440, 1024])
ap.add_argument("--small", action="store_true")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--mixed-precision", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda:0")
warnings.filterwarnings("ignore")
tmp = tempfile.mkdtemp()
with tarfile.open(os.path.join(ROOT, "oracle", "_ref", "reference_raft.tar")) as t:
    t.extractall(tmp)
sys.path.insert(0, os.path.join(tmp, "core"))
import raft as raft_mod  # noqa: E402  (reference core/raft.py)

torch.manual_seed(1234)
args = argparse.Namespace(small=a.small, mixed_precision=a.mixed_precision, alternate_corr=False, dropout=0.0)
model = raft_mod.RAFT(args)
if a.small:
    sd = torch.load(os.path.join(tmp, "raft-small.pth"), map_location="cpu")
    model.load_state_dict({k.replace("module.", "", 1): v for k, v in sd.items()})
model = model.to(dev).eval()
H, W = a.size
g = torch.Generator(device="cpu").manual_seed(7)
# smooth random images (a low-resolution field upsampled) shifted by a few pixels between the frames
base = torch.nn.functional.interpolate(torch.rand(a.batch, 3, H // 8 + 2, W // 8 + 2, generator=g), size=(H + 16, W + 16),
                                       mode="bicubic", align_corners=False).clamp(0, 1) * 255
im1 = base[:, :, 8:8 + H, 8:8 + W].contiguous().to(dev)
im2 = base[:, :, 5:5 + H, 11:11 + W].contiguous().to(dev)


def run():
    with torch.no_grad():
        return model(im1, im2, iters=a.iters, test_mode=True)[1]


def timed():
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        out = run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.reps, out


def epe(x, y):
    d = (x - y).norm(dim=1)
    return {"mean_px": d.mean().item(), "max_px": d.max().item()}


res = {"model": "RAFT-small (raft-small.pth)" if a.small else "RAFT-full (random init)", "batch": a.batch,
       "size": [H, W], "iters": a.iters, "mixed_precision": a.mixed_precision}
t_ref, f_ref = timed()
old = rcb.patch_raft(raft_mod)
t_pat, f_pat = timed()
raft_mod.CorrBlock, raft_mod.AlternateCorrBlock, raft_mod.RAFT.upsample_flow = old
rcb.patch_raft(raft_mod, fuse_motion_encoder=True)
t_fus, f_fus = timed()
res.update(ms_reference=t_ref, ms_patched=t_pat, ms_patched_fused=t_fus,
           speedup_patched=t_ref / t_pat, speedup_patched_fused=t_ref / t_fus,
           flow_max_abs=f_ref.abs().max().item(),
           epe_patched_vs_reference=epe(f_pat, f_ref), epe_fused_vs_reference=epe(f_fus, f_ref))
print(json.dumps(res))
