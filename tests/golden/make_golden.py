"""Generate golden fixtures for the correlation hot path FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference, torch CPU):

    python tests/golden/make_golden.py

It imports the unmodified reference modules (core/corr.py CorrBlock, core/utils/utils.py
coords_grid, core/raft.py RAFT with the shipped raft-small.pth, and the pure-torch
IterativeCorrBlock of liteflownet3_correlation.py as a second formulation of the alternate path),
runs them on seeded inputs on CPU and writes compressed .npz files next to this script.  Nothing
from the reference's source is copied; only numerical inputs/outputs are stored.

The fixtures are the pin for oracle/ (tests/test_oracle_golden.py) and, through the same arrays, for
the CUDA kernels (tests/test_gpu_parity.py).  Random cotangents are regenerated from the recorded
seed with numpy's frozen legacy RandomState, so they are not stored.
"""
import argparse
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("RAFT_REFERENCE", "/root/reference")

warnings.filterwarnings("ignore")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "core"))

from corr import CorrBlock  # noqa: E402  (reference core/corr.py)
from utils.utils import coords_grid  # noqa: E402  (reference core/utils/utils.py)

torch.set_grad_enabled(True)


def rs(seed):
    return np.random.RandomState(seed)


def make_inputs(seed, B, C, H, W, sigma, mean=0.0, std=0.75):
    g = rs(seed)
    f1 = (mean + std * g.standard_normal((B, C, H, W))).astype(np.float32)
    f2 = (mean + std * g.standard_normal((B, C, H, W))).astype(np.float32)
    grid = coords_grid(B, H, W, "cpu").numpy()
    coords = (grid + sigma * g.standard_normal((B, 2, H, W))).astype(np.float32)
    return f1, f2, coords


def cotangent(seed, shape):
    return rs(seed).standard_normal(shape).astype(np.float32)


def run_reference(f1, f2, coords, L, r, grad_seed=None):
    t1 = torch.from_numpy(f1).clone().requires_grad_(grad_seed is not None)
    t2 = torch.from_numpy(f2).clone().requires_grad_(grad_seed is not None)
    tc = torch.from_numpy(coords).clone().requires_grad_(grad_seed is not None)
    blk = CorrBlock(t1, t2, num_levels=L, radius=r)
    out = blk(tc)
    res = {"out": out.detach().numpy()}
    for i, p in enumerate(blk.corr_pyramid):
        res[f"pyr{i}"] = p.detach().numpy()[:, 0]
    if grad_seed is not None:
        go = torch.from_numpy(cotangent(grad_seed, tuple(out.shape)))
        out.backward(go)
        res.update(df1=t1.grad.numpy(), df2=t2.grad.numpy(), dcoords=tc.grad.numpy())
    return res


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB  " +
          " ".join(f"{k}{tuple(v.shape)}" for k, v in arrays.items() if hasattr(v, 'shape') and v.ndim))


def case_odd():
    """Odd sizes: floor-mode pooling drops the trailing row/col (11x13 -> 5x6 -> 2x3)."""
    B, C, H, W, L, r = 2, 24, 11, 13, 3, 3
    f1, f2, coords = make_inputs(101, B, C, H, W, sigma=2.0)
    res = run_reference(f1, f2, coords, L, r, grad_seed=102)
    save("corrblock_odd", meta=np.array([B, C, H, W, L, r, 102]), fmap1=f1, fmap2=f2, coords=coords, **res)


def case_full():
    """RAFT-full style: r=4, L=4, 16x20 grid (levels 16x20, 8x10, 4x5, 2x2), with backward."""
    B, C, H, W, L, r = 1, 32, 16, 20, 4, 4
    f1, f2, coords = make_inputs(201, B, C, H, W, sigma=3.0)
    res = run_reference(f1, f2, coords, L, r, grad_seed=202)
    res.pop("pyr0")  # 410 KB; level 0 is pinned by the odd case and re-derivable from levels >= 1 checks
    save("corrblock_full_r4", meta=np.array([B, C, H, W, L, r, 202]), fmap1=f1, fmap2=f2, coords=coords, **res)


def case_small():
    """RAFT-small style: r=3, L=4, mean-heavy features (SURVEY 7 'Precision')."""
    B, C, H, W, L, r = 2, 16, 16, 16, 4, 3
    f1, f2, coords = make_inputs(301, B, C, H, W, sigma=4.0, mean=1.0, std=1.45)
    res = run_reference(f1, f2, coords, L, r)
    res.pop("pyr0")
    save("corrblock_small_r3", meta=np.array([B, C, H, W, L, r, 0]), fmap1=f1, fmap2=f2, coords=coords, **res)


def case_edges():
    """Border / out-of-range / integer coordinates (SURVEY 8c iii)."""
    B, C, H, W, L, r = 1, 8, 16, 16, 4, 4
    f1, f2, _ = make_inputs(401, B, C, H, W, sigma=0.0)
    grid = coords_grid(B, H, W, "cpu").numpy()
    coords = grid.copy()
    coords[0, :, 0, :] += 1000.0  # far outside -> zeros
    coords[0, :, 1, :] -= 1000.0
    coords[0, 0, 2, :] = 0.0  # left border
    coords[0, 0, 3, :] = W - 1.0  # right border, exactly on the last pixel
    coords[0, 1, 4, :] = -0.5  # half a pixel above the top
    coords[0, 1, 5, :] = H - 0.5
    coords[0, 0, 6, :] = -4.0  # window only partially overlaps
    coords[0, 0, 7, :] = W + 3.25
    coords[0, :, 8, :] += 0.5  # half-pixel offsets
    coords[0, :, 9, :] -= 1e-4  # just below integers
    coords[0, :, 10, :] += 1e-4
    coords = coords.astype(np.float32)
    res = run_reference(f1, f2, coords, L, r)
    res.pop("pyr0")
    save("corrblock_edges", meta=np.array([B, C, H, W, L, r, 0]), fmap1=f1, fmap2=f2, coords=coords, **res)


def case_known_answers():
    """centre channel = <F1[q],F2[q]>/sqrt(C); one-hot window order; +1000 -> zeros."""
    B, C, H, W, L, r = 1, 4, 16, 16, 4, 2
    f1 = np.ones((B, C, H, W), np.float32)
    f2 = np.zeros((B, C, H, W), np.float32)
    f2[0, 0, 5, 10] = 2.0  # one-hot at (x=10, y=5): <f1,f2>/sqrt(4) = 1.0
    coords = coords_grid(B, H, W, "cpu").numpy().astype(np.float32)
    res = run_reference(f1, f2, coords, L, r)
    out = res["out"]
    rd = 2 * r + 1
    lit_a = np.nonzero(out[0, :rd * rd, 5, 9])[0]  # query (x=9,y=5): dx=+1, dy=0
    lit_b = np.nonzero(out[0, :rd * rd, 4, 10])[0]  # query (x=10,y=4): dx=0, dy=+1
    print("one-hot channels lit:", lit_a, out[0, lit_a, 5, 9], lit_b, out[0, lit_b, 4, 10])
    assert list(lit_a) == [(1 + r) * rd + (0 + r)] and list(lit_b) == [(0 + r) * rd + (1 + r)]
    save("corrblock_onehot", meta=np.array([B, C, H, W, L, r, 0]), fmap1=f1, fmap2=f2, coords=coords, out=out)


def case_alt_formulation():
    """The reference's second, pure-torch formulation of the on-the-fly path (IterativeCorrBlock,
    liteflownet3_correlation.py:442-515, 'designed to mimic AlternateCorrBlock'); alt_cuda_corr
    itself only runs on a GPU.  Pooled *features* instead of a pooled volume."""
    from liteflownet3_correlation import IterativeCorrBlock
    B, C, H, W, L, r = 1, 32, 16, 20, 4, 4
    f1, f2, coords = make_inputs(201, B, C, H, W, sigma=3.0)  # same inputs as corrblock_full_r4
    blk = IterativeCorrBlock(torch.from_numpy(f1), torch.from_numpy(f2), radius=r, num_levels=L)
    with torch.no_grad():
        out = blk(torch.from_numpy(coords)).numpy()
    ref = run_reference(f1, f2, coords, L, r)["out"]
    print("IterativeCorrBlock vs CorrBlock max abs diff:", np.abs(out - ref).max())
    save("altcorr_iterative_r4", meta=np.array([B, C, H, W, L, r, 0]), out=out)


def case_raft_small_crop():
    """Real features: RAFT-small (shipped raft-small.pth) on a 128x256 crop of demo frames 0016/0017,
    12 GRU iterations on CPU; records fnet features, the coords fed to the corr block at iterations
    0/5/11 and the corr block's output there (every 2nd query pixel), plus the final flow."""
    import argparse as ap
    import cv2
    import raft as raft_mod
    args = ap.Namespace(small=True, mixed_precision=False, alternate_corr=False)
    model = torch.nn.DataParallel(raft_mod.RAFT(args))
    model.load_state_dict(torch.load(os.path.join(REF, "raft-small.pth"), map_location="cpu"))
    model = model.module.eval()

    def load(name):
        img = cv2.imread(os.path.join(REF, "demo-frames", name))[:, :, ::-1]
        img = img[150:278, 400:656]  # 128 x 256 crop
        return torch.from_numpy(np.ascontiguousarray(img)).permute(2, 0, 1).float()[None]

    im1, im2 = load("frame_0016.png"), load("frame_0017.png")
    rec = {"coords": [], "out": [], "fmaps": None}
    orig_init, orig_call = raft_mod.CorrBlock.__init__, raft_mod.CorrBlock.__call__

    def init(self, fmap1, fmap2, num_levels=4, radius=4):
        rec["fmaps"] = (fmap1.detach().numpy().copy(), fmap2.detach().numpy().copy())
        orig_init(self, fmap1, fmap2, num_levels=num_levels, radius=radius)

    def call(self, coords):
        out = orig_call(self, coords)
        rec["coords"].append(coords.detach().numpy().copy())
        rec["out"].append(out.detach().numpy().copy())
        return out

    raft_mod.CorrBlock.__init__, raft_mod.CorrBlock.__call__ = init, call
    try:
        with torch.no_grad():
            flow_lo, flow_up = model(im1, im2, iters=12, test_mode=True)
    finally:
        raft_mod.CorrBlock.__init__, raft_mod.CorrBlock.__call__ = orig_init, orig_call
    f1, f2 = rec["fmaps"]
    its = [0, 5, 11]
    print("raft-small crop: fmap std %.3f |max| %.2f; flow_lo |max| %.3f" %
          (f1.std(), np.abs(f1).max(), flow_lo.abs().max()))
    save("raft_small_crop", meta=np.array([1, f1.shape[1], f1.shape[2], f1.shape[3], 4, 3, 0]),
         fmap1=f1, fmap2=f2, iters=np.array(its),
         coords=np.stack([rec["coords"][i] for i in its]),
         out_sub=np.stack([rec["out"][i][:, :, ::2, ::2] for i in its]),
         flow_lo=flow_lo.numpy())


def case_upsample_flow():
    """Convex upsampling (next row of the scope table): RAFT.upsample_flow (core/raft.py:112-142) on seeded flow and
    mask tensors of an odd size, with the gradients autograd yields for a seeded cotangent."""
    import raft as raft_mod
    N, H, W, seed = 2, 7, 11, 77
    r = rs(seed)
    flow = torch.from_numpy((3.0 * r.standard_normal((N, 2, H, W))).astype(np.float32)).requires_grad_(True)
    mask = torch.from_numpy((2.0 * r.standard_normal((N, 576, H, W))).astype(np.float32)).requires_grad_(True)
    out = raft_mod.RAFT.upsample_flow(None, flow, mask)  # the method does not use self
    g = cotangent(seed, tuple(out.shape))
    out.backward(torch.from_numpy(g))
    save("upsample_flow", meta=np.array([N, H, W, seed]), flow=flow.detach().numpy(), mask=mask.detach().numpy(),
         out=out.detach().numpy(), dflow=flow.grad.numpy(), dmask=mask.grad.numpy())


def case_motion_encoder_convc1():
    """Lookup followed by the motion encoder's first layer, cor = relu(convc1(corr)) (core/update.py:154,202), with the
    reference's own encoder modules (seeded default initialisation) on the reference CorrBlock output."""
    import argparse as ap
    import update as update_mod
    for name, enc_cls, r, seed, dims in (("basic", update_mod.BasicMotionEncoder, 4, 81, (2, 32, 16, 24)),
                                         ("small", update_mod.SmallMotionEncoder, 3, 82, (1, 24, 17, 19))):
        B, C, H, W = dims
        f1, f2, coords = make_inputs(seed, B, C, H, W, sigma=2.5)
        torch.manual_seed(seed)
        enc = enc_cls(ap.Namespace(corr_levels=4, corr_radius=r))
        with torch.no_grad():
            corr = CorrBlock(torch.from_numpy(f1), torch.from_numpy(f2), num_levels=4, radius=r)(torch.from_numpy(coords))
            cor = torch.relu(enc.convc1(corr))
        save("convc1_" + name, meta=np.array([B, C, H, W, 4, r, seed]), fmap1=f1, fmap2=f2, coords=coords,
             weight=enc.convc1.weight.detach().numpy(), bias=enc.convc1.bias.detach().numpy(), cor=cor.numpy())


def case_sequence_loss():
    """Training harness row (scope table 8f, f4): the reference's sequence_loss (train.py:47-106).  train.py itself cannot
    be imported here (matplotlib / tensorboard / datasets at module level), so the function definition is lifted out of
    the file's syntax tree and executed unchanged; stores loss, metrics and d loss / d predictions."""
    import ast
    src = open(os.path.join(REF, "train.py")).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "sequence_loss")
    ns = {"torch": torch, "MAX_FLOW": 400}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "train.py", "exec"), ns)
    g = rs(901)
    N, H, W, n = 2, 12, 20, 5
    gt = (30.0 * g.standard_normal((N, 2, H, W))).astype(np.float32)
    gt[0, :, 3, 4] = 500.0  # beyond MAX_FLOW: excluded
    preds = [(gt + (4.0 / (i + 1)) * g.standard_normal((N, 2, H, W))).astype(np.float32) for i in range(n)]
    valid = (g.uniform(size=(N, H, W)) > 0.2).astype(np.float32)
    tp = [torch.from_numpy(p).clone().requires_grad_(True) for p in preds]
    loss, metrics = ns["sequence_loss"](tp, torch.from_numpy(gt), torch.from_numpy(valid), gamma=0.8)
    loss.backward()
    save("sequence_loss", meta=np.array([N, H, W, n]), flow_gt=gt, valid=valid, preds=np.stack(preds),
         loss=np.float32(loss.item()), metrics=np.array([metrics[k] for k in ("epe", "1px", "3px", "5px")], np.float32),
         dpreds=np.stack([t.grad.numpy() for t in tp]))


if __name__ == "__main__":
    p = argparse.ArgumentParser()
    p.add_argument("--only", default=None)
    a = p.parse_args()
    cases = [case_odd, case_full, case_small, case_edges, case_known_answers, case_alt_formulation,
             case_raft_small_crop, case_upsample_flow, case_motion_encoder_convc1, case_sequence_loss]
    for c in cases:
        if a.only is None or a.only in c.__name__:
            c()
