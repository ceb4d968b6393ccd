"""Parity of the CUDA path (through the C ABI / Python mirror) against
  (1) golden fixtures produced by the reference itself (tests/golden/),
  (2) the CPU oracle on seeded inputs at sizes it finishes in seconds,
  (3) size-independent properties at BASELINE.json's full sizes,
  (4) the reference's own compiled alt_cuda_corr extension (oracle/_ref) when it travelled to the box.
Tolerance (BASELINE.json north_star): fp32-parity modes max-abs error <= 1e-4 relative to the max-abs of
the reference output; gradients 2e-4; the single-pass bf16 fast mode 5e-3.
"""
import importlib.util
import os

import numpy as np
import pytest
import torch

from conftest import cotangent, load_golden, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-4
GRAD_TOL = 2e-4
BF16_TOL = 5e-3
CONVC1_TOL = 1e-3  # fused lookup + convc1: fp16 tensor-core operands (section 9)
PARITY_MODES = os.environ.get("RCB_TEST_MODES", "fp32,bf16x3,f16f8").split(",")


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device (no CPU fallback exists)"
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def rcb():
    import raft_optical_flow_b200 as pkg
    from raft_optical_flow_b200 import _cabi
    _cabi.lib()
    return pkg


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


def t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def seeded(seed, B, C, H, W, sigma=4.0, mean=0.0, std=0.75):
    rs = np.random.RandomState(seed)
    f1 = (mean + std * rs.standard_normal((B, C, H, W))).astype(np.float32)
    f2 = (mean + std * rs.standard_normal((B, C, H, W))).astype(np.float32)
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    coords = (np.stack([xs, ys])[None] + sigma * rs.standard_normal((B, 2, H, W))).astype(np.float32)
    return f1, f2, coords


# ---------------------------------------------------------------------------------------------
# (1) golden fixtures from the reference
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", PARITY_MODES)
@pytest.mark.parametrize("name", ["corrblock_odd", "corrblock_full_r4", "corrblock_small_r3", "corrblock_edges",
                                  "corrblock_onehot"])
def test_corrblock_golden(rcb, dev, name, mode):
    g = load_golden(name)
    B, C, H, W, L, r, _ = [int(v) for v in g["meta"]]
    blk = rcb.CorrBlock(t(g["fmap1"], dev), t(g["fmap2"], dev), num_levels=L, radius=r, mode=mode)
    assert blk.num_levels == L and blk.radius == r and len(blk.corr_pyramid) == L
    for i in range(L):
        lvl = blk.corr_pyramid[i]
        assert lvl.shape[0] == B * H * W and lvl.shape[1] == 1
        if f"pyr{i}" in g:
            assert rel_err(lvl[:, 0].cpu().numpy(), g[f"pyr{i}"]) < TOL, f"level {i}"
    out = blk(t(g["coords"], dev))
    assert out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == g["out"].shape
    assert rel_err(out.cpu().numpy(), g["out"]) < TOL


def test_known_answers(rcb, dev):
    g = load_golden("corrblock_onehot")
    B, C, H, W, L, r, _ = [int(v) for v in g["meta"]]
    rd = 2 * r + 1
    out = rcb.CorrBlock(t(g["fmap1"], dev), t(g["fmap2"], dev), num_levels=L, radius=r, mode="fp32")(
        t(g["coords"], dev)).cpu().numpy()
    assert list(np.nonzero(out[0, :rd * rd, 5, 9])[0]) == [(1 + r) * rd + r]  # x offset is the slow index
    assert list(np.nonzero(out[0, :rd * rd, 4, 10])[0]) == [r * rd + (1 + r)]
    assert out[0, 17, 5, 9] == 1.0 and out[0, 13, 4, 10] == 1.0
    far = rcb.CorrBlock(t(g["fmap1"], dev), t(g["fmap2"], dev), num_levels=L, radius=r, mode="fp32")(
        t(g["coords"] + 1000.0, dev))
    assert not far.any().item()


def test_real_features_golden(rcb, dev):
    g = load_golden("raft_small_crop")
    B, C, H, W, L, r, _ = [int(v) for v in g["meta"]]
    for mode in PARITY_MODES:
        blk = rcb.CorrBlock(t(g["fmap1"], dev), t(g["fmap2"], dev), num_levels=L, radius=r, mode=mode)
        alt = rcb.AlternateCorrBlock(t(g["fmap1"], dev), t(g["fmap2"], dev), num_levels=L, radius=r)
        for k in range(len(g["iters"])):
            c = t(g["coords"][k], dev)
            assert rel_err(blk(c)[:, :, ::2, ::2].cpu().numpy(), g["out_sub"][k]) < TOL
            assert rel_err(alt(c)[:, :, ::2, ::2].cpu().numpy(), g["out_sub"][k]) < TOL


@pytest.mark.parametrize("mode", PARITY_MODES)  # fp32: SIMT contraction backward, bf16x3: the tcgen05 GEMMs
@pytest.mark.parametrize("name", ["corrblock_odd", "corrblock_full_r4"])
def test_corrblock_backward_golden(rcb, dev, name, mode):
    """Gradients w.r.t. fmap1, fmap2 AND coords vs autograd through the reference CorrBlock."""
    g = load_golden(name)
    B, C, H, W, L, r, seed = [int(v) for v in g["meta"]]
    f1 = t(g["fmap1"], dev).requires_grad_(True)
    f2 = t(g["fmap2"], dev).requires_grad_(True)
    co = t(g["coords"], dev).requires_grad_(True)
    blk = rcb.CorrBlock(f1, f2, num_levels=L, radius=r, mode=mode)
    out = blk(co)
    out.backward(t(cotangent(seed, g["out"].shape), dev))
    assert rel_err(f1.grad.cpu().numpy(), g["df1"]) < GRAD_TOL
    assert rel_err(f2.grad.cpu().numpy(), g["df2"]) < GRAD_TOL
    assert rel_err(co.grad.cpu().numpy(), g["dcoords"]) < GRAD_TOL


@pytest.mark.parametrize("mode", PARITY_MODES)
@pytest.mark.parametrize("dims", [(1, 16, 12, 16), (2, 24, 11, 13), (1, 200, 23, 39)])  # odd sizes, C not a multiple of 16
def test_backward_accumulates_over_iterations(rcb, dev, orc, dims, mode):
    """train.py runs 12 lookups per forward; their gradients must add up in one backward pass."""
    f1n, f2n, c0 = seeded(5, *dims, sigma=2.0)
    _, _, c1 = seeded(6, *dims, sigma=2.0)
    L, r = 3, 3
    f1 = t(f1n, dev).requires_grad_(True)
    f2 = t(f2n, dev).requires_grad_(True)
    blk = rcb.CorrBlock(f1, f2, num_levels=L, radius=r, mode=mode)
    o0, o1 = blk(t(c0, dev)), blk(t(c1, dev))
    g0, g1 = cotangent(11, tuple(o0.shape)), cotangent(12, tuple(o1.shape))
    (o0 * t(g0, dev)).sum().add((o1 * t(g1, dev)).sum()).backward()
    ob = orc.OracleCorrBlock(f1n, f2n, L, r)
    a1, a2, _ = ob.backward(c0, g0)
    b1, b2, _ = ob.backward(c1, g1)
    assert rel_err(f1.grad.cpu().numpy(), a1 + b1) < GRAD_TOL
    assert rel_err(f2.grad.cpu().numpy(), a2 + b2) < GRAD_TOL


@pytest.mark.parametrize("mode", PARITY_MODES)
def test_backward_vs_oracle_training_geometry(rcb, dev, orc, mode):
    """BASELINE.json cfg5 geometry (FlyingChairs 368x496 -> 46x62, C = 256, r = 4, 4 levels), one frame pair, two GRU
    iterations: dF1 / dF2 / dcoords against OracleCorrBlock.backward (what autograd yields through the reference
    CorrBlock, train.py:212).  Q = 2852 = 23 query tiles, K = 23 k-blocks of the tcgen05 backward GEMMs."""
    B, C, H, W, L, r = 1, 256, 46, 62, 4, 4
    f1n, f2n, c0 = seeded(51, B, C, H, W, sigma=3.0)
    _, _, c1 = seeded(52, B, C, H, W, sigma=3.0)
    f1 = t(f1n, dev).requires_grad_(True)
    f2 = t(f2n, dev).requires_grad_(True)
    co0 = t(c0, dev).requires_grad_(True)
    blk = rcb.CorrBlock(f1, f2, num_levels=L, radius=r, mode=mode)
    o0, o1 = blk(co0), blk(t(c1, dev))
    g0, g1 = cotangent(53, tuple(o0.shape)), cotangent(54, tuple(o1.shape))
    (o0 * t(g0, dev)).sum().add((o1 * t(g1, dev)).sum()).backward()
    ob = orc.OracleCorrBlock(f1n, f2n, L, r)
    assert rel_err(o0.detach().cpu().numpy(), ob(c0)) < TOL
    a1, a2, ac = ob.backward(c0, g0)
    b1, b2, _ = ob.backward(c1, g1)
    assert rel_err(f1.grad.cpu().numpy(), a1 + b1) < GRAD_TOL
    assert rel_err(f2.grad.cpu().numpy(), a2 + b2) < GRAD_TOL
    assert rel_err(co0.grad.cpu().numpy(), ac) < GRAD_TOL


# ---------------------------------------------------------------------------------------------
# (2) CPU oracle on seeded inputs
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", PARITY_MODES)
@pytest.mark.parametrize("shape", [(2, 64, 23, 39, 4, 4), (1, 128, 55, 128, 4, 3), (1, 256, 46, 62, 4, 4),
                                   (3, 32, 9, 17, 2, 1), (1, 48, 16, 33, 3, 2),
                                   # the full geometry of BASELINE.json cfg2 / cfg3 (C = 256, r = 4), one frame pair
                                   (1, 256, 55, 128, 4, 4), (1, 256, 47, 156, 4, 4)])
def test_corrblock_vs_oracle(rcb, dev, orc, shape, mode):
    B, C, H, W, L, r = shape
    f1, f2, coords = seeded(100 + H, B, C, H, W)
    want_blk = orc.OracleCorrBlock(f1, f2, L, r)
    blk = rcb.CorrBlock(t(f1, dev), t(f2, dev), num_levels=L, radius=r, mode=mode)
    for i in range(L):
        assert rel_err(blk.corr_pyramid[i][:, 0].cpu().numpy(), want_blk.corr_pyramid[i]) < TOL, f"level {i}"
    assert rel_err(blk(t(coords, dev)).cpu().numpy(), want_blk(coords)) < TOL


def test_mean_heavy_features(rcb, dev, orc):
    """Random-init RAFT-full features are mean-heavy (SURVEY 7 'Precision'): corr mean ~17, |max| ~39."""
    f1, f2, coords = seeded(77, 1, 256, 24, 40, mean=1.0, std=1.45)
    want = orc.OracleCorrBlock(f1, f2, 4, 4)(coords)
    for mode in PARITY_MODES:
        got = rcb.CorrBlock(t(f1, dev), t(f2, dev), mode=mode)(t(coords, dev)).cpu().numpy()
        assert rel_err(got, want) < TOL, mode
    if "bf16x3" in PARITY_MODES:
        fast = rcb.CorrBlock(t(f1, dev), t(f2, dev), mode="bf16")(t(coords, dev)).cpu().numpy()
        assert rel_err(fast, want) < BF16_TOL


def test_corr_static_method(rcb, dev, orc):
    f1, f2, _ = seeded(3, 2, 32, 10, 14)
    v = rcb.CorrBlock.corr(t(f1, dev), t(f2, dev), mode="fp32")
    assert tuple(v.shape) == (2, 10, 14, 1, 10, 14)
    assert rel_err(v.reshape(2, 140, 140).cpu().numpy(), orc.corr_volume(f1, f2)) < TOL


@pytest.mark.parametrize("shape", [(2, 64, 23, 39, 4, 4), (1, 128, 30, 44, 4, 3), (1, 256, 16, 24, 3, 4)])
def test_alternate_block_vs_oracle(rcb, dev, orc, shape):
    B, C, H, W, L, r = shape
    f1, f2, coords = seeded(200 + H, B, C, H, W)
    want = orc.OracleAlternateCorrBlock(f1, f2, L, r)(coords)
    alt = rcb.AlternateCorrBlock(t(f1, dev), t(f2, dev), num_levels=L, radius=r)
    assert rel_err(alt(t(coords, dev)).cpu().numpy(), want) < TOL
    pyr = alt.pyramid
    assert len(pyr) == L + 1 and tuple(pyr[1][1].shape) == (B, C, H // 2, W // 2)


def _ext_inputs(seed, B, N, H1, W1, H2, W2, C, sigma=3.0):
    rs = np.random.RandomState(seed)
    f1 = rs.standard_normal((B, H1, W1, C)).astype(np.float32)
    f2 = rs.standard_normal((B, H2, W2, C)).astype(np.float32)
    ys, xs = np.meshgrid(np.arange(H1), np.arange(W1), indexing="ij")
    base = np.stack([xs, ys], -1)[None, None].astype(np.float32) * (W2 / W1)
    coords = (base + sigma * rs.standard_normal((B, N, H1, W1, 2))).astype(np.float32)
    return f1, f2, coords


@pytest.mark.parametrize("dims", [(2, 1, 12, 20, 12, 20, 64, 4), (1, 2, 16, 24, 8, 12, 128, 3),
                                  (1, 1, 9, 13, 4, 6, 32, 2)])
def test_alt_cuda_corr_extension_vs_oracle(rcb, dev, orc, dims):
    B, N, H1, W1, H2, W2, C, r = dims
    f1, f2, coords = _ext_inputs(31, B, N, H1, W1, H2, W2, C)
    (corr,) = rcb.alt_cuda_corr.forward(t(f1, dev), t(f2, dev), t(coords, dev), r)
    want = orc.altcorr_forward(f1, f2, coords, r)
    assert tuple(corr.shape) == want.shape
    assert rel_err(corr.cpu().numpy(), want) < TOL
    cg = cotangent(32, want.shape)
    g1, g2, gc = rcb.alt_cuda_corr.backward(t(f1, dev), t(f2, dev), t(coords, dev), t(cg, dev), r)
    w1, w2, _ = orc.altcorr_backward(f1, f2, coords, cg, r)
    assert rel_err(g1.cpu().numpy(), w1) < GRAD_TOL
    assert rel_err(g2.cpu().numpy(), w2) < GRAD_TOL
    assert not gc.any().item()  # reference quirk: coords_grad is never written (correlation_kernel.cu:307)
    _, _, gct = rcb.alt_cuda_corr.backward(t(f1, dev), t(f2, dev), t(coords, dev), t(cg, dev), r,
                                           true_coords_grad=True)
    _, _, wct = orc.altcorr_backward(f1, f2, coords, cg, r, true_coords_grad=True)
    assert rel_err(gct.cpu().numpy(), wct) < GRAD_TOL


# ---------------------------------------------------------------------------------------------
# (4) the reference's own compiled extension, on the same GPU
# ---------------------------------------------------------------------------------------------
def _load_ref_ext():
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "alt_cuda_corr.so")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("alt_cuda_corr", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_alt_cuda_corr_vs_reference_extension(rcb, dev):
    ref = _load_ref_ext()
    if ref is None:
        pytest.skip("oracle/_ref/alt_cuda_corr.so not built (reference checkout absent at build time)")
    # H1, W1 multiples of the reference's 4x8 block: its backward reads coords out of bounds otherwise
    B, N, H1, W1, H2, W2, C, r = 2, 1, 16, 24, 8, 12, 256, 4
    f1, f2, coords = _ext_inputs(41, B, N, H1, W1, H2, W2, C)
    a = [t(x, dev) for x in (f1, f2, coords)]
    torch.cuda.synchronize()
    (want,) = ref.forward(*a, r)  # launches on the legacy default stream
    torch.cuda.synchronize()
    (got,) = rcb.alt_cuda_corr.forward(*a, r)
    assert rel_err(got.cpu().numpy(), want.cpu().numpy()) < TOL
    cg = t(cotangent(42, tuple(want.shape)), dev)
    w1, w2, wc = ref.backward(*a, cg, r)
    torch.cuda.synchronize()
    g1, g2, gc = rcb.alt_cuda_corr.backward(*a, cg, r)
    assert rel_err(g1.cpu().numpy(), w1.cpu().numpy()) < GRAD_TOL
    assert rel_err(g2.cpu().numpy(), w2.cpu().numpy()) < GRAD_TOL
    assert not wc.any().item() and not gc.any().item()


# ---------------------------------------------------------------------------------------------
# (3) size-independent properties at full size (BASELINE.json configs)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", PARITY_MODES)
@pytest.mark.parametrize("cfg", [(2, 256, 55, 128, 4), (2, 256, 47, 156, 4), (1, 128, 55, 128, 3)])
def test_full_size_properties(rcb, dev, cfg, mode):
    B, C, H, W, r = cfg
    L, rd = 4, 2 * r + 1
    gen = torch.Generator(device="cpu").manual_seed(1234)
    f1 = (0.75 * torch.randn(B, C, H, W, generator=gen)).to(dev)
    f2 = (0.75 * torch.randn(B, C, H, W, generator=gen)).to(dev)
    blk = rcb.CorrBlock(f1, f2, num_levels=L, radius=r, mode=mode)
    pyr = blk.corr_pyramid
    # level 0 against a plain fp32 (non-TF32) torch contraction in float64 on a slab of queries
    q = torch.arange(0, H * W, 97, device=dev)
    a = f1.reshape(B, C, H * W)[0][:, q].double().T @ f2.reshape(B, C, H * W)[0].double() / np.sqrt(C)
    got0 = pyr[0][:H * W, 0].reshape(H * W, H * W)[q].double()
    assert ((got0 - a).abs().max() / a.abs().max()).item() < TOL
    # level l is the floor-mode 2x2 mean of level l-1 (core/corr.py:52-54)
    for l in range(1, L):
        want = torch.nn.functional.avg_pool2d(pyr[l - 1].contiguous(), 2, stride=2)
        assert tuple(want.shape) == tuple(pyr[l].shape)
        assert ((pyr[l] - want).abs().max() / want.abs().max()).item() < 1e-6
    # integer coords: centre channel of level 0 is <F1[q], F2[q]> / sqrt(C)
    ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    grid = torch.stack([xs, ys]).float()[None].repeat(B, 1, 1, 1)
    out = blk(grid)
    centre = out[:, r * rd + r]
    want = (f1.double() * f2.double()).sum(1) / np.sqrt(C)
    assert ((centre.double() - want).abs().max() / want.abs().max()).item() < TOL
    # far outside -> exact zeros; lookup is linear in the pyramid -> all-pairs == on-the-fly formulation
    assert not blk(grid + 4096.0).any().item()
    coords = grid + 4.0 * torch.randn(B, 2, H, W, generator=gen).to(dev)
    alt = rcb.AlternateCorrBlock(f1, f2, num_levels=L, radius=r)
    o1, o2 = blk(coords), alt(coords)
    assert ((o1 - o2).abs().max() / o2.abs().max()).item() < TOL
    # transposition symmetry of the volume: corr(f1,f2)[q,p] == corr(f2,f1)[p,q]
    v12 = pyr[0][:H * W, 0].reshape(H * W, H * W)
    v21 = rcb.CorrBlock(f2, f1, num_levels=1, radius=r, mode=mode).corr_pyramid[0][:H * W, 0].reshape(H * W, H * W)
    assert ((v12 - v21.T).abs().max() / v12.abs().max()).item() < TOL


@pytest.mark.parametrize("mode", PARITY_MODES)
def test_largest_config_offsets_beyond_32_bits(rcb, dev, mode):
    """cfg4 of BASELINE.json (1088x1920 -> 136x240, batch 4): 4.26e9 level-0 elements (17 GB), i.e. element offsets
    past 2^31 and byte offsets past 2^34.  The on-the-fly formulation never stores the volume, so agreement of the two
    paths on every batch element checks the build's stores and the lookup's gathers at those offsets."""
    B, C, H, W, r, L = 4, 256, 136, 240, 4, 4
    rd = 2 * r + 1
    gen = torch.Generator(device="cpu").manual_seed(4)
    f1 = (0.75 * torch.randn(B, C, H, W, generator=gen)).to(dev)
    f2 = (0.75 * torch.randn(B, C, H, W, generator=gen)).to(dev)
    ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    grid = torch.stack([xs, ys]).float()[None].repeat(B, 1, 1, 1)
    coords = grid + 6.0 * torch.randn(B, 2, H, W, generator=gen).to(dev)
    blk = rcb.CorrBlock(f1, f2, num_levels=L, radius=r, mode=mode)
    assert sum(b.numel() for b in blk._state.pyr.bufs) > 2 ** 32
    centre = blk(grid)[:, r * rd + r]
    want = (f1.double() * f2.double()).sum(1) / np.sqrt(C)
    assert ((centre.double() - want).abs().max() / want.abs().max()).item() < TOL
    o1 = blk(coords)
    wgt = (torch.randn(256, L * rd * rd, 1, 1, generator=gen) / 18.0).to(dev)
    fused = blk.lookup_conv(coords, rcb.PackedConvC1(wgt, None, L, r), relu=False)
    old_tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        unfused = torch.nn.functional.conv2d(o1, wgt)
    finally:
        torch.backends.cudnn.allow_tf32 = old_tf32
    assert ((fused - unfused).abs().max() / unfused.abs().max()).item() < CONVC1_TOL
    del blk, fused, unfused
    o2 = rcb.AlternateCorrBlock(f1, f2, num_levels=L, radius=r)(coords)
    for b in range(B):
        assert ((o1[b] - o2[b]).abs().max() / o2[b].abs().max()).item() < TOL, b


# ---------------------------------------------------------------------------------------------
# boundary behaviour
# ---------------------------------------------------------------------------------------------
def test_non_finite_coordinates_and_single_pixel_levels(rcb, dev, orc):
    """NaN / inf coordinates give zeros (the window is clamped far outside the plane), never NaN or a fault; a level
    that is a single pixel (8 x 8 maps, 4 levels: 8, 4, 2, 1) is sampled in pixel space like every other level --
    the reference's normalisation divides by W_i - 1 = 0 there."""
    rs = np.random.RandomState(5)
    B, C, H, W, L, r = 1, 16, 8, 8, 4, 4
    f1 = rs.standard_normal((B, C, H, W)).astype(np.float32)
    f2 = rs.standard_normal((B, C, H, W)).astype(np.float32)
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    coords = (np.stack([xs, ys])[None] + 1.5 * rs.standard_normal((B, 2, H, W))).astype(np.float32)
    blk = rcb.CorrBlock(t(f1, dev), t(f2, dev), num_levels=L, radius=r)
    alt = rcb.AlternateCorrBlock(t(f1, dev), t(f2, dev), num_levels=L, radius=r)
    want = orc.OracleCorrBlock(f1, f2, num_levels=L, radius=r)(coords, roundtrip=False)
    assert rel_err(blk(t(coords, dev)).cpu().numpy(), want) < TOL
    assert rel_err(alt(t(coords, dev)).cpu().numpy(), want) < TOL
    bad = coords.copy()
    bad[0, 0, 0, 0] = np.nan
    bad[0, 1, 1, 1] = np.inf
    bad[0, 0, 2, 2] = -np.inf
    bad[0, :, 3, 3] = np.nan
    mask = np.ones((H, W), bool)
    for (y, x) in [(0, 0), (1, 1), (2, 2), (3, 3)]:
        mask[y, x] = False
    for block in (blk, alt):  # all-pairs and on-the-fly kernels clamp the same way
        got = block(t(bad, dev)).cpu().numpy()
        assert np.isfinite(got).all()
        for (y, x) in [(0, 0), (1, 1), (2, 2), (3, 3)]:
            assert not got[0, :, y, x].any()
        assert rel_err(got[0][:, mask], want[0][:, mask]) < TOL  # the other queries are untouched


def test_empty_batch_passes_through(rcb, dev):
    """N = 0 goes through the reference's torch ops (empty matmul / pooling / grid_sample) and yields empty tensors of
    the right shape; so it does here, without a launch."""
    f = torch.empty((0, 32, 12, 16), device=dev)
    c = torch.empty((0, 2, 12, 16), device=dev)
    blk = rcb.CorrBlock(f, f, num_levels=3, radius=2)
    assert tuple(blk(c).shape) == (0, 3 * 25, 12, 16)
    assert [tuple(p.shape) for p in blk.corr_pyramid] == [(0, 1, 12, 16), (0, 1, 6, 8), (0, 1, 3, 4)]
    assert tuple(rcb.CorrBlock.corr(f, f).shape) == (0, 12, 16, 1, 12, 16)
    alt = rcb.AlternateCorrBlock(f, f, num_levels=3, radius=2)
    assert tuple(alt(c).shape) == (0, 3 * 25, 12, 16)
    assert len(alt.pyramid) == 4
    with pytest.raises(RuntimeError):
        blk(torch.empty((1, 2, 12, 16), device=dev))
    torch.cuda.synchronize()


def test_error_behaviour_matches_reference_extension(rcb, dev):
    f = torch.zeros(1, 8, 8, 16, device=dev)
    c = torch.zeros(1, 1, 8, 8, 2, device=dev)
    with pytest.raises(RuntimeError, match="fmap2 must be a CUDA tensor"):
        rcb.alt_cuda_corr.forward(f, f.cpu(), c, 4)
    with pytest.raises(RuntimeError, match="fmap1 must be contiguous"):
        rcb.alt_cuda_corr.forward(f.permute(0, 2, 1, 3), f, c, 4)
    with pytest.raises(RuntimeError):
        rcb.CorrBlock(torch.zeros(1, 8, 8, 8, device=dev), torch.zeros(1, 8, 8, 8, device=dev), radius=9)


def test_runs_on_the_callers_stream(rcb, dev, orc):
    f1, f2, coords = seeded(9, 1, 32, 16, 24)
    want = orc.OracleCorrBlock(f1, f2, 4, 4)(coords)
    s = torch.cuda.Stream(device=dev)
    a, b, c = t(f1, dev), t(f2, dev), t(coords, dev)
    torch.cuda.synchronize()
    with torch.cuda.stream(s):
        out = rcb.CorrBlock(a, b, mode="fp32")(c)
    s.synchronize()
    assert rel_err(out.cpu().numpy(), want) < TOL


F16_TOL = 1e-3  # fp16 storage: 2^-11 relative per stored value, at most ~5e-4 of max-abs through the bilinear weights


@pytest.mark.parametrize("shape", [(2, 64, 23, 39, 4, 4), (1, 128, 55, 128, 4, 3), (1, 256, 46, 62, 4, 4),
                                   (3, 32, 9, 17, 2, 1), (1, 48, 16, 33, 3, 2), (1, 256, 47, 156, 4, 4)])
def test_fp16_pyramid_fast_mode(rcb, dev, orc, shape):
    """pyramid_dtype="f16" (RCB_F16): same values as the fp32 pyramid up to fp16 rounding, at every level and through
    the lookup, including plane widths that are not multiples of the 8-column fp16 tiles and out-of-plane windows."""
    B, C, H, W, L, r = shape
    f1n, f2n, cn = seeded(31, B, C, H, W)
    f1, f2, c = t(f1n, dev), t(f2n, dev), t(cn, dev)
    ref = rcb.CorrBlock(f1, f2, num_levels=L, radius=r, mode="bf16x3")
    fast = rcb.CorrBlock(f1, f2, num_levels=L, radius=r, mode="bf16x3", pyramid_dtype="f16")
    for a, b in zip(ref.corr_pyramid, fast.corr_pyramid):
        assert b.dtype == torch.float16 and a.shape == b.shape
        assert rel_err(b.float().cpu().numpy(), a.cpu().numpy()) < F16_TOL
    want = orc.OracleCorrBlock(f1n, f2n, L, r)(cn)
    assert rel_err(fast(c).cpu().numpy(), want) < F16_TOL
    far = c + 1000.0  # every tap outside the plane: exact zeros
    assert fast(far).abs().max().item() == 0.0
    edge = c.clone()
    edge[:, 0] = edge[:, 0] * 0.0 + (W - 1.5)  # windows hanging over the right edge
    want_e = orc.OracleCorrBlock(f1n, f2n, L, r)(edge.cpu().numpy())
    assert rel_err(fast(edge).cpu().numpy(), want_e) < F16_TOL
    both = rcb.CorrBlock(f1, f2, num_levels=L, radius=r, mode="bf16", pyramid_dtype="f16")
    assert rel_err(both(c).cpu().numpy(), want) < BF16_TOL


def test_fp16_pyramid_is_inference_only(rcb, dev):
    f1n, f2n, cn = seeded(32, 1, 32, 12, 16)
    f1 = t(f1n, dev).requires_grad_(True)
    blk = rcb.CorrBlock(f1, t(f2n, dev), num_levels=2, radius=2, pyramid_dtype="f16")
    out = blk(t(cn, dev))
    with pytest.raises(RuntimeError, match="unsupported"):
        out.sum().backward()
    with pytest.raises(RuntimeError, match="unsupported"):
        rcb.CorrBlock(t(f1n, dev), t(f2n, dev), num_levels=2, radius=2, mode="fp32", pyramid_dtype="f16")


def test_planned_and_unplanned_lookup_agree(rcb, dev):
    """rcb_corr_lookup encodes the TMA tensor maps per call, rcb_corr_lookup_planned reuses a caller-owned plan:
    same kernel, bit-identical output (also exercises the raw C ABI without the Python mirror)."""
    from raft_optical_flow_b200 import _cabi
    f1n, f2n, cn = seeded(21, 2, 32, 19, 27)
    L, r = 4, 4
    blk = rcb.CorrBlock(t(f1n, dev), t(f2n, dev), num_levels=L, radius=r)
    c = t(cn, dev)
    want = blk(c)
    got = torch.empty_like(want)
    st = blk._state
    s = torch.cuda.current_stream(dev).cuda_stream
    _cabi.check(_cabi.lib().rcb_corr_lookup(st.pyr.ptrs, c.data_ptr(), got.data_ptr(), 2, 19, 27, L, r, _cabi.F32, s),
                "rcb_corr_lookup")
    torch.cuda.synchronize()
    assert torch.equal(got, want)


@pytest.mark.parametrize("mode", PARITY_MODES + ["bf16"])
def test_build_is_deterministic(rcb, dev, mode):
    """The volume is built by persistent CTAs whose two MMA-issuing warps take alternate tiles and whose tail units are
    cut along the patch sweep: none of that may show in the result.  Three builds of the same pair are bit-identical,
    level by level (odd width, two batches, a tail round)."""
    rs = np.random.RandomState(3)
    f1 = t((0.75 * rs.standard_normal((2, 256, 47, 78))).astype(np.float32), dev)
    f2 = t((0.75 * rs.standard_normal((2, 256, 47, 78))).astype(np.float32), dev)
    ref = [lv.clone() for lv in rcb.CorrBlock(f1, f2, num_levels=4, radius=4, mode=mode).corr_pyramid]
    for _ in range(2):
        again = rcb.CorrBlock(f1, f2, num_levels=4, radius=4, mode=mode).corr_pyramid
        for a_, b_ in zip(ref, again):
            assert torch.equal(a_, b_)


def test_dependent_launches_keep_results_and_stream_order(rcb, dev):
    """Consecutive lookups are launched as programmatic dependents (the next one gathers while the previous one
    drains, its stores wait).  Back-to-back launches without any synchronisation in between must give the results of
    the same calls made one at a time, also when the allocator recycles the output memory of earlier calls."""
    f1n, f2n, _ = seeded(41, 2, 64, 31, 45)
    blk = rcb.CorrBlock(t(f1n, dev), t(f2n, dev), num_levels=4, radius=4)
    cs = [t(seeded(50 + i, 2, 64, 31, 45)[2], dev) for i in range(8)]
    want = []
    for c in cs:  # one at a time, fully synchronised
        want.append(blk(c).clone())
        torch.cuda.synchronize()
    for rep in range(10):
        outs = [blk(c) for c in cs]  # 8 launches back to back, all outputs alive
        torch.cuda.synchronize()
        for o, w in zip(outs, want):
            assert torch.equal(o, w)
        del outs
        ptrs, out, prev = set(), None, None
        for i in range(64):  # rebinding frees the block before last: it is handed to the next call
            prev, out = out, blk(cs[i % 8])
            ptrs.add(out.data_ptr())
        torch.cuda.synchronize()
        assert torch.equal(out, want[63 % 8]) and torch.equal(prev, want[62 % 8])
        assert len(ptrs) <= 4  # the allocator did recycle output memory


# ---------------------------------------------------------------------------------------------
# next row of the scope table: convex upsampling (RAFT.upsample_flow, core/raft.py:112-142)
# ---------------------------------------------------------------------------------------------
def test_upsample_flow_golden_forward_and_backward(rcb, dev):
    g = load_golden("upsample_flow")
    N, H, W, seed = [int(v) for v in g["meta"]]
    flow = t(g["flow"], dev).requires_grad_(True)
    mask = t(g["mask"], dev).requires_grad_(True)
    out = rcb.upsample_flow(flow, mask)
    assert tuple(out.shape) == (N, 2, 8 * H, 8 * W)
    assert rel_err(out.detach().cpu().numpy(), g["out"]) < TOL
    out.backward(t(cotangent(seed, g["out"].shape), dev))
    assert rel_err(flow.grad.cpu().numpy(), g["dflow"]) < GRAD_TOL
    assert rel_err(mask.grad.cpu().numpy(), g["dmask"]) < GRAD_TOL


@pytest.mark.parametrize("dims", [(1, 55, 128), (3, 46, 62), (2, 1, 1), (1, 5, 33)])
def test_upsample_flow_vs_oracle(rcb, dev, orc, dims):
    N, H, W = dims
    rs = np.random.RandomState(5)
    flow = (4.0 * rs.standard_normal((N, 2, H, W))).astype(np.float32)
    mask = (3.0 * rs.standard_normal((N, 576, H, W))).astype(np.float32)
    got = rcb.upsample_flow(t(flow, dev), t(mask, dev)).cpu().numpy()
    assert rel_err(got, orc.upsample_flow(flow, mask)) < TOL
    # uniform logits: every fine pixel is the mean of its 3x3 neighbourhood of 8*flow (zero padded)
    flat = rcb.upsample_flow(t(flow, dev), torch.zeros(N, 576, H, W, device=dev)).cpu().numpy()
    assert rel_err(flat, orc.upsample_flow(flow, np.zeros_like(mask))) < TOL
    assert np.abs(flat[:, :, ::8, ::8] - flat[:, :, 7::8, 7::8]).max() < 1e-5  # constant inside a coarse cell


# ---------------------------------------------------------------------------------------------
# (9) lookup fused with the motion encoder's first layer (scope table 8f, f1): relu(convc1(corr))
# fp16 tensor-core operands, fp32 accumulate -> the stated bound is 1e-3 of max-abs of the fp32 result (measured
# 3e-4); the reference itself runs this convolution in TF32 (same 11-bit significands) under cuDNN's defaults.
# ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize("name", ["convc1_basic", "convc1_small"])
def test_lookup_convc1_golden(rcb, dev, name):
    g = load_golden(name)
    B, C, H, W, L, r, _ = [int(v) for v in g["meta"]]
    blk = rcb.CorrBlock(t(g["fmap1"], dev), t(g["fmap2"], dev), num_levels=L, radius=r)
    packed = rcb.PackedConvC1(t(g["weight"], dev), t(g["bias"], dev), L, r)
    got = blk.lookup_conv(t(g["coords"], dev), packed)
    assert got.shape == g["cor"].shape and got.dtype == torch.float32 and got.is_contiguous()
    assert rel_err(got.cpu().numpy(), g["cor"]) < CONVC1_TOL


@pytest.mark.parametrize("relu,with_bias", [(True, True), (False, False)])
@pytest.mark.parametrize("shape", [(2, 32, 23, 39, 4, 256), (1, 48, 30, 44, 3, 96), (3, 16, 16, 17, 4, 16),
                                   (1, 32, 9, 130, 3, 80)])
def test_lookup_convc1_vs_oracle(rcb, dev, orc, shape, relu, with_bias):
    B, C, H, W, r, cout = shape  # Q = H*W is not a multiple of the 128-query tile
    L = 4 if min(H, W) >= 16 else 3
    f1, f2, coords = seeded(900 + cout, B, C, H, W, sigma=3.0)
    coords[:, :, 0, 0] = -40.0  # a window entirely outside
    rs = np.random.RandomState(cout)
    cin = L * (2 * r + 1) ** 2
    weight = (rs.standard_normal((cout, cin, 1, 1)) / np.sqrt(cin)).astype(np.float32)
    bias = (0.2 * rs.standard_normal(cout)).astype(np.float32) if with_bias else None
    blk = rcb.CorrBlock(t(f1, dev), t(f2, dev), num_levels=L, radius=r)
    packed = rcb.PackedConvC1(t(weight, dev), None if bias is None else t(bias, dev), L, r)
    got = blk.lookup_conv(t(coords, dev), packed, relu=relu).cpu().numpy()
    want = orc.convc1_relu(orc.OracleCorrBlock(f1, f2, num_levels=L, radius=r)(coords, roundtrip=False), weight, bias,
                           relu=relu)
    assert np.isfinite(got).all()
    assert rel_err(got, want) < CONVC1_TOL


def test_lookup_convc1_full_size_matches_unfused_pair(rcb, dev):
    """cfg2 geometry: the fused kernel against its own unfused pair (lookup, then an fp32 torch convolution)."""
    B, C, H, W, r, L, cout = 2, 256, 55, 128, 4, 4, 256
    f1, f2, coords = seeded(31, B, C, H, W)
    rs = np.random.RandomState(5)
    cin = L * (2 * r + 1) ** 2
    weight = t((rs.standard_normal((cout, cin, 1, 1)) / np.sqrt(cin)).astype(np.float32), dev)
    bias = t((0.1 * rs.standard_normal(cout)).astype(np.float32), dev)
    blk = rcb.CorrBlock(t(f1, dev), t(f2, dev), num_levels=L, radius=r)
    packed = rcb.PackedConvC1(weight, bias, L, r)
    c = t(coords, dev)
    got = blk.lookup_conv(c, packed)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        want = torch.relu(torch.nn.functional.conv2d(blk(c), weight, bias))
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert rel_err(got.cpu().numpy(), want.cpu().numpy()) < CONVC1_TOL
    # output channels are independent: permuting the rows of the weight permutes the result bit for bit
    perm = torch.from_numpy(np.random.RandomState(6).permutation(cout)).to(dev)
    p2 = rcb.PackedConvC1(weight[perm].contiguous(), bias[perm].contiguous(), L, r)
    assert torch.equal(blk.lookup_conv(c, p2), got[:, perm])


def test_lookup_convc1_rejects_what_it_does_not_support(rcb, dev):
    f1, f2, coords = seeded(3, 1, 16, 16, 24)
    w = torch.randn(32, 4 * 81, 1, 1, device=dev)
    blk16 = rcb.CorrBlock(t(f1, dev), t(f2, dev), pyramid_dtype="f16")
    with pytest.raises(RuntimeError):
        blk16.lookup_conv(t(coords, dev), rcb.PackedConvC1(w, None))
    with pytest.raises(RuntimeError):
        rcb.PackedConvC1(torch.randn(32, 100, 1, 1, device=dev), None)  # wrong input channel count
    with pytest.raises(RuntimeError):
        rcb.PackedConvC1(torch.randn(24, 4 * 81, 1, 1, device=dev), None)  # cout not a multiple of 16
    with pytest.raises(RuntimeError):
        rcb.PackedConvC1(w.cpu(), None)
    blk = rcb.CorrBlock(t(f1, dev), t(f2, dev), radius=3)
    with pytest.raises(RuntimeError):
        blk.lookup_conv(t(coords, dev), rcb.PackedConvC1(w, None))  # packed for radius 4


# ---------------------------------------------------------------------------------------------
# (10) CUDA graphs: the per-iteration entry points allocate nothing, never synchronise and take the caller's stream,
# so a GRU loop can be captured once and replayed on new coordinates
# ---------------------------------------------------------------------------------------------
def test_lookups_can_be_captured_in_a_cuda_graph(rcb, dev):
    B, C, H, W, r, L = 2, 64, 24, 40, 4, 4
    f1, f2, c0 = seeded(77, B, C, H, W)
    _, _, c1 = seeded(78, B, C, H, W, sigma=6.0)
    blk = rcb.CorrBlock(t(f1, dev), t(f2, dev), num_levels=L, radius=r)
    rs = np.random.RandomState(3)
    wgt = t((rs.standard_normal((96, L * 81, 1, 1)) / 18.0).astype(np.float32), dev)
    packed = rcb.PackedConvC1(wgt, None, L, r)
    static_c = t(c0, dev)
    eager = [(blk(x).clone(), blk.lookup_conv(x, packed).clone()) for x in (t(c0, dev), t(c1, dev))]
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):  # warm-up on the capture stream (allocator pools, function attributes)
        blk(static_c), blk.lookup_conv(static_c, packed)
    torch.cuda.current_stream(dev).wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        a = blk(static_c)          # two dependent-launch lookups back to back, as in the GRU loop
        b = blk(static_c)
        c = blk.lookup_conv(static_c, packed)
    for x, (want_corr, want_conv) in zip((c0, c1), eager):
        static_c.copy_(t(x, dev))
        graph.replay()
        torch.cuda.synchronize(dev)
        assert torch.equal(a, want_corr) and torch.equal(b, want_corr)
        assert torch.equal(c, want_conv)


# ---------------------------------------------------------------------------------------------
# (11) several GPUs in one process, as under nn.DataParallel (reference train.py:172, evaluate.py: one Python thread
# per GPU): blocks on a device that is not the current one, and two devices driven concurrently from two threads
# ---------------------------------------------------------------------------------------------
def test_second_device_and_concurrent_threads(rcb, orc):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    import threading
    B, C, H, W, r, L = 2, 64, 23, 39, 4, 4
    f1, f2, coords = seeded(21, B, C, H, W)
    want = orc.OracleCorrBlock(f1, f2, num_levels=L, radius=r)(coords, roundtrip=False)
    rs = np.random.RandomState(9)
    wgt = (rs.standard_normal((96, L * 81, 1, 1)) / 18.0).astype(np.float32)
    want_conv = orc.convc1_relu(want, wgt, None)
    torch.cuda.set_device(0)
    d1 = torch.device("cuda:1")
    blk = rcb.CorrBlock(t(f1, d1), t(f2, d1), num_levels=L, radius=r)  # current device is 0
    assert rel_err(blk(t(coords, d1)).cpu().numpy(), want) < TOL
    alt = rcb.AlternateCorrBlock(t(f1, d1), t(f2, d1), num_levels=L, radius=r)
    assert rel_err(alt(t(coords, d1)).cpu().numpy(), want) < TOL
    got = blk.lookup_conv(t(coords, d1), rcb.PackedConvC1(t(wgt, d1), None, L, r))
    assert got.device == d1 and rel_err(got.cpu().numpy(), want_conv) < CONVC1_TOL
    errs = {}

    def worker(idx):
        try:
            dev = torch.device("cuda", idx)
            torch.cuda.set_device(dev)
            for _ in range(5):
                b = rcb.CorrBlock(t(f1, dev), t(f2, dev), num_levels=L, radius=r)
                out = b(t(coords, dev))
                conv = b.lookup_conv(t(coords, dev), rcb.PackedConvC1(t(wgt, dev), None, L, r))
                errs[idx] = max(rel_err(out.cpu().numpy(), want), rel_err(conv.cpu().numpy(), want_conv) * TOL / CONVC1_TOL)
        except Exception as e:  # noqa: BLE001 - reported through the assertion below
            errs[idx] = e

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    torch.cuda.set_device(0)
    for i in range(2):
        assert not isinstance(errs.get(i), Exception), errs.get(i)
        assert errs[i] < TOL, (i, errs[i])


# ---------------------------------------------------------------------------------------------
# (12) seeded sweep over small odd geometries: windows that start in every phase of the 4x4 tiles, planes whose last
# tile row / column is ragged, pooled levels down to 1-2 pixels, coordinates far outside on every side
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(12))
def test_lookup_sweep_of_small_geometries_vs_oracle(rcb, dev, orc, seed):
    rs = np.random.RandomState(1000 + seed)
    B, C = int(rs.randint(1, 3)), int(rs.choice([8, 20, 32]))
    H, W = int(rs.randint(8, 41)), int(rs.randint(8, 41))
    r = int(rs.choice([3, 4, 4, 4]))
    L = 4 if min(H, W) >= 16 else 3
    f1 = rs.standard_normal((B, C, H, W)).astype(np.float32)
    f2 = rs.standard_normal((B, C, H, W)).astype(np.float32)
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    coords = (np.stack([xs, ys])[None] + 5.0 * rs.standard_normal((B, 2, H, W))).astype(np.float32)
    coords[:, 0, 0, :] = np.linspace(-12.0, W + 11.0, W, dtype=np.float32)  # sweeps across the left and right borders
    coords[:, 1, :, 0] = np.linspace(-12.0, H + 11.0, H, dtype=np.float32)  # ... and the top and bottom ones
    coords[:, :, -1, -1] = np.float32(3.0) + np.float32(1.0) / np.float32(3.0)  # window starting in tile row phase 3
    want = orc.OracleCorrBlock(f1, f2, num_levels=L, radius=r)(coords, roundtrip=False)
    for pdt, tol in (("f32", TOL), ("f16", 2e-3)):
        blk = rcb.CorrBlock(t(f1, dev), t(f2, dev), num_levels=L, radius=r, pyramid_dtype=pdt)
        got = blk(t(coords, dev)).cpu().numpy()
        assert np.isfinite(got).all()
        assert rel_err(got, want) < tol, (seed, pdt, B, C, H, W, r, L)
        if pdt == "f32":  # both lane mappings of the fp32 kernel (grids this small run 4 lanes per query by default)
            for lanes in (2, 4):
                blk._state.plan.set_lanes(lanes)
                assert np.array_equal(blk(t(coords, dev)).cpu().numpy(), got), (seed, lanes)


@pytest.mark.parametrize("r,L", [(1, 2), (2, 3), (3, 4), (4, 4)])
def test_lookup_lanes_per_query_are_bit_identical(rcb, dev, orc, r, L):
    """rcb_corr_lookup_plan_set_lanes: the two lane mappings of the lookup kernel (2 lanes per query for grids of
    several waves, 4 for small ones) are the same arithmetic in the same order; ragged width, windows beyond every
    border, every radius; and the set_lanes argument is validated."""
    rs = np.random.RandomState(77 + r)
    B, C, H, W = 2, 16, 27, 37
    f1 = rs.standard_normal((B, C, H, W)).astype(np.float32)
    f2 = rs.standard_normal((B, C, H, W)).astype(np.float32)
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    coords = (np.stack([xs, ys])[None] + 6.0 * rs.standard_normal((B, 2, H, W))).astype(np.float32)
    coords[:, 0, 0, :] = np.linspace(-15.0, W + 14.0, W, dtype=np.float32)
    coords[:, 1, :, 0] = np.linspace(-15.0, H + 14.0, H, dtype=np.float32)
    want = orc.OracleCorrBlock(f1, f2, num_levels=L, radius=r)(coords, roundtrip=False)
    for pdt, tol in (("f32", TOL), ("f16", 2e-3)):  # the fp32 kernel and the fp16-pyramid kernel of the fast mode
        blk = rcb.CorrBlock(t(f1, dev), t(f2, dev), num_levels=L, radius=r, pyramid_dtype=pdt)
        outs = {}
        for lanes in (0, 2, 4):
            blk._state.plan.set_lanes(lanes)
            outs[lanes] = blk(t(coords, dev)).cpu().numpy()
            assert rel_err(outs[lanes], want) < tol, (pdt, lanes, r, L)
        assert np.array_equal(outs[2], outs[4]) and np.array_equal(outs[0], outs[4])
        with pytest.raises(RuntimeError):
            blk._state.plan.set_lanes(3)


# ---------------------------------------------------------------------------------------------
# scope table 8f, f2: operands packed once from the encoder's [2N, C, H, W] output
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", [m for m in PARITY_MODES if m != "fp32"] + ["bf16"])
@pytest.mark.parametrize("shape", [(2, 64, 23, 39, 4, 4), (1, 128, 30, 44, 4, 3), (1, 200, 17, 21, 3, 4)])
def test_from_packed_equals_unpacked_and_oracle(rcb, dev, orc, shape, mode):
    """CorrBlock.from_packed(PackedFmaps(fnet output)) == CorrBlock(fmap1, fmap2) bit for bit (the same two kernels,
    launched through rcb_corr_pack_fmaps + rcb_corr_build_packed), and both match the oracle."""
    B, C, H, W, L, r = shape
    f1, f2, coords = seeded(300 + H, B, C, H, W)
    fmaps = t(np.concatenate([f1, f2], 0), dev)  # what fnet([image1, image2]) holds before torch.split
    packed = rcb.PackedFmaps(fmaps, mode=mode)
    a = rcb.CorrBlock.from_packed(packed, num_levels=L, radius=r)
    b = rcb.CorrBlock(fmaps[:B], fmaps[B:], num_levels=L, radius=r, mode=mode)
    for la, lb in zip(a.corr_pyramid, b.corr_pyramid):
        assert torch.equal(la, lb)
    out = a(t(coords, dev))
    assert torch.equal(out, b(t(coords, dev)))
    if mode != "bf16":
        assert rel_err(out.cpu().numpy(), orc.OracleCorrBlock(f1, f2, L, r)(coords)) < TOL
    with pytest.raises(RuntimeError):
        rcb.PackedFmaps(fmaps, mode="fp32")


# ---------------------------------------------------------------------------------------------
# autograd through AlternateCorrBlock (the reference has none; gradients must equal CorrBlock's)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["corrblock_odd", "corrblock_full_r4"])
def test_alternate_block_backward_golden(rcb, dev, name):
    """d/d(fmap1, fmap2, coords) through AlternateCorrBlock against autograd through the REFERENCE CorrBlock (the two
    formulations are equal by linearity, SURVEY 8c vi)."""
    g = load_golden(name)
    B, C, H, W, L, r, seed = [int(v) for v in g["meta"]]
    f1 = t(g["fmap1"], dev).requires_grad_(True)
    f2 = t(g["fmap2"], dev).requires_grad_(True)
    co = t(g["coords"], dev).requires_grad_(True)
    out = rcb.AlternateCorrBlock(f1, f2, num_levels=L, radius=r)(co)
    assert rel_err(out.detach().cpu().numpy(), g["out"]) < TOL
    out.backward(t(cotangent(seed, g["out"].shape), dev))
    assert rel_err(f1.grad.cpu().numpy(), g["df1"]) < GRAD_TOL
    assert rel_err(f2.grad.cpu().numpy(), g["df2"]) < GRAD_TOL
    assert rel_err(co.grad.cpu().numpy(), g["dcoords"]) < GRAD_TOL


def test_alternate_block_backward_accumulates_vs_oracle(rcb, dev, orc):
    """Two calls of one block (two GRU iterations), odd sizes: gradients add up; checked against the oracle's
    CorrBlock backward and, per level, against oracle.altcorr_backward(true_coords_grad=True)."""
    B, C, H, W, L, r = 2, 24, 11, 13, 3, 3
    f1n, f2n, c0 = seeded(61, B, C, H, W, sigma=2.0)
    _, _, c1 = seeded(62, B, C, H, W, sigma=2.0)
    f1 = t(f1n, dev).requires_grad_(True)
    f2 = t(f2n, dev).requires_grad_(True)
    co = t(c0, dev).requires_grad_(True)
    blk = rcb.AlternateCorrBlock(f1, f2, num_levels=L, radius=r)
    o0, o1 = blk(co), blk(t(c1, dev))
    g0, g1 = cotangent(63, tuple(o0.shape)), cotangent(64, tuple(o1.shape))
    (o0 * t(g0, dev)).sum().add((o1 * t(g1, dev)).sum()).backward()
    ob = orc.OracleCorrBlock(f1n, f2n, L, r)
    a1, a2, ac = ob.backward(c0, g0)
    b1, b2, _ = ob.backward(c1, g1)
    assert rel_err(f1.grad.cpu().numpy(), a1 + b1) < GRAD_TOL
    assert rel_err(f2.grad.cpu().numpy(), a2 + b2) < GRAD_TOL
    assert rel_err(co.grad.cpu().numpy(), ac) < GRAD_TOL
    # level 0 of the first call through the extension-level oracle
    rd2 = (2 * r + 1) ** 2
    cg = (g0[:, :rd2] / np.sqrt(C)).reshape(B, 1, rd2, H, W)
    _, _, wc = orc.altcorr_backward(f1n.transpose(0, 2, 3, 1), f2n.transpose(0, 2, 3, 1),
                                    c0.transpose(0, 2, 3, 1).reshape(B, 1, H, W, 2), cg, r, true_coords_grad=True)
    only0 = np.zeros_like(g0)
    only0[:, :rd2] = g0[:, :rd2]
    co2 = t(c0, dev).requires_grad_(True)
    rcb.AlternateCorrBlock(t(f1n, dev), t(f2n, dev), num_levels=L, radius=r)(co2).backward(t(only0, dev))
    assert rel_err(co2.grad.cpu().numpy(), wc.reshape(B, H, W, 2).transpose(0, 3, 1, 2)) < GRAD_TOL
