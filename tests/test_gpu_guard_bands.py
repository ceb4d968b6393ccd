"""Out-of-bounds write detector for every kernel behind the C ABI.

Each output (and workspace) pointer handed to the library is the interior of a larger allocation whose
leading and trailing bands hold a sentinel bit pattern; after the call the bands must be untouched.  The
shapes are the awkward ones (odd sizes, channel counts that are not multiples of the tile sizes, ragged last
tiles) because that is where a clipped tile or a tail loop would write past the end.  Results themselves are
checked by test_gpu_parity.py; here only the footprint is.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

GUARD = 4096  # bytes on either side; a multiple of every alignment the library asks for
SENTINEL = 0x5A


class Guarded:
    """`nbytes` of device memory between two sentinel bands."""

    def __init__(self, nbytes, dev, fill=None):
        self.nbytes = int(nbytes)
        body = (self.nbytes + 255) // 256 * 256
        self.raw = torch.full((GUARD + body + GUARD,), SENTINEL, dtype=torch.uint8, device=dev)
        self.tail_from = GUARD + self.nbytes
        if fill is not None:
            self.raw[GUARD:self.tail_from] = fill

    @property
    def ptr(self):
        return self.raw.data_ptr() + GUARD

    def view(self, dtype, shape):
        return self.raw[GUARD:self.tail_from].view(dtype).view(shape)

    def check(self, what):
        torch.cuda.synchronize()
        head = self.raw[:GUARD]
        tail = self.raw[self.tail_from:]
        assert bool((head == SENTINEL).all()), f"{what}: bytes before the buffer were overwritten"
        bad = (tail != SENTINEL).nonzero()
        assert bad.numel() == 0, f"{what}: {bad.numel()} bytes past the end were overwritten (first at +{int(bad[0])})"


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def cabi():
    from raft_optical_flow_b200 import _cabi
    _cabi.lib()
    return _cabi


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _inputs(B, C, H, W, dev, sigma=3.0):
    g = torch.Generator(device="cpu").manual_seed(B * 1000 + C + H * W)
    f1 = (0.75 * torch.randn(B, C, H, W, generator=g)).to(dev)
    f2 = (0.75 * torch.randn(B, C, H, W, generator=g)).to(dev)
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    grid = torch.stack([xs, ys]).float()[None].repeat(B, 1, 1, 1)
    coords = (grid + sigma * torch.randn(B, 2, H, W, generator=g)).to(dev)
    # some windows entirely outside, some straddling every border
    coords[:, :, 0, 0] = -50.0
    coords[:, 0, -1, -1] = W + 2.5
    coords[:, 1, -1, -1] = H + 1.25
    return f1.contiguous(), f2.contiguous(), coords.contiguous()


SHAPES = [(1, 20, 9, 11), (2, 96, 13, 22), (1, 256, 23, 39), (1, 136, 17, 8)]


@pytest.mark.parametrize("pyr_dtype", ["f32", "f16"])
@pytest.mark.parametrize("mode", ["f16f8", "bf16x3", "bf16", "fp32"])
@pytest.mark.parametrize("shape", SHAPES)
def test_build_and_lookup_stay_inside_their_buffers(cabi, dev, shape, mode, pyr_dtype):
    if mode == "fp32" and pyr_dtype == "f16":
        pytest.skip("the SIMT build has no fp16 pyramid")
    B, C, H, W = shape
    lib = cabi.lib()
    f1, f2, coords = _inputs(B, C, H, W, dev)
    levels = 4 if min(H, W) >= 8 else 3
    dt = {"f32": cabi.F32, "f16": cabi.F16}[pyr_dtype]
    lay = cabi.pyramid_layout(B, H, W, levels, dt)
    lv = [Guarded(lay.level_bytes[l], dev) for l in range(levels)]
    ptrs = cabi.ptr_array([g.ptr for g in lv])
    m = cabi.BUILD_MODES[mode]
    nws = lib.rcb_corr_build_workspace_bytes(B, C, H, W, m)
    ws = Guarded(max(nws, 16), dev)
    cabi.check(lib.rcb_corr_build(f1.data_ptr(), f2.data_ptr(), ptrs, B, C, H, W, levels, m, dt, ws.ptr, nws,
                                  _stream()), "rcb_corr_build")
    for l, g in enumerate(lv):
        g.check(f"build {mode}/{pyr_dtype} level {l}")
    ws.check(f"build {mode} workspace")

    for radius in (3, 4):
        rd = 2 * radius + 1
        out = Guarded(B * levels * rd * rd * H * W * 4, dev)
        cabi.check(lib.rcb_corr_lookup(ptrs, coords.data_ptr(), out.ptr, B, H, W, levels, radius, dt, _stream()),
                   "rcb_corr_lookup")
        out.check(f"lookup r={radius}")
        plan = cabi.LookupPlan(ptrs, B, H, W, levels, radius, dt)
        out2 = Guarded(B * levels * rd * rd * H * W * 4, dev)
        cabi.check(lib.rcb_corr_lookup_planned(plan.ptr, coords.data_ptr(), out2.ptr, _stream()),
                   "rcb_corr_lookup_planned")
        out2.check(f"planned lookup r={radius}")
        a = out.view(torch.float32, (B, levels * rd * rd, H, W))
        b = out2.view(torch.float32, (B, levels * rd * rd, H, W))
        assert torch.isfinite(a).all() and torch.equal(a, b)


@pytest.mark.parametrize("shape", SHAPES)
def test_backward_kernels_stay_inside_their_buffers(cabi, dev, shape):
    B, C, H, W = shape
    lib = cabi.lib()
    f1, f2, coords = _inputs(B, C, H, W, dev)
    levels = 4 if min(H, W) >= 8 else 3
    radius = 4
    rd = 2 * radius + 1
    lay = cabi.pyramid_layout(B, H, W, levels, cabi.F32)
    pyr = [torch.randn(lay.level_bytes[l] // 4, device=dev) for l in range(levels)]
    pptrs = cabi.ptr_array([p.data_ptr() for p in pyr])
    go = torch.randn(B, levels * rd * rd, H, W, device=dev)
    dp = [Guarded(lay.level_bytes[l], dev, fill=0) for l in range(levels)]
    dptrs = cabi.ptr_array([g.ptr for g in dp])
    dco = Guarded(B * 2 * H * W * 4, dev)
    cabi.check(lib.rcb_corr_lookup_backward(pptrs, coords.data_ptr(), go.data_ptr(), dptrs, dco.ptr, B, H, W, levels,
                                            radius, cabi.F32, _stream()), "rcb_corr_lookup_backward")
    cabi.check(lib.rcb_corr_pool_backward(dptrs, B, H, W, levels, _stream()), "rcb_corr_pool_backward")
    for l, g in enumerate(dp):
        g.check(f"lookup/pool backward level {l}")
    dco.check("lookup backward dcoords")

    n = B * C * H * W * 4
    for tc in (False, True):
        if tc and C > 256:
            continue
        df1, df2 = Guarded(n, dev), Guarded(n, dev)
        if tc:
            nws = lib.rcb_corr_contract_backward_tc_workspace_bytes(B, C, H, W)
            ws = Guarded(max(nws, 256), dev)
            cabi.check(lib.rcb_corr_contract_backward_tc(f1.data_ptr(), f2.data_ptr(), dp[0].ptr, df1.ptr, df2.ptr,
                                                         B, C, H, W, ws.ptr, nws, _stream()),
                       "rcb_corr_contract_backward_tc")
            ws.check("contract backward (tensor cores) workspace")
        else:
            cabi.check(lib.rcb_corr_contract_backward(f1.data_ptr(), f2.data_ptr(), dp[0].ptr, df1.ptr, df2.ptr,
                                                      B, C, H, W, _stream()), "rcb_corr_contract_backward")
        df1.check(f"contract backward tc={tc} dfmap1")
        df2.check(f"contract backward tc={tc} dfmap2")
        assert torch.isfinite(df1.view(torch.float32, (B, C, H, W))).all()
        assert torch.isfinite(df2.view(torch.float32, (B, C, H, W))).all()


@pytest.mark.parametrize("dims", [(1, 1, 9, 11, 9, 11, 20, 4), (2, 2, 13, 22, 6, 11, 96, 3),
                                  (1, 1, 23, 39, 11, 19, 128, 4), (1, 1, 17, 8, 17, 8, 256, 4)])
def test_altcorr_kernels_stay_inside_their_buffers(cabi, dev, dims):
    B, N, H1, W1, H2, W2, C, r = dims
    lib = cabi.lib()
    rd = 2 * r + 1
    f1 = torch.randn(B, H1, W1, C, device=dev)
    f2 = torch.randn(B, H2, W2, C, device=dev)
    coords = torch.rand(B, N, H1, W1, 2, device=dev) * torch.tensor([W2 + 6.0, H2 + 6.0], device=dev) - 3.0
    coords = coords.contiguous()
    corr = Guarded(B * N * rd * rd * H1 * W1 * 4, dev, fill=0)
    cabi.check(lib.rcb_altcorr_forward(f1.data_ptr(), f2.data_ptr(), coords.data_ptr(), corr.ptr, B, N, H1, W1, H2,
                                       W2, C, r, _stream()), "rcb_altcorr_forward")
    corr.check("altcorr forward")
    cg = torch.randn(B, N, rd * rd, H1, W1, device=dev)
    for true_cg in (0, 1):
        g1 = Guarded(f1.numel() * 4, dev, fill=0)
        g2 = Guarded(f2.numel() * 4, dev, fill=0)
        gc = Guarded(coords.numel() * 4, dev, fill=0)
        cabi.check(lib.rcb_altcorr_backward(f1.data_ptr(), f2.data_ptr(), coords.data_ptr(), cg.data_ptr(), g1.ptr,
                                            g2.ptr, gc.ptr, B, N, H1, W1, H2, W2, C, r, true_cg, _stream()),
                   "rcb_altcorr_backward")
        g1.check("altcorr backward fmap1_grad")
        g2.check("altcorr backward fmap2_grad")
        gc.check("altcorr backward coords_grad")


@pytest.mark.parametrize("shape", SHAPES)
def test_altcorr_pyramid_path_stays_inside_its_buffers(cabi, dev, shape):
    B, C, H, W = shape
    lib = cabi.lib()
    f1, f2, coords = _inputs(B, C, H, W, dev)
    levels = 4 if min(H, W) >= 8 else 3
    f1n = Guarded(B * H * W * C * 4, dev)
    f2n, h, w = [], H, W
    for _ in range(levels):
        f2n.append(Guarded(B * h * w * C * 4, dev))
        h, w = h // 2, w // 2
    f2p = cabi.ptr_array([g.ptr for g in f2n])
    cabi.check(lib.rcb_altcorr_prepare(f1.data_ptr(), f2.data_ptr(), f1n.ptr, f2p, B, C, H, W, levels, _stream()),
               "rcb_altcorr_prepare")
    f1n.check("altcorr prepare fmap1")
    for l, g in enumerate(f2n):
        g.check(f"altcorr prepare fmap2 level {l}")
    for r in (3, 4):
        rd = 2 * r + 1
        out = Guarded(B * levels * rd * rd * H * W * 4, dev)
        cabi.check(lib.rcb_altcorr_pyramid_forward(f1n.ptr, f2p, coords.data_ptr(), out.ptr, B, C, H, W, levels, r,
                                                   1.0 / math.sqrt(C), _stream()), "rcb_altcorr_pyramid_forward")
        out.check(f"altcorr pyramid forward r={r}")
        assert torch.isfinite(out.view(torch.float32, (B, levels * rd * rd, H, W))).all()


@pytest.mark.parametrize("dims", [(1, 1, 1), (2, 5, 33), (1, 13, 22), (3, 46, 62)])
def test_upsample_kernels_stay_inside_their_buffers(cabi, dev, dims):
    N, H, W = dims
    lib = cabi.lib()
    flow = torch.randn(N, 2, H, W, device=dev)
    mask = torch.randn(N, 576, H, W, device=dev)
    out = Guarded(N * 2 * 8 * H * 8 * W * 4, dev)
    cabi.check(lib.rcb_upsample_flow(flow.data_ptr(), mask.data_ptr(), out.ptr, N, H, W, _stream()),
               "rcb_upsample_flow")
    out.check("upsample forward")
    go = torch.randn(N, 2, 8 * H, 8 * W, device=dev)
    dflow = Guarded(flow.numel() * 4, dev)
    dmask = Guarded(mask.numel() * 4, dev)
    nws = lib.rcb_upsample_flow_backward_workspace_bytes(N, H, W)
    ws = Guarded(max(nws, 16), dev)
    cabi.check(lib.rcb_upsample_flow_backward(flow.data_ptr(), mask.data_ptr(), go.data_ptr(), dflow.ptr, dmask.ptr,
                                              ws.ptr, nws, N, H, W, _stream()), "rcb_upsample_flow_backward")
    dflow.check("upsample backward dflow")
    dmask.check("upsample backward dmask")
    ws.check("upsample backward workspace")


@pytest.mark.parametrize("shape,radius,cout", [((2, 32, 13, 22), 4, 256), ((1, 24, 17, 19), 3, 96),
                                               ((1, 16, 9, 130), 4, 16)])
def test_fused_lookup_convc1_stays_inside_its_buffers(cabi, dev, shape, radius, cout):
    B, C, H, W = shape
    lib = cabi.lib()
    f1, f2, coords = _inputs(B, C, H, W, dev)
    levels = 4 if min(H, W) >= 16 else 3
    lay = cabi.pyramid_layout(B, H, W, levels, cabi.F32)
    lv = [torch.empty(lay.level_bytes[l] // 4, device=dev) for l in range(levels)]
    ptrs = cabi.ptr_array([x.data_ptr() for x in lv])
    m = cabi.BUILD_MODES["bf16x3"]
    nws = lib.rcb_corr_build_workspace_bytes(B, C, H, W, m)
    ws = torch.empty(max(nws, 16), dtype=torch.uint8, device=dev)
    cabi.check(lib.rcb_corr_build(f1.data_ptr(), f2.data_ptr(), ptrs, B, C, H, W, levels, m, cabi.F32, ws.data_ptr(),
                                  nws, _stream()), "rcb_corr_build")
    plan = cabi.LookupPlan(ptrs, B, H, W, levels, radius, cabi.F32)
    cin = levels * (2 * radius + 1) ** 2
    weight = torch.randn(cout, cin, device=dev) / cin ** 0.5
    bias = torch.randn(cout, device=dev)
    npk = lib.rcb_corr_convc1_pack_bytes(cout, levels, radius)
    assert npk > 0
    wp = Guarded(npk, dev)
    cabi.check(lib.rcb_corr_convc1_pack(weight.data_ptr(), wp.ptr, cout, levels, radius, _stream()),
               "rcb_corr_convc1_pack")
    wp.check("convc1 weight pack")
    out = Guarded(B * cout * H * W * 4, dev)
    cabi.check(lib.rcb_corr_lookup_convc1(plan.ptr, coords.data_ptr(), wp.ptr, bias.data_ptr(), out.ptr, cout, 1,
                                          _stream()), "rcb_corr_lookup_convc1")
    out.check("fused lookup + convc1")
    assert torch.isfinite(out.view(torch.float32, (B, cout, H, W))).all()
