"""Drop-in replacement for the reference's ``core/corr.py``.

Same classes, constructor/call signatures, attribute names and output layout as the reference
(``CorrBlock`` core/corr.py:12-127, ``AlternateCorrBlock`` core/corr.py:130-198); the arithmetic runs in
hand-written sm_100a kernels behind the C ABI of ``include/raft_corr_b200.h``.  PyTorch is used for
device memory, streams and autograd plumbing only.  There is no CPU path: tensors must live on a CUDA
device (same error text as the reference extension, alt_cuda_corr/correlation.cpp:19).

Build modes of the all-pairs volume (``mode=`` / env ``RAFT_CORR_MODE``):
  "f16f8" (default)   tcgen05 tensor cores: fp16 hi x hi pass + the two cross terms in 8-bit e4m3 (kind::f8f6f4),
                      fp32 accumulate -> fp32-parity (~1e-5 of max-abs, bound 1e-4); |features| < 65504
  "bf16x3"            tcgen05 tensor cores, hi/lo bf16 split, three passes, fp32 accumulate -> fp32-parity (~5e-6)
  "fp32"              fp32 FMA pipe, the reference's exact arithmetic class (also used for C > 256)
  "bf16"              single-pass bf16 operands (fast mode, ~1e-3 rel)
"""
import math
import os

import torch
import torch.nn.functional as F

from . import _cabi

__all__ = ["CorrBlock", "AlternateCorrBlock", "PackedConvC1", "PackedFmaps"]

DEFAULT_MODE = os.environ.get("RAFT_CORR_MODE", "f16f8")


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _check_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")


def _prep(t, name):
    _check_cuda(t, name)
    if t.dtype is torch.float32 and not t.requires_grad and t.is_contiguous():
        return t  # the common case (core/raft.py:181-182,216 hands over detached fp32 tensors): no torch calls
    return t.detach().float().contiguous()


class _Pyramid:
    """Device buffers of one correlation pyramid in the layout rcb_pyramid_layout_query reports."""

    def __init__(self, B, H, W, levels, device, dtype=_cabi.F32, zero=False):
        self.B, self.H, self.W, self.levels, self.dtype = B, H, W, levels, dtype
        self.layout = _cabi.pyramid_layout(B, H, W, levels, dtype)
        tdtype = torch.float32 if dtype == _cabi.F32 else torch.float16
        esize = 4 if dtype == _cabi.F32 else 2
        alloc = torch.zeros if zero else torch.empty
        self.bufs = [alloc(self.layout.level_bytes[l] // esize, dtype=tdtype, device=device) for l in range(levels)]
        self.ptrs = _cabi.ptr_array([b.data_ptr() for b in self.bufs])

    def level(self, l):
        """Level l as a dense [B*H*W, 1, H_l, W_l] tensor, the shape of the reference's corr_pyramid entries
        (core/corr.py:44-54).  The kernels keep planes as 64-byte tiles (rcb_pyramid_layout); this un-tiles a
        copy with plain tensor ops and is meant for inspection/tests, not for the hot path."""
        lay, n = self.layout, self.B * self.H * self.W
        ty, tx, tw = lay.tiles_y[l], lay.tiles_x[l], lay.tile_w
        t = self.bufs[l].view(n, ty, tx, 4, tw).permute(0, 1, 3, 2, 4).reshape(n, ty * 4, tx * tw)
        return t[:, :lay.H[l], :lay.W[l]].unsqueeze(1)

    def views(self):
        return [self.level(l) for l in range(self.levels)]

    def zero_(self):
        for b in self.bufs:
            b.zero_()


def _build(f1, f2, levels, mode, pyr):
    B, C, H, W = f1.shape
    lib = _cabi.lib()
    if C > 256 and pyr.dtype == _cabi.F32:
        mode = _cabi.BUILD_FP32_SIMT  # the tensor-core kernels keep the query operand in tensor memory: C <= 256
    ws_bytes = lib.rcb_corr_build_workspace_bytes(B, C, H, W, mode)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=f1.device)
    with torch.cuda.device(f1.device):
        _cabi.check(lib.rcb_corr_build(f1.data_ptr(), f2.data_ptr(), pyr.ptrs, B, C, H, W, levels, mode, pyr.dtype,
                                       ws.data_ptr(), ws_bytes, _stream(f1)), "rcb_corr_build")
    ws.record_stream(torch.cuda.current_stream(f1.device))


class _BuildFn(torch.autograd.Function):
    """Volume + pyramid.  The pyramid lives in ``state``; autograd sees a scalar token so that this node's
    backward runs once, after every lookup of the forward pass has scattered its gradient."""

    @staticmethod
    def forward(ctx, fmap1, fmap2, state):
        ctx.state = state
        return torch.zeros((), device=fmap1.device)

    @staticmethod
    def backward(ctx, _grad_token):
        st = ctx.state
        B, C, H, W = st.f1.shape
        lib = _cabi.lib()
        df1 = torch.empty_like(st.f1)
        df2 = torch.empty_like(st.f2)
        if st.dpyr is None:  # no lookup contributed a gradient
            return df1.zero_(), df2.zero_(), None
        with torch.cuda.device(st.f1.device):
            s = _stream(st.f1)
            _cabi.check(lib.rcb_corr_pool_backward(st.dpyr.ptrs, B, H, W, st.levels, s), "rcb_corr_pool_backward")
            if st.mode == _cabi.BUILD_FP32_SIMT or C > 256:  # the reference's exact arithmetic class: fp32 FMA tiles
                _cabi.check(lib.rcb_corr_contract_backward(st.f1.data_ptr(), st.f2.data_ptr(), st.dpyr.bufs[0].data_ptr(),
                                                           df1.data_ptr(), df2.data_ptr(), B, C, H, W, s),
                            "rcb_corr_contract_backward")
            else:  # tensor cores, hi/lo bf16 split like the forward build
                nws = lib.rcb_corr_contract_backward_tc_workspace_bytes(B, C, H, W)
                ws = torch.empty(max(nws, 256), dtype=torch.uint8, device=st.f1.device)
                _cabi.check(lib.rcb_corr_contract_backward_tc(st.f1.data_ptr(), st.f2.data_ptr(),
                                                              st.dpyr.bufs[0].data_ptr(), df1.data_ptr(), df2.data_ptr(),
                                                              B, C, H, W, ws.data_ptr(), nws, s),
                            "rcb_corr_contract_backward_tc")
                ws.record_stream(torch.cuda.current_stream(st.f1.device))
        st.dpyr = None  # a second backward pass starts from zero again
        return df1, df2, None


class _LookupFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coords, token, state):
        ctx.state = state
        c = _prep(coords, "coords")
        ctx.save_for_backward(c)
        ctx.needs_pyr = token is not None and token.requires_grad
        return state.lookup(c)

    @staticmethod
    def backward(ctx, grad_out):
        st = ctx.state
        (c,) = ctx.saved_tensors
        B, _, H, W = c.shape
        go = grad_out.float().contiguous()
        want_c = ctx.needs_input_grad[0]
        dco = torch.empty_like(c) if want_c else None
        if ctx.needs_pyr and st.dpyr is None:
            st.dpyr = _Pyramid(B, H, W, st.levels, c.device, _cabi.F32, zero=True)
        if ctx.needs_pyr or want_c:
            with torch.cuda.device(c.device):
                _cabi.check(_cabi.lib().rcb_corr_lookup_backward(
                    st.pyr.ptrs, c.data_ptr(), go.data_ptr(), st.dpyr.ptrs if ctx.needs_pyr else None,
                    dco.data_ptr() if want_c else None, B, H, W, st.levels, st.radius, st.pyr.dtype, _stream(c)),
                    "rcb_corr_lookup_backward")
        dtoken = torch.zeros((), device=c.device) if ctx.needs_pyr else None
        return dco, dtoken, None


class PackedFmaps:
    """Tensor-core operands of the volume build, packed once from the feature encoder's output for the concatenated
    frame pair (scope table 8f, f2): ``fmaps`` is the [2N, C, H, W] tensor fnet returns before ``torch.split``
    (reference core/extractor.py:184-190, core/raft.py:178-182) -- first N maps = fmap1 (queries), last N = fmap2
    (targets).  ``CorrBlock.from_packed(packed)`` then runs only the volume + pyramid kernel.  The operand layout
    (rcb_corr_pack_fmaps) is what a fused epilogue of the encoder's last 1x1 convolution would have to emit."""

    def __init__(self, fmaps, mode=None):
        _check_cuda(fmaps, "fmaps")
        if fmaps.dim() != 4 or fmaps.shape[0] % 2:
            raise RuntimeError("fmaps must be the [2N, C, H, W] encoder output of the concatenated frame pair")
        f = _prep(fmaps, "fmaps")
        self.shape = (f.shape[0] // 2,) + tuple(f.shape[1:])
        B, C, H, W = self.shape
        self.mode = _cabi.BUILD_MODES[mode or DEFAULT_MODE]
        if self.mode == _cabi.BUILD_FP32_SIMT or C > 256:
            raise RuntimeError("packed operands exist for the tensor-core build modes only (C <= 256)")
        lib = _cabi.lib()
        self.nbytes = lib.rcb_corr_build_workspace_bytes(B, C, H, W, self.mode)
        self.buf = torch.empty(self.nbytes, dtype=torch.uint8, device=f.device)
        with torch.cuda.device(f.device):
            _cabi.check(lib.rcb_corr_pack_fmaps(f.data_ptr(), self.buf.data_ptr(), self.nbytes, B, C, H, W, self.mode,
                                                _stream(f)), "rcb_corr_pack_fmaps")


class _State:
    """Everything one CorrBlock owns on the device."""

    def __init__(self, f1, f2, levels, radius, mode, pyr_dtype, packed=None):
        B, C, H, W = f1.shape if packed is None else packed.shape
        dev = f1.device if packed is None else packed.buf.device
        self.f1, self.f2, self.levels, self.radius, self.mode = f1, f2, levels, radius, mode
        self.pyr = _Pyramid(B, H, W, levels, dev, pyr_dtype)
        self.dpyr = None
        if packed is None:
            _build(f1, f2, levels, mode, self.pyr)
        else:
            with torch.cuda.device(dev):
                _cabi.check(_cabi.lib().rcb_corr_build_packed(packed.buf.data_ptr(), packed.nbytes, self.pyr.ptrs, B, C, H,
                                                              W, levels, mode, pyr_dtype, _stream(packed.buf)),
                            "rcb_corr_build_packed")
        with torch.cuda.device(dev):
            self.plan = _cabi.LookupPlan(self.pyr.ptrs, B, H, W, levels, radius, pyr_dtype)
        self.out_channels = levels * (2 * radius + 1) ** 2
        self.device_index = dev.index if dev.index is not None else torch.cuda.current_device()
        self._lookup = _cabi.lib().rcb_corr_lookup_planned

    def lookup(self, coords):
        B, two, H, W = coords.shape
        if (B, H, W) != (self.pyr.B, self.pyr.H, self.pyr.W) or two != 2:
            raise RuntimeError(f"coords shape {tuple(coords.shape)} does not match the feature maps")
        out = torch.empty((B, self.out_channels, H, W), dtype=torch.float32, device=coords.device)
        if torch.cuda.current_device() == self.device_index:  # kernels launch on the CURRENT device
            st = self._lookup(self.plan.ptr, coords.data_ptr(), out.data_ptr(), _stream(coords))
        else:
            with torch.cuda.device(coords.device):
                st = self._lookup(self.plan.ptr, coords.data_ptr(), out.data_ptr(), _stream(coords))
        if st != 0:
            _cabi.check(st, "rcb_corr_lookup_planned")
        return out


class PackedConvC1:
    """Weights of the motion encoder's first layer (``convc1``, reference core/update.py:136,182) in the operand
    layout of the fused lookup kernel (``rcb_corr_convc1_pack``).  Pack once per set of weights, reuse every
    iteration: ``block.lookup_conv(coords, packed)``."""

    def __init__(self, weight, bias, num_levels=4, radius=4):
        _check_cuda(weight, "weight")
        cout = weight.shape[0]
        cin = num_levels * (2 * radius + 1) ** 2
        w = weight.detach().float().reshape(cout, -1).contiguous()
        if w.shape[1] != cin:
            raise RuntimeError(f"convc1 weight has {w.shape[1]} input channels, the lookup produces {cin}")
        lib = _cabi.lib()
        n = lib.rcb_corr_convc1_pack_bytes(cout, num_levels, radius)
        if n == 0:
            raise RuntimeError("fused lookup + convc1 supports radius 3/4 and 16..256 output channels in steps of 16")
        self.cout, self.num_levels, self.radius = cout, num_levels, radius
        self.wpack = torch.empty(n, dtype=torch.uint8, device=weight.device)
        self.bias = None if bias is None else bias.detach().float().contiguous()
        with torch.cuda.device(weight.device):
            _cabi.check(lib.rcb_corr_convc1_pack(w.data_ptr(), self.wpack.data_ptr(), cout, num_levels, radius,
                                                 _stream(weight)), "rcb_corr_convc1_pack")


class CorrBlock:
    """All-pairs correlation pyramid with per-iteration window lookup (reference core/corr.py:12-94)."""

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4, mode=None, pyramid_dtype="f32"):
        self.num_levels = num_levels
        self.radius = radius
        if fmap1.shape != fmap2.shape or fmap1.dim() != 4:
            raise RuntimeError("fmap1 and fmap2 must be [N, C, H, W] tensors of equal shape")
        if not 1 <= num_levels <= _cabi.MAX_LEVELS or not 1 <= radius <= _cabi.MAX_RADIUS:
            raise RuntimeError(f"num_levels must be in 1..{_cabi.MAX_LEVELS} and radius in 1..{_cabi.MAX_RADIUS}")
        mode = _cabi.BUILD_MODES[mode or DEFAULT_MODE]
        pyr_dtype = {"f32": _cabi.F32, "f16": _cabi.F16}[pyramid_dtype]
        f1, f2 = _prep(fmap1, "fmap1"), _prep(fmap2, "fmap2")
        self._token = None
        self._empty_shape = None
        if f1.shape[0] == 0:  # an empty batch passes through the reference's torch ops; there is nothing to launch
            self._state = None
            self._empty_shape = tuple(f1.shape)
            self._empty_device = f1.device
            return
        self._state = _State(f1, f2, num_levels, radius, mode, pyr_dtype)
        if torch.is_grad_enabled() and (fmap1.requires_grad or fmap2.requires_grad):
            self._token = _BuildFn.apply(fmap1, fmap2, self._state)

    @classmethod
    def from_packed(cls, packed, num_levels=4, radius=4, pyramid_dtype="f32"):
        """The block of a frame pair whose features were packed with ``PackedFmaps`` (inference: no autograd)."""
        if not 1 <= num_levels <= _cabi.MAX_LEVELS or not 1 <= radius <= _cabi.MAX_RADIUS:
            raise RuntimeError(f"num_levels must be in 1..{_cabi.MAX_LEVELS} and radius in 1..{_cabi.MAX_RADIUS}")
        self = cls.__new__(cls)
        self.num_levels, self.radius = num_levels, radius
        pyr_dtype = {"f32": _cabi.F32, "f16": _cabi.F16}[pyramid_dtype]
        self._state = _State(None, None, num_levels, radius, packed.mode, pyr_dtype, packed=packed)
        self._token = None
        self._empty_shape = None
        return self

    @property
    def corr_pyramid(self):
        """List of num_levels tensors [N*H*W, 1, H_i, W_i] like the reference attribute (core/corr.py:38,46-54);
        nothing outside corr.py reads it in the reference.  Materialised on access from the tiled buffers."""
        if self._state is None:
            _, _, H, W = self._empty_shape
            return [torch.empty((0, 1, H >> l, W >> l), dtype=torch.float32, device=self._empty_device)
                    for l in range(self.num_levels)]
        return self._state.pyr.views()

    def _empty_out(self, coords, channels):
        _, _, H, W = self._empty_shape
        if tuple(coords.shape) != (0, 2, H, W):
            raise RuntimeError(f"coords shape {tuple(coords.shape)} does not match the feature maps")
        return torch.empty((0, channels, H, W), dtype=torch.float32, device=coords.device)

    def __call__(self, coords):
        if self._state is None:
            return self._empty_out(coords, self.num_levels * (2 * self.radius + 1) ** 2)
        if torch.is_grad_enabled() and (self._token is not None or coords.requires_grad):
            return _LookupFn.apply(coords, self._token, self._state)
        return self._state.lookup(_prep(coords, "coords"))

    def lookup_conv(self, coords, packed, relu=True):
        """relu(convc1(self(coords))) in one kernel (reference core/raft.py:219 + core/update.py:154,202) without
        materialising the correlation tensor; ``packed`` is a PackedConvC1.  Inference only (no autograd), fp32
        pyramids; fp16 tensor-core operands with fp32 accumulation.  Returns [N, cout, H, W] fp32."""
        st = self._state
        if st is None:
            return self._empty_out(coords, packed.cout)
        if st.pyr.dtype != _cabi.F32:
            raise RuntimeError("lookup_conv needs an fp32 pyramid")
        if (packed.num_levels, packed.radius) != (self.num_levels, self.radius):
            raise RuntimeError("PackedConvC1 was packed for a different num_levels / radius")
        c = _prep(coords, "coords")
        B, two, H, W = c.shape
        if (B, H, W) != (st.pyr.B, st.pyr.H, st.pyr.W) or two != 2:
            raise RuntimeError(f"coords shape {tuple(c.shape)} does not match the feature maps")
        out = torch.empty((B, packed.cout, H, W), dtype=torch.float32, device=c.device)
        with torch.cuda.device(c.device):
            _cabi.check(_cabi.lib().rcb_corr_lookup_convc1(
                st.plan.ptr, c.data_ptr(), packed.wpack.data_ptr(),
                None if packed.bias is None else packed.bias.data_ptr(), out.data_ptr(), packed.cout, int(relu),
                _stream(c)), "rcb_corr_lookup_convc1")
        return out

    @staticmethod
    def corr(fmap1, fmap2, mode=None):
        """[N,C,H,W] x2 -> [N,H,W,1,H,W] = fmap1^T fmap2 / sqrt(C) (reference core/corr.py:96-127)."""
        f1, f2 = _prep(fmap1, "fmap1"), _prep(fmap2, "fmap2")
        B, C, H, W = f1.shape
        if B == 0:
            return torch.empty((0, H, W, 1, H, W), dtype=torch.float32, device=f1.device)
        pyr = _Pyramid(B, H, W, 1, f1.device)
        _build(f1, f2, 1, _cabi.BUILD_MODES[mode or DEFAULT_MODE], pyr)
        return pyr.level(0).reshape(B, H, W, 1, H, W)


class _AltBuildFn(torch.autograd.Function):
    """Autograd anchor of AlternateCorrBlock: like _BuildFn, a scalar token makes this node's backward run once, after
    every call of the forward pass has accumulated its NHWC feature gradients in the block."""

    @staticmethod
    def forward(ctx, fmap1, fmap2, block):
        ctx.block = block
        return torch.zeros((), device=fmap1.device)

    @staticmethod
    def backward(ctx, _grad_token):
        blk = ctx.block
        B, C, H, W = blk._shape
        dev = blk._f1n.device
        if blk._df1n is None:
            z = torch.zeros((B, C, H, W), dtype=torch.float32, device=dev)
            return z, z.clone(), None
        # fmap2's pooled levels fold back onto level 0: avg_pool2d backward (core/corr.py:159-160), floor mode
        g2 = None
        for l in range(blk.num_levels - 1, -1, -1):
            cur = blk._df2n[l]
            if g2 is not None:
                up = g2.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2) * 0.25
                cur = cur.clone() if cur is not None else torch.zeros_like(blk._f2n[l])
                cur[:, :up.shape[1], :up.shape[2]] += up
            g2 = cur if cur is not None else torch.zeros_like(blk._f2n[l])
        df1 = blk._df1n.permute(0, 3, 1, 2).contiguous()
        df2 = g2.permute(0, 3, 1, 2).contiguous()
        blk._df1n, blk._df2n = None, [None] * blk.num_levels  # a second backward pass starts from zero again
        return df1, df2, None


class _AltLookupFn(torch.autograd.Function):
    """One AlternateCorrBlock call.  Backward = the extension's backward kernel per level (rcb_altcorr_backward,
    replacing alt_cuda_corr.backward, correlation.cpp:36-48) with the TRUE coordinate gradient the reference kernel
    leaves at zero (correlation_kernel.cu:307), which is what autograd through CorrBlock yields."""

    @staticmethod
    def forward(ctx, coords, token, block):
        ctx.block = block
        c = _prep(coords, "coords")
        ctx.save_for_backward(c)
        ctx.needs_f = token is not None and token.requires_grad
        return block._forward(c)

    @staticmethod
    def backward(ctx, grad_out):
        blk = ctx.block
        (c,) = ctx.saved_tensors
        B, C, H, W = blk._shape
        rd2 = (2 * blk.radius + 1) ** 2
        want_c = ctx.needs_input_grad[0]
        lib = _cabi.lib()
        scale = 1.0 / math.sqrt(C)
        cn = c.permute(0, 2, 3, 1)  # [B, H, W, 2] like core/corr.py:174
        dco = torch.zeros((B, H, W, 2), dtype=torch.float32, device=c.device) if want_c else None
        with torch.cuda.device(c.device):
            s = _stream(c)
            for l in range(blk.num_levels):
                cg = (grad_out[:, l * rd2:(l + 1) * rd2].float() * scale).reshape(B, 1, rd2, H, W).contiguous()
                cl = (cn / float(2 ** l)).reshape(B, 1, H, W, 2).contiguous()  # core/corr.py:187
                f2 = blk._f2n[l]
                g1, g2, gc = torch.empty_like(blk._f1n), torch.empty_like(f2), torch.empty_like(cl)
                _cabi.check(lib.rcb_altcorr_backward(blk._f1n.data_ptr(), f2.data_ptr(), cl.data_ptr(), cg.data_ptr(),
                                                     g1.data_ptr(), g2.data_ptr(), gc.data_ptr(), B, 1, H, W,
                                                     f2.shape[1], f2.shape[2], C, blk.radius, int(want_c), s),
                            "rcb_altcorr_backward")
                if ctx.needs_f:
                    blk._df1n = g1 if blk._df1n is None else blk._df1n.add_(g1)
                    blk._df2n[l] = g2 if blk._df2n[l] is None else blk._df2n[l].add_(g2)
                if want_c:
                    dco.add_(gc.reshape(B, H, W, 2), alpha=1.0 / float(2 ** l))
        dtoken = torch.zeros((), device=c.device) if ctx.needs_f else None
        return (dco.permute(0, 3, 1, 2).contiguous() if want_c else None), dtoken, None


class AlternateCorrBlock:
    """On-the-fly correlation (reference core/corr.py:130-198): no Q x Q volume is ever stored; every call
    evaluates the (2r+2)^2 tap dot products of each query at each level from pooled fmap2 features."""

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4):
        self.num_levels = num_levels
        self.radius = radius
        if fmap1.shape != fmap2.shape or fmap1.dim() != 4:
            raise RuntimeError("fmap1 and fmap2 must be [N, C, H, W] tensors of equal shape")
        if not 1 <= num_levels <= _cabi.MAX_LEVELS or not 1 <= radius <= _cabi.MAX_RADIUS:
            raise RuntimeError(f"num_levels must be in 1..{_cabi.MAX_LEVELS} and radius in 1..{_cabi.MAX_RADIUS}")
        f1, f2 = _prep(fmap1, "fmap1"), _prep(fmap2, "fmap2")
        B, C, H, W = f1.shape
        self._shape = (B, C, H, W)
        self._src = (f1, f2)
        self._token = None
        if B == 0:  # empty batch: nothing to prepare, calls return empty tensors
            self._f1n, self._f2n = None, []
            return
        self._f1n = torch.empty((B, H, W, C), dtype=torch.float32, device=f1.device)
        self._f2n = []
        h, w = H, W
        for _ in range(num_levels):
            if h < 1 or w < 1:
                raise RuntimeError("feature map too small for the requested number of levels")
            self._f2n.append(torch.empty((B, h, w, C), dtype=torch.float32, device=f1.device))
            h, w = h // 2, w // 2
        self._f2ptrs = _cabi.ptr_array([t.data_ptr() for t in self._f2n])
        with torch.cuda.device(f1.device):
            _cabi.check(_cabi.lib().rcb_altcorr_prepare(f1.data_ptr(), f2.data_ptr(), self._f1n.data_ptr(),
                                                        self._f2ptrs, B, C, H, W, num_levels, _stream(f1)),
                        "rcb_altcorr_prepare")
        # autograd (the reference has none on this path: alt_cuda_corr.backward is never called from Python): gradients
        # with respect to the feature maps accumulate here over the calls of one forward pass
        self._df1n, self._df2n = None, [None] * num_levels
        self._token = None
        if torch.is_grad_enabled() and (fmap1.requires_grad or fmap2.requires_grad):
            self._token = _AltBuildFn.apply(fmap1, fmap2, self)

    @property
    def pyramid(self):
        """[(fmap1_i, fmap2_i)] with num_levels+1 entries like the reference (core/corr.py:154-161).  Only
        pyramid[0][0] and pyramid[i][1] are ever used (core/corr.py:183-184); fmap2 entries are views of the
        NHWC buffers the kernels read, the unused pooled fmap1 entries are materialised on demand."""
        f1, f2 = self._src
        if self._shape[0] == 0:  # empty batch: plain pooled (empty) tensors
            out = [(f1, f2)]
            for _ in range(self.num_levels):
                f1, f2 = F.avg_pool2d(f1, 2, stride=2), F.avg_pool2d(f2, 2, stride=2)
                out.append((f1, f2))
            return out
        out = [(f1, self._f2n[0].permute(0, 3, 1, 2))]
        for i in range(1, self.num_levels + 1):
            f1 = F.avg_pool2d(f1, 2, stride=2)
            f2i = self._f2n[i].permute(0, 3, 1, 2) if i < self.num_levels else F.avg_pool2d(out[-1][1], 2, stride=2)
            out.append((f1, f2i))
        return out

    def __call__(self, coords):
        if self._shape[0] == 0:
            return self._forward(coords)
        if torch.is_grad_enabled() and (self._token is not None or coords.requires_grad):
            return _AltLookupFn.apply(coords, self._token, self)
        return self._forward(_prep(coords, "coords"))

    def _forward(self, c):
        B, C, H, W = self._shape
        if tuple(c.shape) != (B, 2, H, W):
            raise RuntimeError(f"coords shape {tuple(c.shape)} does not match the feature maps")
        rd = 2 * self.radius + 1
        out = torch.empty((B, self.num_levels * rd * rd, H, W), dtype=torch.float32, device=c.device)
        if B == 0:
            return out
        with torch.cuda.device(c.device):
            _cabi.check(_cabi.lib().rcb_altcorr_pyramid_forward(
                self._f1n.data_ptr(), self._f2ptrs, c.data_ptr(), out.data_ptr(), B, C, H, W, self.num_levels,
                self.radius, 1.0 / math.sqrt(C), _stream(c)), "rcb_altcorr_pyramid_forward")
        return out
