"""Lookup fused with the motion encoder's first layer inside the unmodified reference model (scope table 8f, f1).

The reference computes ``corr = corr_fn(coords1)`` (core/raft.py:219) and, as the first thing inside the update
block, ``cor = F.relu(self.convc1(corr))`` (core/update.py:154 SmallMotionEncoder, :202 BasicMotionEncoder).
``LazyCorrBlock`` defers the lookup: its call returns a handle, and the patched encoder ``forward`` resolves the handle
with one ``CorrBlock.lookup_conv`` launch, so the correlation tensor never exists.  Inference only -- whenever autograd
is recording, the handle is never created and the reference's own two steps run on the materialised tensor.
Accuracy class: the fused layer feeds fp16 operands to the tensor cores (~3e-4 of max-abs against an fp32 convolution,
the class of the TF32 convolution cuDNN runs for the reference by default); ``patch_raft(fuse_motion_encoder=True)``
enables it for ALL no-grad inference of the patched model.
"""
import sys

import torch
import torch.nn.functional as F

from .corr import CorrBlock, PackedConvC1

__all__ = ["LazyCorrBlock", "install_fused_motion_encoder"]


class _LazyCorr:
    """corr_fn(coords) that has not been evaluated yet."""
    __slots__ = ("block", "coords")

    def __init__(self, block, coords):
        self.block, self.coords = block, coords

    def materialise(self):
        return CorrBlock.__call__(self.block, self.coords)


class LazyCorrBlock(CorrBlock):
    def __call__(self, coords):
        if torch.is_grad_enabled() or self._state.pyr.dtype != 0:
            return super().__call__(coords)
        return _LazyCorr(self, coords)


_PACKED = {}      # (device, weight ptr, weight version, bias ptr, bias version, dtype, levels, radius) -> PackedConvC1
_PACKED_MAX = 32  # a handful of models x devices; oldest entries are dropped


def _packed_for(conv, block):
    """Packed convc1 weights, cached at MODULE level and keyed by the parameter storage itself: nn.DataParallel
    (the reference's demo.py / evaluate.py) rebuilds its replicas on every forward, so a cache hung on the conv module
    would never hit and the weights would be re-packed on every GRU iteration; replicas of one device share the
    broadcast parameter memory only within a forward, so the key also carries the version counters."""
    w, b = conv.weight, conv.bias
    key = (w.device, w.data_ptr(), w._version, None if b is None else b.data_ptr(), None if b is None else b._version,
           w.dtype, block.num_levels, block.radius)
    packed = _PACKED.get(key)
    if packed is None:
        packed = PackedConvC1(w, b, block.num_levels, block.radius)
        if len(_PACKED) >= _PACKED_MAX:
            _PACKED.pop(next(iter(_PACKED)))
        _PACKED[key] = packed
    return packed


def _supported(conv):
    return (conv.kernel_size == (1, 1) and conv.stride == (1, 1) and conv.padding == (0, 0) and conv.groups == 1
            and conv.out_channels % 16 == 0 and 16 <= conv.out_channels <= 256)


def _encoder_forward(original):
    def forward(self, flow, corr):
        if not isinstance(corr, _LazyCorr):
            return original(self, flow, corr)
        if torch.is_grad_enabled() or corr.block.radius not in (3, 4) or not _supported(self.convc1):
            return original(self, flow, corr.materialise())
        cor = corr.block.lookup_conv(corr.coords, _packed_for(self.convc1, corr.block), relu=True)
        # the rest of the encoder as the reference has it (core/update.py:155-166 small, :203-216 basic)
        if hasattr(self, "convc2"):
            cor = F.relu(self.convc2(cor))
        flo = F.relu(self.convf1(flow))
        flo = F.relu(self.convf2(flo))
        out = F.relu(self.conv(torch.cat([cor, flo], dim=1)))
        return torch.cat([out, flow], dim=1)
    forward._rcb_original = original
    return forward


def install_fused_motion_encoder(raft_module):
    """Installs LazyCorrBlock as raft_module.CorrBlock and wraps the forward of the reference's motion encoders (the
    classes of the module core/raft.py:9 imports its update blocks from).  Returns a function that undoes both."""
    upd = sys.modules[raft_module.BasicUpdateBlock.__module__]
    encoders = [upd.BasicMotionEncoder, upd.SmallMotionEncoder]
    old_block = raft_module.CorrBlock
    for cls in encoders:
        if not hasattr(cls.forward, "_rcb_original"):
            cls.forward = _encoder_forward(cls.forward)
    raft_module.CorrBlock = LazyCorrBlock

    def undo():
        raft_module.CorrBlock = old_block
        for cls in encoders:
            if hasattr(cls.forward, "_rcb_original"):
                cls.forward = cls.forward._rcb_original
    return undo
