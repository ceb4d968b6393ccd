"""Convex upsampling of the 1/8-resolution flow -- the next row of the scope table after the correlation path
(SURVEY 8f, f3).  ``upsample_flow(flow, mask)`` computes what ``RAFT.upsample_flow`` does (reference
core/raft.py:112-142) in one fused sm_100a kernel; ``patch_raft(raft_module)`` installs it together with the
correlation blocks at the names the reference looks up.  CUDA tensors only, like the rest of the package.
"""
import torch

from . import _cabi

__all__ = ["upsample_flow", "patch_raft"]


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


class _UpsampleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flow, mask):
        for t, name in ((flow, "flow"), (mask, "mask")):
            if not t.is_cuda:
                raise RuntimeError(f"{name} must be a CUDA tensor")
        f, m = flow.detach().float().contiguous(), mask.detach().float().contiguous()
        N, two, H, W = f.shape
        if two != 2 or tuple(m.shape) != (N, 576, H, W):
            raise RuntimeError(f"expected flow [N,2,H,W] and mask [N,576,H,W], got {tuple(f.shape)} / {tuple(m.shape)}")
        out = torch.empty((N, 2, 8 * H, 8 * W), dtype=torch.float32, device=f.device)
        with torch.cuda.device(f.device):
            _cabi.check(_cabi.lib().rcb_upsample_flow(f.data_ptr(), m.data_ptr(), out.data_ptr(), N, H, W, _stream(f)),
                        "rcb_upsample_flow")
        ctx.save_for_backward(f, m)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        f, m = ctx.saved_tensors
        N, _, H, W = f.shape
        g = grad_out.float().contiguous()
        dflow, dmask = torch.empty_like(f), torch.empty_like(m)
        lib = _cabi.lib()
        nws = lib.rcb_upsample_flow_backward_workspace_bytes(N, H, W)
        ws = torch.empty(nws, dtype=torch.uint8, device=f.device)
        with torch.cuda.device(f.device):
            _cabi.check(lib.rcb_upsample_flow_backward(f.data_ptr(), m.data_ptr(), g.data_ptr(), dflow.data_ptr(),
                                                       dmask.data_ptr(), ws.data_ptr(), nws, N, H, W, _stream(f)),
                        "rcb_upsample_flow_backward")
        return dflow, dmask


def upsample_flow(flow, mask):
    """[N,2,H,W] flow, [N,576,H,W] mask -> [N,2,8H,8W] (reference core/raft.py:112-142)."""
    return _UpsampleFn.apply(flow, mask)


def patch_raft(raft_module, fuse_motion_encoder=False):
    """Installs CorrBlock / AlternateCorrBlock at the module globals core/raft.py:187,189 looks up and the fused
    convex upsampling as RAFT.upsample_flow (core/raft.py:112,240).  Returns the replaced objects.
    fuse_motion_encoder=True additionally defers every lookup into the motion encoder's first layer (fused.py:
    one kernel for corr_fn(coords1) + relu(convc1(.)), inference only); undo that part with the function stored as
    ``raft_module._rcb_undo_fused``."""
    from .corr import AlternateCorrBlock, CorrBlock
    old = (raft_module.CorrBlock, raft_module.AlternateCorrBlock, raft_module.RAFT.upsample_flow)
    raft_module.CorrBlock, raft_module.AlternateCorrBlock = CorrBlock, AlternateCorrBlock
    raft_module.RAFT.upsample_flow = lambda self, flow, mask: upsample_flow(flow, mask)
    if fuse_motion_encoder:
        from .fused import install_fused_motion_encoder
        raft_module._rcb_undo_fused = install_fused_motion_encoder(raft_module)
    return old
