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


class _GraphedInference:
    """RAFT.forward(test_mode=True) under no_grad as ONE CUDA graph per (model, input shape, iters): the reference's
    GRU loop (core/raft.py:214-243) issues ~150 small kernels per iteration from Python, and at batch 1 the host
    cannot feed the GPU -- the lookup then runs at a quarter of its roofline behind launch latency.  The first call
    with a new key runs the model eagerly twice (cuDNN autotuning, allocator warm-up) and captures the third run;
    later calls copy the frames into the graph's input buffers, replay, and return copies of the outputs.  Every
    kernel of this package is capture-safe (no allocation, no synchronisation, plans are host-side data)."""

    def __init__(self, original):
        self.original = original
        self.cache = {}

    def __call__(self, module, image1, image2, iters=12, flow_init=None, upsample=True, test_mode=False):
        if (not test_mode or torch.is_grad_enabled() or flow_init is not None or not image1.is_cuda
                or torch.cuda.is_current_stream_capturing()):
            return self.original(module, image1, image2, iters=iters, flow_init=flow_init, upsample=upsample,
                                 test_mode=test_mode)
        key = (id(module), tuple(image1.shape), image1.dtype, image1.device, int(iters), module.training)
        entry = self.cache.get(key)
        if entry is None:
            in1, in2 = image1.clone(), image2.clone()
            side = torch.cuda.Stream(image1.device)
            side.wait_stream(torch.cuda.current_stream(image1.device))
            with torch.cuda.stream(side):
                for _ in range(2):
                    self.original(module, in1, in2, iters=iters, test_mode=True)
            torch.cuda.current_stream(image1.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self.original(module, in1, in2, iters=iters, test_mode=True)
            entry = (graph, in1, in2, out)
            self.cache[key] = entry
        graph, in1, in2, out = entry
        in1.copy_(image1)
        in2.copy_(image2)
        graph.replay()
        return tuple(o.clone() for o in out)


def patch_raft(raft_module, fuse_motion_encoder=False, cuda_graph=False):
    """Installs CorrBlock / AlternateCorrBlock at the module globals core/raft.py:187,189 looks up and the fused
    convex upsampling as RAFT.upsample_flow (core/raft.py:112,240).  Returns the replaced objects.
    fuse_motion_encoder=True additionally defers every lookup into the motion encoder's first layer (fused.py:
    one kernel for corr_fn(coords1) + relu(convc1(.)), inference only); undo that part with the function stored as
    ``raft_module._rcb_undo_fused``.
    cuda_graph=True additionally replays no-grad ``test_mode`` forwards as one CUDA graph per input shape
    (``_GraphedInference``; the parameters are read in place, so weight updates are seen; undo with
    ``raft_module._rcb_undo_graph()``)."""
    from .corr import AlternateCorrBlock, CorrBlock
    old = (raft_module.CorrBlock, raft_module.AlternateCorrBlock, raft_module.RAFT.upsample_flow)
    raft_module.CorrBlock, raft_module.AlternateCorrBlock = CorrBlock, AlternateCorrBlock
    raft_module.RAFT.upsample_flow = lambda self, flow, mask: upsample_flow(flow, mask)
    if fuse_motion_encoder:
        from .fused import install_fused_motion_encoder
        raft_module._rcb_undo_fused = install_fused_motion_encoder(raft_module)
    if cuda_graph:
        original = raft_module.RAFT.forward
        graphed = _GraphedInference(original)

        def forward(self, image1, image2, iters=12, flow_init=None, upsample=True, test_mode=False):
            return graphed(self, image1, image2, iters=iters, flow_init=flow_init, upsample=upsample, test_mode=test_mode)

        raft_module.RAFT.forward = forward

        def undo():
            raft_module.RAFT.forward = original
            graphed.cache.clear()
        raft_module._rcb_undo_graph = undo
    return old
