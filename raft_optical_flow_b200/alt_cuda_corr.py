"""Drop-in replacement for the reference's ``alt_cuda_corr`` extension module
(alt_cuda_corr/correlation.cpp:51-54: ``forward`` and ``backward``).

Same argument order, shapes, return type (a list of tensors) and error behaviour as the pybind11
module: non-CUDA or non-contiguous inputs raise ``RuntimeError("<name> must be a CUDA tensor")`` /
``RuntimeError("<name> must be contiguous")`` (correlation.cpp:19-21).  Outputs are freshly allocated and
owned by the caller; inputs are never modified.
"""
import torch

from . import _cabi

__all__ = ["forward", "backward"]


def _check_input(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if t.dtype != torch.float32:
        # the reference accessor is packed_accessor32<float,...> (correlation_kernel.cu:279-282)
        raise RuntimeError(f"expected scalar type Float but found {t.dtype} for {name}")


def _dims(fmap1, fmap2, coords):
    if fmap1.dim() != 4 or fmap2.dim() != 4 or coords.dim() != 5 or coords.shape[-1] != 2:
        raise RuntimeError("expected fmap1 [B,H1,W1,C], fmap2 [B,H2,W2,C], coords [B,N,H1,W1,2]")
    B, H1, W1, C = fmap1.shape
    _, H2, W2, C2 = fmap2.shape
    N = coords.shape[1]
    if fmap2.shape[0] != B or C2 != C or tuple(coords.shape) != (B, N, H1, W1, 2):
        raise RuntimeError("fmap1 / fmap2 / coords shapes are inconsistent")
    return B, N, H1, W1, H2, W2, C


def forward(fmap1, fmap2, coords, radius):
    """fmap1 [B,H1,W1,C], fmap2 [B,H2,W2,C], coords [B,N,H1,W1,2] -> [corr [B,N,(2r+1)^2,H1,W1]] (unscaled)."""
    _check_input(fmap1, "fmap1")
    _check_input(fmap2, "fmap2")
    _check_input(coords, "coords")
    B, N, H1, W1, H2, W2, C = _dims(fmap1, fmap2, coords)
    rd = 2 * int(radius) + 1
    corr = torch.empty((B, N, rd * rd, H1, W1), dtype=torch.float32, device=fmap1.device)
    with torch.cuda.device(fmap1.device):
        s = torch.cuda.current_stream(fmap1.device).cuda_stream
        _cabi.check(_cabi.lib().rcb_altcorr_forward(fmap1.data_ptr(), fmap2.data_ptr(), coords.data_ptr(),
                                                    corr.data_ptr(), B, N, H1, W1, H2, W2, C, int(radius), s),
                    "alt_cuda_corr.forward")
    return [corr]


def backward(fmap1, fmap2, coords, corr_grad, radius, true_coords_grad=False):
    """-> [fmap1_grad, fmap2_grad, coords_grad].  coords_grad is all zeros exactly like the reference kernel
    (correlation_kernel.cu:307,323) unless ``true_coords_grad=True`` asks for the real derivative."""
    _check_input(fmap1, "fmap1")
    _check_input(fmap2, "fmap2")
    _check_input(coords, "coords")
    _check_input(corr_grad, "corr_grad")
    B, N, H1, W1, H2, W2, C = _dims(fmap1, fmap2, coords)
    rd = 2 * int(radius) + 1
    if tuple(corr_grad.shape) != (B, N, rd * rd, H1, W1):
        raise RuntimeError("corr_grad must be [B,N,(2r+1)^2,H1,W1]")
    g1 = torch.empty_like(fmap1)
    g2 = torch.empty_like(fmap2)
    gc = torch.empty_like(coords)
    with torch.cuda.device(fmap1.device):
        s = torch.cuda.current_stream(fmap1.device).cuda_stream
        _cabi.check(_cabi.lib().rcb_altcorr_backward(fmap1.data_ptr(), fmap2.data_ptr(), coords.data_ptr(),
                                                     corr_grad.data_ptr(), g1.data_ptr(), g2.data_ptr(),
                                                     gc.data_ptr(), B, N, H1, W1, H2, W2, C, int(radius),
                                                     int(bool(true_coords_grad)), s), "alt_cuda_corr.backward")
    return [g1, g2, gc]
