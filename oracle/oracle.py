"""numpy front-end of the CPU oracle (oracle/corr_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / ``--impl reference`` legs of bench.py.  The product package
(raft_optical_flow_b200) never imports this module.

The two classes below restate the *composition* the reference performs around the
primitive ops (reference paths relative to its checkout):

* ``OracleCorrBlock``          -- core/corr.py:12-127   (CorrBlock)
* ``OracleAlternateCorrBlock`` -- core/corr.py:130-198  (AlternateCorrBlock -> alt_cuda_corr.forward)

Parity pinning: tests/test_oracle_golden.py checks them against fixtures generated
from the reference itself (tests/golden/make_golden.py).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_SRC_PATH = os.path.join(_HERE, "corr_oracle.c")

_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_i32p = ctypes.POINTER(ctypes.c_int)


def build(force=False):
    """Compile corr_oracle.c with the system gcc (see oracle/Makefile)."""
    if not force and os.path.exists(_LIB_PATH) and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(_SRC_PATH):
        return _LIB_PATH
    os.makedirs(os.path.dirname(_LIB_PATH), exist_ok=True)
    base = ["-O3", "-mavx2", "-ffp-contract=off", "-fPIC", "-shared", "-fvisibility=hidden",
            "-o", _LIB_PATH, _SRC_PATH, "-lm"]
    last = None
    for cc in ("/usr/bin/gcc", "gcc"):
        for omp in (["-fopenmp"], []):
            try:
                subprocess.run([cc] + omp + base, check=True, capture_output=True)
                return _LIB_PATH
            except (OSError, subprocess.CalledProcessError) as e:  # try the next recipe
                last = e
    raise RuntimeError(f"could not build the CPU oracle: {last}")


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        try:
            _lib = ctypes.CDLL(_LIB_PATH)
        except OSError:
            build(force=True)
            _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_num_threads.restype = ctypes.c_int
    return _lib


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    lib().orc_set_num_threads(ctypes.c_int(int(n)))


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(_f32p)


def pyramid_shapes(H, W, L):
    """floor-mode halving, core/corr.py:52-54."""
    hs, ws = [H], [W]
    for _ in range(L - 1):
        hs.append(hs[-1] // 2)
        ws.append(ws[-1] // 2)
    return hs, ws


def corr_volume(fmap1, fmap2, acc64=True):
    """core/corr.py:96-127 -> [B, Q, Q] (the reference views it as [B,H,W,1,H,W])."""
    f1, f2 = _c(fmap1), _c(fmap2)
    B, C, H, W = f1.shape
    Q = H * W
    vol = np.empty((B, Q, Q), dtype=np.float32)
    lib().orc_corr_volume(_p(f1), _p(f2), B, C, H, W, _p(vol), int(bool(acc64)))
    return vol


def avg_pool2(x):
    """F.avg_pool2d(x, 2, stride=2) over the last two dims (core/corr.py:53,159-160)."""
    x = _c(x)
    H, W = x.shape[-2:]
    n = int(np.prod(x.shape[:-2], dtype=np.int64))
    out = np.empty(x.shape[:-2] + (H // 2, W // 2), dtype=np.float32)
    lib().orc_avg_pool2(_p(x), ctypes.c_long(n), H, W, _p(out))
    return out


def lookup(pyramid, coords, radius, roundtrip=True):
    """core/corr.py:56-94.  pyramid[i]: [B*H*W, H_i, W_i]; coords [B,2,H,W] -> [B, L*rd*rd, H, W]."""
    coords = _c(coords)
    B, _, H, W = coords.shape
    L = len(pyramid)
    pyr = [_c(p).reshape(B * H * W, p.shape[-2], p.shape[-1]) for p in pyramid]
    Hs = (ctypes.c_int * L)(*[p.shape[-2] for p in pyr])
    Ws = (ctypes.c_int * L)(*[p.shape[-1] for p in pyr])
    ptrs = (_f32p * L)(*[_p(p) for p in pyr])
    rd = 2 * radius + 1
    out = np.empty((B, L * rd * rd, H, W), dtype=np.float32)
    lib().orc_lookup(ptrs, Hs, Ws, _p(coords), B, H, W, L, radius, _p(out), int(bool(roundtrip)))
    return out


def altcorr_forward(fmap1, fmap2, coords, radius):
    """alt_cuda_corr.forward, correlation_kernel.cu:18-119,260-286.
    fmap1 [B,H1,W1,C], fmap2 [B,H2,W2,C], coords [B,N,H1,W1,2] -> corr [B,N,rd*rd,H1,W1] (unscaled)."""
    f1, f2, co = _c(fmap1), _c(fmap2), _c(coords)
    B, H1, W1, C = f1.shape
    _, H2, W2, _ = f2.shape
    N = co.shape[1]
    rd = 2 * radius + 1
    out = np.empty((B, N, rd * rd, H1, W1), dtype=np.float32)
    lib().orc_altcorr_forward(_p(f1), _p(f2), _p(co), B, N, H1, W1, H2, W2, C, radius, _p(out))
    return out


def altcorr_backward(fmap1, fmap2, coords, corr_grad, radius, true_coords_grad=False):
    """alt_cuda_corr.backward, correlation_kernel.cu:122-256,288-324.
    Returns (fmap1_grad, fmap2_grad, coords_grad); coords_grad is zeros like the reference unless
    ``true_coords_grad`` asks for the real derivative."""
    f1, f2, co, cg = _c(fmap1), _c(fmap2), _c(coords), _c(corr_grad)
    B, H1, W1, C = f1.shape
    _, H2, W2, _ = f2.shape
    N = co.shape[1]
    g1 = np.empty_like(f1)
    g2 = np.empty_like(f2)
    gc = np.empty_like(co)
    lib().orc_altcorr_backward(_p(f1), _p(f2), _p(co), _p(cg), B, N, H1, W1, H2, W2, C, radius,
                               _p(g1), _p(g2), _p(gc), int(bool(true_coords_grad)))
    return g1, g2, gc


class OracleCorrBlock:
    """core/corr.py:12-94: volume, (L-1) poolings, per-call window lookup."""

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4, acc64=True):
        self.num_levels = num_levels
        self.radius = radius
        self.fmap1, self.fmap2 = _c(fmap1), _c(fmap2)
        B, C, H, W = self.fmap1.shape
        vol = corr_volume(self.fmap1, self.fmap2, acc64=acc64).reshape(B * H * W, H, W)
        self.corr_pyramid = [vol]
        for _ in range(num_levels - 1):
            vol = avg_pool2(vol)
            self.corr_pyramid.append(vol)

    def __call__(self, coords, roundtrip=True):
        return lookup(self.corr_pyramid, coords, self.radius, roundtrip=roundtrip)

    def backward(self, coords, grad_out):
        """What autograd yields through CorrBlock for (fmap1, fmap2, coords) (train.py:212)."""
        coords, grad_out = _c(coords), _c(grad_out)
        B, C, H, W = self.fmap1.shape
        L = self.num_levels
        Hs, Ws = pyramid_shapes(H, W, L)
        scratch = [np.zeros(B * H * W * Hs[i] * Ws[i], dtype=np.float64) for i in range(L)]
        df1 = np.empty_like(self.fmap1)
        df2 = np.empty_like(self.fmap2)
        dco = np.empty_like(coords)
        pptr = (_f32p * L)(*[_p(p) for p in self.corr_pyramid])
        sptr = (_f64p * L)(*[s.ctypes.data_as(_f64p) for s in scratch])
        lib().orc_corrblock_backward(_p(self.fmap1), _p(self.fmap2), pptr, (ctypes.c_int * L)(*Hs),
                                     (ctypes.c_int * L)(*Ws), _p(coords), _p(grad_out), B, C, H, W, L,
                                     self.radius, sptr, _p(df1), _p(df2), _p(dco))
        return df1, df2, dco


class OracleAlternateCorrBlock:
    """core/corr.py:130-198: pooled feature pyramid, per-level alt_cuda_corr.forward, stack, / sqrt(C)."""

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4):
        self.num_levels = num_levels
        self.radius = radius
        f1, f2 = _c(fmap1), _c(fmap2)
        self.pyramid = [(f1, f2)]
        for _ in range(num_levels):  # corr.py:157-161 builds L+1 entries
            f1, f2 = avg_pool2(f1), avg_pool2(f2)
            self.pyramid.append((f1, f2))

    def __call__(self, coords):
        coords = _c(coords).transpose(0, 2, 3, 1)  # corr.py:174
        B, H, W, _ = coords.shape
        dim = self.pyramid[0][0].shape[1]
        f1 = np.ascontiguousarray(self.pyramid[0][0].transpose(0, 2, 3, 1))  # corr.py:183
        outs = []
        for i in range(self.num_levels):
            f2 = np.ascontiguousarray(self.pyramid[i][1].transpose(0, 2, 3, 1))  # corr.py:184
            ci = np.ascontiguousarray((coords / np.float32(2 ** i)).reshape(B, 1, H, W, 2))  # corr.py:187
            outs.append(altcorr_forward(f1, f2, ci, self.radius)[:, 0])  # corr.py:190-191
        corr = np.stack(outs, axis=1).reshape(B, -1, H, W)  # corr.py:194-195
        return corr / np.sqrt(np.float32(dim))  # corr.py:198


# ---- next row: convex upsampling ----------------------------------------------------------
def _unfold3x3(x8):
    """F.unfold(x, [3, 3], padding=1) of core/raft.py:132 as an array [N, 2, 9, H, W]: neighbour k = ky*3 + kx is
    x[.., h + ky - 1, w + kx - 1], zero outside."""
    N, C, H, W = x8.shape
    pad = np.zeros((N, C, H + 2, W + 2), dtype=x8.dtype)
    pad[:, :, 1:-1, 1:-1] = x8
    return np.stack([pad[:, :, ky:ky + H, kx:kx + W] for ky in range(3) for kx in range(3)], axis=2)


def _softmax9(mask):
    N, _, H, W = mask.shape
    m = mask.reshape(N, 9, 8, 8, H, W).astype(np.float64)  # core/raft.py:128: view(N, 1, 9, 8, 8, H, W)
    m = np.exp(m - m.max(axis=1, keepdims=True))
    return m / m.sum(axis=1, keepdims=True)                # core/raft.py:130: softmax over the 9 neighbours


def upsample_flow(flow, mask):
    """RAFT.upsample_flow (reference core/raft.py:112-142) in numpy (float64 accumulation):
    flow [N,2,H,W], mask [N,576,H,W] -> [N,2,8H,8W]."""
    N, _, H, W = flow.shape
    p = _softmax9(mask)                                     # [N, 9, 8, 8, H, W]
    nb = _unfold3x3(8.0 * flow.astype(np.float64))          # [N, 2, 9, H, W]   (core/raft.py:132-133)
    up = np.einsum("nkijhw,nckhw->ncijhw", p, nb)           # core/raft.py:136: sum(mask * up_flow, dim=2)
    return up.transpose(0, 1, 4, 2, 5, 3).reshape(N, 2, 8 * H, 8 * W).astype(np.float32)  # core/raft.py:138-140


def upsample_flow_backward(flow, mask, grad_out):
    """Gradients of upsample_flow w.r.t. flow and mask for the cotangent grad_out [N,2,8H,8W] (what autograd yields
    for core/raft.py:112-142)."""
    N, _, H, W = flow.shape
    p = _softmax9(mask)
    nb = _unfold3x3(8.0 * flow.astype(np.float64))
    g = grad_out.astype(np.float64).reshape(N, 2, H, 8, W, 8).transpose(0, 1, 3, 5, 2, 4)  # [N, 2, i, j, H, W]
    gn = np.einsum("ncijhw,nckhw->nkijhw", g, nb)           # d out / d p_k contracted with the cotangent
    dmask = p * (gn - (p * gn).sum(axis=1, keepdims=True))  # softmax Jacobian
    dnb = np.einsum("nkijhw,ncijhw->nckhw", p, g)           # [N, 2, 9, H, W]
    dpad = np.zeros((N, 2, H + 2, W + 2))
    for ky in range(3):
        for kx in range(3):
            dpad[:, :, ky:ky + H, kx:kx + W] += dnb[:, :, ky * 3 + kx]
    dflow = 8.0 * dpad[:, :, 1:-1, 1:-1]
    return dflow.astype(np.float32), dmask.reshape(N, 576, H, W).astype(np.float32)


def convc1_relu(corr, weight, bias=None, relu=True):
    """First layer of the motion encoder on the lookup output: relu(convc1(corr)) with convc1 a 1x1 convolution
    (reference core/update.py:136,154 SmallMotionEncoder, :182,202 BasicMotionEncoder), float64 accumulation.
    corr [N,K,H,W], weight [Cout,K] or [Cout,K,1,1], bias [Cout] -> [N,Cout,H,W]."""
    w = np.asarray(weight, dtype=np.float64).reshape(weight.shape[0], -1)
    out = np.einsum("nkhw,ck->nchw", np.asarray(corr, dtype=np.float64), w)
    if bias is not None:
        out = out + np.asarray(bias, dtype=np.float64).reshape(1, -1, 1, 1)
    if relu:
        out = np.maximum(out, 0.0)
    return out.astype(np.float32)
