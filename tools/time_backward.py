"""Times the three kernels of the all-pairs backward separately through the C ABI (CUDA events):
lookup_backward (one GRU iteration), pool_backward, contract_backward.   python tools/time_backward.py [--config cfg5]"""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CONFIGS, SEED  # noqa: E402
from raft_optical_flow_b200 import CorrBlock, _cabi  # noqa: E402
from raft_optical_flow_b200.corr import _Pyramid  # noqa: E402
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg5")
a = ap.parse_args()
B, C, H, W, r, L, iters, _ = CONFIGS[a.config]
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(SEED)
f1 = (0.75 * torch.randn(B, C, H, W, generator=g)).to(dev)
f2 = (0.75 * torch.randn(B, C, H, W, generator=g)).to(dev)
ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
coords = (torch.stack([xs, ys]).float()[None] + 4.0 * torch.randn(B, 2, H, W, generator=g)).to(dev).contiguous()
rd = 2 * r + 1
go = torch.randn(B, L * rd * rd, H, W, device=dev)
blk = CorrBlock(f1, f2, num_levels=L, radius=r)
st = blk._state
dpyr = _Pyramid(B, H, W, L, dev, _cabi.F32, zero=True)
dco = torch.empty_like(coords)
df1, df2 = torch.empty_like(f1), torch.empty_like(f2)
lib = _cabi.lib()
s = torch.cuda.current_stream().cuda_stream
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
t_lb = timed(lambda: _cabi.check(lib.rcb_corr_lookup_backward(st.pyr.ptrs, coords.data_ptr(), go.data_ptr(), dpyr.ptrs, dco.data_ptr(), B, H, W, L, r, 0, s), "lb"))
t_lb_nc = timed(lambda: _cabi.check(lib.rcb_corr_lookup_backward(st.pyr.ptrs, coords.data_ptr(), go.data_ptr(), dpyr.ptrs, None, B, H, W, L, r, 0, s), "lb"))
t_pb = timed(lambda: _cabi.check(lib.rcb_corr_pool_backward(dpyr.ptrs, B, H, W, L, s), "pb"))
t_cb = timed(lambda: _cabi.check(lib.rcb_corr_contract_backward(f1.data_ptr(), f2.data_ptr(), dpyr.bufs[0].data_ptr(), df1.data_ptr(), df2.data_ptr(), B, C, H, W, s), "cb"))
nws = lib.rcb_corr_contract_backward_tc_workspace_bytes(B, C, H, W)
ws = torch.empty(max(nws, 256), dtype=torch.uint8, device=dev)
df1t, df2t = torch.empty_like(f1), torch.empty_like(f2)
t_cbt = timed(lambda: _cabi.check(lib.rcb_corr_contract_backward_tc(f1.data_ptr(), f2.data_ptr(), dpyr.bufs[0].data_ptr(), df1t.data_ptr(), df2t.data_ptr(), B, C, H, W, ws.data_ptr(), nws, s), "cbt"))
e1 = ((df1t - df1).abs().max() / df1.abs().max()).item(); e2 = ((df2t - df2).abs().max() / df2.abs().max()).item()
t_zero = timed(lambda: dpyr.zero_())
Q = H * W
print(f"{a.config}: lookup_backward {t_lb:.0f} us (without dcoords {t_lb_nc:.0f}) x{iters} iterations, pool_backward {t_pb:.0f} us, "
      f"contract_backward {t_cb:.0f} us ({4.0 * B * Q * Q * C / t_cb / 1e6:.1f} TFLOP/s fp32), tensor-core form {t_cbt:.0f} us "
      f"({4.0 * B * Q * Q * C / t_cbt / 1e6:.1f} TFLOP/s algorithmic, workspace {nws / 1e9:.2f} GB, rel diff vs fp32 tiles dF1 {e1:.1e} dF2 {e2:.1e}), "
      f"zero dpyr {t_zero:.0f} us")
