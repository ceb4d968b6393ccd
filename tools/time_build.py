"""Times rcb_corr_build (pack + main kernel) through the C ABI with preallocated buffers (no Python/torch
allocation inside the timed loop), CUDA events on the launching stream.
    python tools/time_build.py [--config cfg2] [--mode bf16x3] [--reps 20]"""
import argparse, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CONFIGS, SEED  # noqa: E402
from raft_optical_flow_b200 import _cabi  # noqa: E402
from raft_optical_flow_b200.corr import _Pyramid  # noqa: E402
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg2")
ap.add_argument("--mode", default="bf16x3")
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
B, C, H, W, r, L, iters, _ = CONFIGS[a.config]
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(SEED)
f1 = (0.75 * torch.randn(B, C, H, W, generator=g)).to(dev)
f2 = (0.75 * torch.randn(B, C, H, W, generator=g)).to(dev)
lib = _cabi.lib()
mode = _cabi.BUILD_MODES[a.mode]
pyr = _Pyramid(B, H, W, L, dev)
nws = lib.rcb_corr_build_workspace_bytes(B, C, H, W, mode)
ws = torch.empty(max(nws, 16), dtype=torch.uint8, device=dev)
s = torch.cuda.current_stream().cuda_stream
def call():
    _cabi.check(lib.rcb_corr_build(f1.data_ptr(), f2.data_ptr(), pyr.ptrs, B, C, H, W, L, mode, 0, ws.data_ptr(), nws, s), "build")
for _ in range(3):
    call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(a.reps):
    call()
e1.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
if os.environ.get("RCB_TC_PROF") == "1":
    prof = torch.zeros(16 * 148, dtype=torch.int64, device=dev)
    os.environ["RCB_TC_PROF_PTR"] = hex(prof.data_ptr())
    call()
    torch.cuda.synchronize()
    del os.environ["RCB_TC_PROF_PTR"]
    pr = prof.view(148, 16).double()
    names = ["prod_total", "prod_wait_b_empty", "unused", "mma_total", "mma_wait_b_full", "mma_wait_acc_empty",
             "mma_wait_a_full", "epi_total", "epi_wait_acc_full", "epi_wait_store_read", "epi_tmem_ld", "tiles", "pool_total", "pool_wait_acc_full"]
    print("per-CTA mean cycles:", {n: int(pr[:, i].mean().item()) for i, n in enumerate(names)})
print(f"{a.config} {a.mode} skip={os.environ.get('RCB_TC_DEBUG_SKIP','0')} nstage={os.environ.get('RCB_TC_NSTAGE','-')} "
      f"build {e0.elapsed_time(e1) / a.reps * 1e3:.1f} us (host enqueue {1e6 * (t1 - t0) / a.reps:.1f} us/call)")
