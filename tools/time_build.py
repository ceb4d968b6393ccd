"""(RCB_USE_DEBUG_LIB=1 + `python -m raft_optical_flow_b200.build --debug` for the RCB_TC_* hooks.)
Times rcb_corr_build (pack + main kernel) through the C ABI with preallocated buffers (no Python/torch
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
ap.add_argument("--mode", default="f16f8")
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--between", default="none", choices=["none", "lookups", "fill", "sleep"],
                help="what runs between the timed builds: nothing (tight loop), the step's lookups, a 512 MB fill, "
                     "or ~1 ms of idle")
ap.add_argument("--python", action="store_true", help="with --between lookups: go through CorrBlock(...) and blk(coords) "
                "like bench.py instead of the raw C calls on preallocated buffers")
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
if a.between != "none":
    # each build timed on its own, with something else on the stream in between (what the bench step looks like)
    from raft_optical_flow_b200 import CorrBlock
    plan = _cabi.LookupPlan(pyr.ptrs, B, H, W, L, r, 0)
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    coords = (torch.stack([xs, ys]).float()[None] + 4.0 * torch.randn(B, 2, H, W, generator=g)).to(dev).contiguous()
    outs = [torch.empty((B, L * (2 * r + 1) ** 2, H, W), device=dev) for _ in range(2)]
    junk = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.reps)]
    blk = None
    for k in range(a.reps if not a.python else 0):
        if a.between == "lookups":
            for i in range(iters):
                _cabi.check(lib.rcb_corr_lookup_planned(plan.ptr, coords.data_ptr(), outs[i & 1].data_ptr(), s), "lookup")
        elif a.between == "fill":
            junk.fill_(k & 255)
        else:
            torch.cuda._sleep(2_000_000)
        evs[k][0].record()
        call()
        evs[k][1].record()
    cs = [coords.clone() for _ in range(iters)]
    for k in range(a.reps if a.python else 0):
        evs[k][0].record()
        blk = CorrBlock(f1, f2, num_levels=L, radius=r, mode=a.mode)
        evs[k][1].record()
        for i in range(iters):
            out = blk(cs[i])
    torch.cuda.synchronize()
    ts = sorted(x.elapsed_time(y) * 1e3 for x, y in evs[2:])
    print(f"{a.config} {a.mode} between={a.between}: build median {ts[len(ts) // 2]:.1f} us, min {ts[0]:.1f}, max {ts[-1]:.1f}")
if os.environ.get("RCB_TC_PROF") == "1":
    prof = torch.zeros(16 * 148 + 20 * 148, dtype=torch.int64, device=dev)
    os.environ["RCB_TC_PROF_PTR"] = hex(prof.data_ptr())
    call()
    torch.cuda.synchronize()
    del os.environ["RCB_TC_PROF_PTR"]
    pr = prof[:16 * 148].view(148, 16).double()
    phases = ["wait_acc_full", "tmem_ld", "release_acc", "scale+pool", "wait_store_read", "staging+l2_store", "fence",
              "pair_barrier", "tma_issue+commit", "level3"]
    php = prof[16 * 148:].view(148, 2, 10).double().mean(0)
    for b in range(2):
        print(f"band {b} phases (mean cycles per CTA):", {n: int(php[b, i].item()) for i, n in enumerate(phases)})
    names = ["prod_total", "prod_wait_b_empty", "unused", "mma_total", "mma_wait_b_full", "mma_wait_acc_empty",
             "mma_wait_a_full", "band0_total", "band0_wait_acc_full", "band0_wait_store_read", "tiles",
             "band1_total", "band1_wait_acc_full", "band1_wait_store_read"]
    print("per-CTA mean cycles:", {n: int(pr[:, i].mean().item()) for i, n in enumerate(names)})
print(f"{a.config} {a.mode} skip={os.environ.get('RCB_TC_DEBUG_SKIP','0')} nstage={os.environ.get('RCB_TC_NSTAGE','-')} "
      f"build {e0.elapsed_time(e1) / a.reps * 1e3:.1f} us (host enqueue {1e6 * (t1 - t0) / a.reps:.1f} us/call)")
