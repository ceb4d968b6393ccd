"""Step time against time under sustained load, with nvidia-smi clocks / power sampled beside it: is the difference
between a short run (1.68 ms per cfg2 step) and the bench's >= 1 s region (1.86 ms) the power cap?
    python tools/power_timeline.py [--seconds 4] [--what step|build|lookup]"""
import argparse, os, statistics, subprocess, sys, threading, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CONFIGS, SEED  # noqa: E402
from raft_optical_flow_b200 import CorrBlock, _cabi  # noqa: E402
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg2")
ap.add_argument("--seconds", type=float, default=4.0)
ap.add_argument("--what", default="step")
ap.add_argument("--mode", default=None)
a = ap.parse_args()
B, C, H, W, r, L, iters, _ = CONFIGS[a.config]
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(SEED)
fs = [(0.75 * torch.randn(2, B, C, H, W, generator=g)).to(dev) for _ in range(2)]
ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
grid = torch.stack([xs, ys]).float()[None]
dev_c = (grid + 4.0 * torch.randn(iters, B, 2, H, W, generator=g)).to(dev)
blk0 = CorrBlock(fs[0][0], fs[0][1], num_levels=L, radius=r, mode=a.mode)
torch.cuda.synchronize()

samples = []
stop = False
def sampler():
    p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_throttle_reasons.active",
                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
    while not stop:
        line = p.stdout.readline()
        if not line:
            break
        samples.append((time.perf_counter(), line.strip()))
    p.terminate()
th = threading.Thread(target=sampler, daemon=True)
th.start()
time.sleep(1.0)  # idle baseline samples

def one(k, ev):
    ev[0].record()
    if a.what == "lookup":
        for i in range(iters):
            blk0(dev_c[i])
        ev[1].record()
        return
    f = fs[k % 2]
    blk = CorrBlock(f[0], f[1], num_levels=L, radius=r, mode=a.mode)
    ev[1].record()
    if a.what == "step":
        for i in range(iters):
            blk(dev_c[i])

recs = []
t_start = time.perf_counter()
k = 0
with torch.no_grad():
    while time.perf_counter() - t_start < a.seconds:
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(20)]
        end = torch.cuda.Event(enable_timing=True)
        tb = time.perf_counter()
        for i in range(20):
            one(k, evs[i]); k += 1
        end.record()
        torch.cuda.synchronize()
        for i in range(20):
            nxt = evs[i + 1][0] if i + 1 < 20 else end
            recs.append((tb - t_start, evs[i][0].elapsed_time(nxt), evs[i][0].elapsed_time(evs[i][1])))
t_end = time.perf_counter()
time.sleep(0.3)
stop = True
print(f"what={a.what} mode={a.mode or 'default'}: {len(recs)} units in {t_end - t_start:.2f} s")
nb = 10
for j in range(nb):
    lo, hi = j * a.seconds / nb, (j + 1) * a.seconds / nb
    sel = [x for x in recs if lo <= x[0] < hi]
    if not sel:
        continue
    sm = [s[1] for s in samples if t_start + lo <= s[0] < t_start + hi]
    print(f"  t={lo:4.1f}-{hi:4.1f}s  unit {statistics.median(x[1] for x in sel) * 1e3:7.1f} us  first part {statistics.median(x[2] for x in sel) * 1e3:6.1f} us   smi[sm MHz, mem MHz, W, C, reasons]: {sm[len(sm) // 2] if sm else '-'}")
idle = [s[1] for s in samples if s[0] < t_start]
print("  idle before:", idle[-1] if idle else "-")
