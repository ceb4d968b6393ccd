"""H2D / D2H bandwidth from pinned memory: one copy, the copy split over several streams, and both directions at once
(the e2e leg of bench.py moves 130 MB in and 73 MB out per step).  python tools/pcie_probe.py"""
import torch, time
dev = torch.device("cuda:0")
n_in, n_out = 129761280, 72990720
hin = torch.empty(n_in, dtype=torch.uint8).pin_memory()
hout = torch.empty(n_out, dtype=torch.uint8).pin_memory()
din = torch.empty(n_in, dtype=torch.uint8, device=dev)
dout = torch.empty(n_out, dtype=torch.uint8, device=dev)
streams = [torch.cuda.Stream() for _ in range(5)]

def run(nsplit, with_d2h, reps=20):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        step = (n_in + nsplit - 1) // nsplit
        for i in range(nsplit):
            with torch.cuda.stream(streams[i]):
                din[i * step:(i + 1) * step].copy_(hin[i * step:(i + 1) * step], non_blocking=True)
        if with_d2h:
            with torch.cuda.stream(streams[4]):
                hout.copy_(dout, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"H2D in {nsplit} stream(s){' + D2H' if with_d2h else ''}: {dt * 1e3:.3f} ms per step, H2D {n_in / dt / 1e9:.1f} GB/s"
          + (f", D2H {n_out / dt / 1e9:.1f} GB/s" if with_d2h else ""))

for w in (False, True):
    for ns in (1, 2, 4):
        run(ns, w)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    with torch.cuda.stream(streams[4]):
        hout.copy_(dout, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 20
print(f"D2H alone: {n_out / dt / 1e9:.1f} GB/s")
