"""Measures this GPU's pure-write, pure-read and copy DRAM bandwidth with plain torch ops (CUDA events, best of 5) --
the denominators the store-bound build kernel and the gather-bound lookup are compared with.
    python tools/bw_probe.py"""
import torch
dev = torch.device("cuda:0")
n = 2200 * 1024 * 1024 // 4  # 2.2 GB of fp32, the size of the cfg2 pyramid
a = torch.empty(n, dtype=torch.float32, device=dev)
b = torch.empty(n, dtype=torch.float32, device=dev)
def best(fn, reps=5):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)
gb = n * 4 / 1e9
a.fill_(1.0); b.copy_(a); torch.cuda.synchronize()
t = best(lambda: a.fill_(2.0)); print(f"write  (fill_):  {gb / t * 1e3:7.0f} GB/s  ({t * 1e3:.0f} us for {gb:.2f} GB)")
t = best(lambda: a.zero_()); print(f"write  (zero_):  {gb / t * 1e3:7.0f} GB/s")
t = best(lambda: torch.sum(a)); print(f"read   (sum):    {gb / t * 1e3:7.0f} GB/s")
t = best(lambda: b.copy_(a)); print(f"copy   (r+w):    {2 * gb / t * 1e3:7.0f} GB/s")
