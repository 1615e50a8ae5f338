"""Scratch: why is bench.py's device path slow with the tensor-core kernel while bench_chains is fast?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import selenite_lite_b200 as slb
import bench

dev = torch.device("cuda", 0)
C, T = 1024, 480000
def timeit(d, x, y, steps=4):
    for _ in range(3): d.rx_process(x, y)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        d.rx_process(x, y); ev[i + 1].record()
    torch.cuda.synchronize()
    return [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(steps)]

g = torch.Generator(device=dev); g.manual_seed(1)
xr = torch.randint(-8000, 8000, (C, T, 2), dtype=torch.int16, device=dev, generator=g)
y = torch.empty_like(xr)
d = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32)
print("randint:", timeit(d, xr, y))
xt = bench.synth_on_gpu(torch, C, T, dev, 0)
print("tone+noise:", timeit(d, xt, y))
xs = (xt // 64).contiguous()
print("tone+noise / 64:", timeit(d, xs, y))
print("randint again:", timeit(d, xr, y))
xz = torch.zeros_like(xr)
print("zeros:", timeit(d, xz, y))
