#!/usr/bin/env python
"""End-to-end (host buffers) throughput of slb_rx_process_host against the slice size of its pipeline (SELENITE_B200_SLICE_BYTES),
with the plain duplex pinned-copy ceiling of the same bytes beside it. Tuning aid for the default in sl_capi.cu; not the bench contract."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import selenite_lite_b200 as slb  # noqa: E402

C, T = 1024, 480000
dev = torch.device("cuda", 0)
x = torch.randint(-8000, 8000, (C, T, 2), dtype=torch.int16, device=dev)
y = torch.empty_like(x)
xh = torch.empty((C, T, 2), dtype=torch.int16).pin_memory(); yh = torch.empty((C, T, 2), dtype=torch.int16).pin_memory()
xh.copy_(x); torch.cuda.synchronize()
s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)


def duplex():
    with torch.cuda.stream(s_in):
        x.copy_(xh, non_blocking=True)
    with torch.cuda.stream(s_out):
        yh.copy_(y, non_blocking=True)


def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


t_c = timed(duplex)
out = {"ceiling_ms": t_c * 1e3, "ceiling_GBps_each_way": C * T * 4 / t_c / 1e9, "slices": {}}
for mb in (4, 8, 16, 32, 64, 128, 256):
    os.environ["SELENITE_B200_SLICE_BYTES"] = str(mb << 20)
    d = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32)
    t = timed(lambda: d.rx_process_pinned(xh, yh))
    out["slices"]["%d MiB" % mb] = {"ms": t * 1e3, "Gsamples_per_s": C * T / t / 1e9, "frac_of_ceiling": t_c / t}
    del d
print(json.dumps(out))
