"""Device-resident throughput of the RX-SSB-f32 chain with every channel in AM (profiling aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import selenite_lite_b200 as slb
C, T = 1024, 480000
x = torch.randint(-8000, 8000, (C, T, 2), dtype=torch.int16, device="cuda"); y = torch.empty_like(x)
for path in (slb.RX_PATH_AUTO, slb.RX_PATH_FFT):
    d = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32); d.DSP_Set_Mode(slb.MODE_AM); d.set_rx_path(path)
    for _ in range(3): d.rx_process(x, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): d.rx_process(x, y)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("AM, path %d: %.3f ms/step, %.1f Gsamples/s" % (path, ms, C * T / ms / 1e6))
