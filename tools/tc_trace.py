"""Profiling aid: one traced launch of the tensor-core RX kernel, summarised per pipeline role. Needs a library built with
-DSL_TC_TRACE (tools/ab_build.sh trace -DSL_TC_TRACE; SELENITE_B200_LIB=build/ab/libtrace.so); SELENITE_B200_TC_TRACE names the dump."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import selenite_lite_b200 as slb
C, T = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 48000 * 2
x = torch.randint(-8000, 8000, (C, T, 2), dtype=torch.int16, device="cuda"); y = torch.empty_like(x)
d = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32)
for _ in range(2): d.rx_process(x, y)
torch.cuda.synchronize()
out = os.path.join(ROOT, "gpurun_out", "tc_trace.txt")
os.makedirs(os.path.dirname(out), exist_ok=True)
os.environ["SELENITE_B200_TC_TRACE"] = out
d.rx_process(x, y); torch.cuda.synchronize()
t = np.loadtxt(out)
names = ["prod:raw_empty", "conv:raw_full", "conv:a_empty", "conv:done", "mma:a_full", "mma:t_empty", "mma:issued", "epi:t_full", "epi:tmem->reg", "epi:zero-state", "epi:s_bar", "epi:BAR1", "epi:e_bar", "epi:BAR2", "epi:done"]
n = t.shape[0]
print("supertiles", n, "total clks", t[:, 14].max(), "per supertile", t[-1, 14] / n)
for r in list(range(20, 26)):
    print(r, " ".join("%s=%d" % (names[i].split(":")[1], t[r, i] - t[r, 0]) for i in range(15)))
mid = slice(10, n - 2)
def d_(a, b): return np.mean(t[mid, a] - t[mid, b])
print("period (epi done to done): %.0f" % np.mean(np.diff(t[mid, 14])))
print("conv: wait raw %.0f  wait a_empty %.0f  work %.0f" % (d_(1, 0), d_(2, 1), d_(3, 2)))
print("mma: a_full after conv done %.0f  wait t_empty %.0f  issue %.0f ; t_full seen by epi after issue %.0f" % (d_(4, 3), d_(5, 4), d_(6, 5), d_(7, 6)))
if np.mean(t[mid, 9]) > np.mean(t[mid, 13]):   # split epilogue: stamp 9 sits after the envelope walk and the gain
    print("epi: tmem %.0f level1+wait s_bar %.0f BAR1 %.0f corr+wait e_bar %.0f BAR2 %.0f walk+gain %.0f pack+store %.0f total %.0f" % (d_(8, 7), d_(10, 8), d_(11, 10), d_(12, 11), d_(13, 12), d_(9, 13), d_(14, 9), d_(14, 7)))
else:
    print("epi: tmem %.0f zero-state %.0f wait s_bar %.0f BAR1 %.0f corr+wait e_bar %.0f BAR2 %.0f tail %.0f total %.0f" % (d_(8, 7), d_(9, 8), d_(10, 9), d_(11, 10), d_(12, 11), d_(13, 12), d_(14, 13), d_(14, 7)))
