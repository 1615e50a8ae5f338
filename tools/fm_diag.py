import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import oracle_lib, selenite_lite_b200 as slb
from test_gpu_rx_ssb_f32 import run_gpu
from test_golden import GOLD, audio_tolerance
orc = oracle_lib.Oracle("ref")
g = np.load(os.path.join(GOLD, "rx_fm_f32.npz"))
for name in ("centre", "offset"):
    d = slb.DspIf(1, chain=slb.CHAIN_RX_SSB_F32); d.DSP_Set_Mode(slb.MODE_FM)
    y, audio, gain = run_gpu(d, g["fm_%s_in" % name][None])
    ra = g["fm_%s_audio" % name]; err = np.abs(audio[0] - ra); tol = audio_tolerance(ra)
    print(name, "audio err first384 max %.3g at %d; later max %.3g (x tol %.2f) at %d" % (err[:384].max(), err[:384].argmax(), err[384:].max(), (err[384:] / tol[384:]).max(), 384 + (err[384:] / tol[384:]).argmax()))
    print("   gain rel diff max %.3g" % np.abs(gain[0] / g["fm_%s_gain" % name] - 1).max(), " int16 diff max", np.abs(y[0].astype(int) - g["fm_%s_out" % name].astype(int)).max(), "share %.3f" % np.mean(y[0] != g["fm_%s_out" % name]))
    bad = np.nonzero(err > 10 * tol)[0]
    print("   samples with err > 10 tol:", bad[:20], len(bad))
    for i in bad[:6]:
        print("     n=%d gpu %.6g ref %.6g" % (i, audio[0][i], ra[i]))
