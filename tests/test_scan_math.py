"""The time-parallel evaluation of the df2T cascade used by the fused kernel, checked on the CPU:
tables from the product (slb_biquad_scan_tables) + the kernel's algorithm in numpy float32 vs the sequential
arm_biquad_cascade_df2T_f32 restatement."""
import numpy as np

import selenite_lite_b200 as slb
from selenite_lite_b200 import _lib


def test_zero_state_plus_correction_equals_sequential(port, rng):
    p = slb.default_rx_f32_params(48000)
    coef = np.array(p.biquad[:10], np.float32)
    M = np.zeros(80, np.float32); Cr = np.zeros(192, np.float32)
    assert _lib.load().slb_biquad_scan_tables(coef.ctypes.data, M.ctypes.data, Cr.ctypes.data) == 0
    M = M.reshape(5, 4, 4).astype(np.float64); Cr = Cr.reshape(48, 4).astype(np.float64)
    x = (rng.standard_normal(32 * 48) * 0.3).astype(np.float32)
    s0 = (rng.standard_normal(4) * 0.1).astype(np.float32)
    ref_y, ref_state = port.biquad_df2T_f32(coef.reshape(2, 5), 2, s0, x, 48)
    # lanes: zero-state runs
    zs = np.zeros((32, 48), np.float32); z = np.zeros((32, 4), np.float64)
    for k in range(32):
        y, st = port.biquad_df2T_f32(coef.reshape(2, 5), 2, np.zeros(4, np.float32), x[48 * k:48 * k + 48], 48)
        zs[k] = y; z[k] = st
    z[0] += M[0] @ s0
    for lv in range(5):                                   # Kogge-Stone with M^(2^lv)
        d = 1 << lv
        z[d:] = z[d:] + (M[lv] @ z[:-d].T).T
    start = np.vstack([s0[None, :].astype(np.float64), z[:-1]])
    y = zs + (Cr @ start.T).T
    scale = np.sqrt(np.mean(ref_y.astype(np.float64) ** 2))
    assert np.max(np.abs(y.reshape(-1) - ref_y)) < 3e-6 * scale
    assert np.max(np.abs(z[-1] - ref_state)) < 3e-6 * max(1.0, np.max(np.abs(ref_state)))
