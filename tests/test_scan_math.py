"""The time-parallel evaluation of the df2T cascade used by the fused kernel's recurrence warp,
checked on the CPU: tables from the product (slb_biquad_scan_tables) + the kernel's algorithm restated in numpy
(lane = one 48-sample block = two 24-sample runs; zero-state runs, block combine with M24, 5-step scan with
M48^(2^k), start states, zero-input correction) against the sequential arm_biquad_cascade_df2T_f32 restatement."""
import numpy as np

import selenite_lite_b200 as slb
from selenite_lite_b200 import _lib


def test_zero_state_plus_correction_equals_sequential(port, rng):
    p = slb.default_rx_f32_params(48000)
    coef = np.array(p.biquad[:10], np.float32)
    M = np.zeros(96, np.float32); Cr = np.zeros(96, np.float32)
    assert _lib.load().slb_biquad_scan_tables(coef.ctypes.data, M.ctypes.data, Cr.ctypes.data) == 0
    M = M.reshape(6, 4, 4).astype(np.float64); Cr = Cr.reshape(24, 4).astype(np.float64)
    assert np.allclose(M[1], M[0] @ M[0], atol=1e-6) and np.allclose(M[3], M[2] @ M[2], atol=1e-6)
    x = (rng.standard_normal(32 * 48) * 0.3).astype(np.float32)
    s0 = (rng.standard_normal(4) * 0.1).astype(np.float32)
    ref_y, ref_state = port.biquad_df2T_f32(coef.reshape(2, 5), 2, s0, x, 48)
    zs = np.zeros((32, 2, 24), np.float32); zA = np.zeros((32, 4)); zB = np.zeros((32, 4))
    for l in range(32):
        for h, zz in ((0, zA), (1, zB)):
            y, st = port.biquad_df2T_f32(coef.reshape(2, 5), 2, np.zeros(4, np.float32), x[48 * l + 24 * h:48 * l + 24 * h + 24], 24)
            zs[l, h] = y; zz[l] = st
    z = zB + (M[0] @ zA.T).T                              # zero-start end state of each block
    z[0] += M[1] @ s0                                     # lane 0 folds the carried state in
    for lv in range(5):                                   # Kogge-Stone with M48^(2^lv)
        d = 1 << lv
        z[d:] = z[d:] + (M[lv + 1] @ z[:-d].T).T
    startA = np.vstack([s0[None, :].astype(np.float64), z[:-1]])
    startB = zA + (M[0] @ startA.T).T
    y = zs.astype(np.float64)
    y[:, 0] += (Cr @ startA.T).T
    y[:, 1] += (Cr @ startB.T).T
    scale = np.sqrt(np.mean(ref_y.astype(np.float64) ** 2))
    assert np.max(np.abs(y.reshape(-1) - ref_y)) < 3e-6 * scale
    assert np.max(np.abs(z[-1] - ref_state)) < 3e-6 * max(1.0, np.max(np.abs(ref_state)))
