"""Host-side logic of the tensor-core RX-SSB-f32 kernel (sl_rx_ssb_tc.cu), checked on the CPU:
 * the 24-bit FIR taps recovered from the tcgen05 operand planes reproduce the overlap-save FFT filter of the oracle
   chain (arm_cfft_f32 / arm_cmplx_mult_cmplx_f32 / inverse, restated here in float64 numpy) far inside 1e-5;
 * a mask that is not the DFT of a 129-tap filter is refused (the FFT kernel keeps serving it);
 * the kernel's time-parallel biquad (zero-state blocks of 48, warp-local walk with A^48, cross-warp walk with A^192,
   zero-input correction) restated in numpy equals the sequential arm_biquad_cascade_df2T_f32 restatement."""
import ctypes as C

import numpy as np
import pytest

import selenite_lite_b200 as slb
from selenite_lite_b200 import _lib


def tc_taps(mask):
    hr = np.zeros(129, np.int32); hi = np.zeros(129, np.int32); unit = C.c_float(0)
    rc = _lib.load().slb_design_tc_taps(mask.ctypes.data, hr.ctypes.data, hi.ctypes.data, C.byref(unit))
    return rc, hr, hi, unit.value


@pytest.mark.parametrize("mode", [slb.MODE_USB, slb.MODE_LSB, slb.MODE_CW, slb.MODE_CWR, slb.MODE_DIG])
def test_integer_fir_equals_overlap_save(mode, rng):
    mask = np.zeros((512, 2), np.float32)
    assert _lib.load().slb_design_mask(48000, mode, mask.ctypes.data) == 0
    rc, hr, hi, unit = tc_taps(mask)
    assert rc == 0 and 8323000 <= max(np.abs(hr).max(), np.abs(hi).max()) <= 8323072    # the largest tap uses the full 24 bits
    T = 384 * 6
    x = rng.integers(-20000, 20000, (T + 128, 2)).astype(np.int16)  # 128 frames of history + the stream
    # oracle filter in float64: 512-point frames hopping by 384, keep the last 384 (1/32768 of arm_q15_to_float folded in)
    H = mask[:, 0].astype(np.float64) + 1j * mask[:, 1].astype(np.float64)
    z = (x[:, 0].astype(np.float64) + 1j * x[:, 1].astype(np.float64)) / 32768.0
    ref = np.concatenate([np.fft.ifft(np.fft.fft(z[f:f + 512]) * H)[128:] for f in range(0, T, 384)]).real
    # the kernel's form: exact integer FIR, one float scale
    acc = np.zeros(T, np.int64)
    for d in range(129):
        seg = x[128 - d:128 - d + T].astype(np.int64)
        acc += int(hr[d]) * seg[:, 0] - int(hi[d]) * seg[:, 1]
    got = acc.astype(np.float64) * unit
    # white full-band input is the worst case (95 % of its power is rejected): the 2^-24 tap quantisation and the float32
    # rounding of the mask itself both scale with the INPUT level; measured <= 6e-7 of the output rms, bound 1e-6
    assert np.max(np.abs(got - ref)) < 1e-6 * np.sqrt(np.mean(ref ** 2))


def test_non_fir_mask_is_refused(rng):
    mask = rng.standard_normal((512, 2)).astype(np.float32)
    assert tc_taps(mask)[0] == slb.ERR_UNSUPPORTED if hasattr(slb, "ERR_UNSUPPORTED") else tc_taps(mask)[0] != 0
    assert tc_taps(np.zeros((512, 2), np.float32))[0] != 0


def test_block_parallel_biquad_equals_sequential(port, rng):
    p = slb.default_rx_f32_params(48000)
    coef = np.array(p.biquad[:10], np.float32)
    Mp = np.zeros(64, np.float32); M192 = np.zeros(16, np.float32); Cr = np.zeros(192, np.float32)
    assert _lib.load().slb_biquad_tc_tables(coef.ctypes.data, Mp.ctypes.data, M192.ctypes.data, Cr.ctypes.data) == 0
    Mp = Mp.reshape(4, 4, 4).astype(np.float64); M192 = M192.reshape(4, 4).astype(np.float64); Cr = Cr.reshape(48, 4).astype(np.float64)
    assert np.allclose(Mp[0], np.eye(4)) and np.allclose(Mp[2], Mp[1] @ Mp[1], atol=1e-6) and np.allclose(M192, Mp[2] @ Mp[2], atol=1e-6)
    x = (rng.standard_normal(16 * 48) * 0.3).astype(np.float32)
    carry = (rng.standard_normal(4) * 0.1).astype(np.float32)
    ref_y, ref_state = port.biquad_df2T_f32(coef.reshape(2, 5), 2, carry, x, 48)
    zs = np.zeros((16, 48)); z = np.zeros((16, 4))
    for q in range(16):
        y, st = port.biquad_df2T_f32(coef.reshape(2, 5), 2, np.zeros(4, np.float32), x[48 * q:48 * q + 48], 48)
        zs[q] = y; z[q] = st
    start = np.zeros((16, 4)); S = carry.astype(np.float64)
    for w in range(4):
        P = np.zeros(4)
        for a in range(4):
            start[4 * w + a] = P + Mp[a] @ S                # level 1 (inside the warp) + the warp's start state
            P = Mp[1] @ P + z[4 * w + a]
        S = M192 @ S + P                                     # level 2: P is now the warp's zero-start end state
    y = zs + (Cr @ start.T).T
    end = Mp[1] @ start[15] + z[15]
    scale = np.sqrt(np.mean(ref_y.astype(np.float64) ** 2))
    assert np.max(np.abs(y.reshape(-1) - ref_y)) < 3e-6 * scale
    assert np.max(np.abs(end - ref_state)) < 3e-6 * max(1.0, np.max(np.abs(ref_state)))
