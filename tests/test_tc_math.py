"""Host-side logic of the tensor-core RX-SSB-f32 kernel (sl_rx_ssb_tc.cu), checked on the CPU:
 * the 129 taps recovered from a mask reproduce the overlap-save FFT filter of the oracle chain (arm_cfft_f32 /
   arm_cmplx_mult_cmplx_f32 / inverse, restated here in float64 numpy), and the tcgen05 operand planes built from them
   (taps composed with the biquad's zero-state response, 24-bit digits) reproduce FIR + biquad far inside 1e-5;
 * a mask that is not the DFT of a 129-tap filter is refused (the FFT kernel keeps serving it);
 * the kernel's time-parallel biquad (zero-state blocks of 48, warp-local walk with A^48, cross-warp walk with A^192,
   zero-input correction) restated in numpy equals the sequential arm_biquad_cascade_df2T_f32 restatement."""
import ctypes as C

import numpy as np
import pytest

import selenite_lite_b200 as slb
from selenite_lite_b200 import _lib


def tc_taps(mask):
    hr = np.zeros(129, np.float64); hi = np.zeros(129, np.float64)
    rc = _lib.load().slb_design_tc_taps(mask.ctypes.data, hr.ctypes.data, hi.ctypes.data)
    return rc, hr, hi


def default_mask(mode):
    mask = np.zeros((512, 2), np.float32)
    assert _lib.load().slb_design_mask(48000, mode, mask.ctypes.data) == 0
    return mask


def biquad_zero_state_f64(coef, x):
    """arm_biquad_cascade_df2T_f32.c:551-562 in float64, two stages, zero start: output and end state {d1_0,d2_0,d1_1,d2_1}."""
    c = [float(v) for v in coef]
    d = [0.0, 0.0, 0.0, 0.0]; y = np.zeros(len(x))
    for n, xn in enumerate(x):
        y0 = c[0] * xn + d[0]; d[0] = (c[1] * xn + c[3] * y0) + d[1]; d[1] = c[2] * xn + c[4] * y0
        y1 = c[5] * y0 + d[2]; d[2] = (c[6] * y0 + c[8] * y1) + d[3]; d[3] = c[7] * y0 + c[9] * y1
        y[n] = y1
    return y, np.array(d)


@pytest.mark.parametrize("mode", [slb.MODE_USB, slb.MODE_LSB, slb.MODE_CW, slb.MODE_CWR, slb.MODE_DIG])
def test_fir_taps_equal_overlap_save(mode, rng):
    mask = default_mask(mode)
    rc, hr, hi = tc_taps(mask)
    assert rc == 0
    T = 384 * 6
    x = rng.integers(-20000, 20000, (T + 128, 2)).astype(np.int16)  # 128 frames of history + the stream
    # oracle filter in float64: 512-point frames hopping by 384, keep the last 384 (1/32768 of arm_q15_to_float folded in)
    H = mask[:, 0].astype(np.float64) + 1j * mask[:, 1].astype(np.float64)
    z = (x[:, 0].astype(np.float64) + 1j * x[:, 1].astype(np.float64)) / 32768.0
    ref = np.concatenate([np.fft.ifft(np.fft.fft(z[f:f + 512]) * H)[128:] for f in range(0, T, 384)]).real
    got = np.zeros(T)
    for d in range(129):
        seg = x[128 - d:128 - d + T].astype(np.float64) / 32768.0
        got += hr[d] * seg[:, 0] - hi[d] * seg[:, 1]
    # what is left is the part of the mask's impulse response beyond 129 taps (float32 rounding of the mask): ~1e-7 of the INPUT
    assert np.max(np.abs(got - ref)) < 3e-7 * np.sqrt(np.mean(ref ** 2))


@pytest.mark.parametrize("mode", [slb.MODE_USB, slb.MODE_CWR, slb.MODE_DIG])
def test_operand_planes_equal_fir_plus_biquad(mode, rng):
    """The kernel's tcgen05 operand planes (taps composed with the biquad's zero-state response, 24-bit digits, UMMA layout),
    evaluated in integers on raw windows, against the float64 FIR followed by the float64 df2T cascade: audio and end state."""
    mask = default_mask(mode)
    rc, hr, hi = tc_taps(mask)
    coef = np.array(slb.default_rx_f32_params(48000).biquad[:10], np.float32)
    lib = _lib.load()
    for trial in range(3):
        amp = (20000, 300, 32767)[trial]                                  # loud, weak, and rail-to-rail samples
        w = rng.integers(-amp, amp + 1, (176, 2)).astype(np.int16)
        if trial == 2:
            w[::7] = (32767, -32768); w[3::11] = (-32768, 32767)
        out = np.zeros(52, np.float64)
        assert lib.slb_design_tc_block(mask.ctypes.data, coef.ctypes.data, w.ctypes.data, out.ctypes.data) == 0
        xs = w.astype(np.float64) / 32768.0
        y = np.array([sum(hr[d] * xs[128 + m - d, 0] - hi[d] * xs[128 + m - d, 1] for d in range(129)) for m in range(48)])
        a, st = biquad_zero_state_f64(coef, y)
        scale = np.sqrt(np.mean(xs ** 2))                                # 24-bit quantisation error scales with the INPUT level
        assert np.max(np.abs(out[:48] - a)) < 2e-7 * scale, (trial, np.max(np.abs(out[:48] - a)) / scale)
        assert np.max(np.abs(out[48:] - st)) < 2e-7 * scale * max(1.0, np.max(np.abs(st)) / max(np.max(np.abs(a)), 1e-12)), trial


def test_non_fir_mask_is_refused(rng):
    mask = rng.standard_normal((512, 2)).astype(np.float32)
    assert tc_taps(mask)[0] != 0
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


@pytest.mark.parametrize("mode", [slb.MODE_USB, slb.MODE_LSB])
def test_tx_operand_planes_equal_complex_fir(mode, rng):
    """TX (sl_tx_ssb_tc.cu): real mic samples x complex taps, two rails of tcgen05 operand planes, against the float64 FIR."""
    mask = default_mask(mode)
    rc, hr, hi = tc_taps(mask)
    assert rc == 0
    lib = _lib.load()
    for amp in (20000, 200, 32767):
        w = rng.integers(-amp, amp + 1, 192).astype(np.int16)
        if amp == 32767:
            w[::5] = 32767; w[2::7] = -32768
        out = np.zeros(96, np.float64)
        assert lib.slb_design_tc_tx_block(mask.ctypes.data, w.ctypes.data, out.ctypes.data) == 0
        m = w.astype(np.float64) / 32768.0
        ref_i = np.array([sum(hr[d] * m[128 + n - d] for d in range(129)) for n in range(48)])
        ref_q = np.array([sum(hi[d] * m[128 + n - d] for d in range(129)) for n in range(48)])
        scale = np.sqrt(np.mean(m ** 2))
        assert np.max(np.abs(out[0::2] - ref_i)) < 2e-7 * scale and np.max(np.abs(out[1::2] - ref_q)) < 2e-7 * scale


def test_q15_operand_planes_equal_arm_fir_q15(port, rng):
    """sl_rx_q15_tc.cu: the tcgen05 operand planes of the two 64-tap q15 FIRs (balanced byte digits, both rails in one operand)
    and the reconstruction (acc >> 15) = 2 S2 + ((256 S1 + S0) >> 15), bit for bit against arm_fir_q15 — random taps,
    full-scale samples, saturating outputs; taps that do not split into two signed bytes are refused."""
    lib = _lib.load()
    for trial in range(4):
        ti = rng.integers(-32639, 32640, 64).astype(np.int16); tq = rng.integers(-32639, 32640, 64).astype(np.int16)
        if trial == 0:
            p = slb.default_rx_q15_params(48000); ti = np.array(p.taps_i[:64], np.int16); tq = np.array(p.taps_q[:64], np.int16)
        w = rng.integers(-32768, 32768, (112, 2)).astype(np.int16)
        if trial == 3:
            w[:] = 32767; ti[:] = 32639; tq[:] = -32639                            # both rails saturate
        out = np.zeros(96, np.int32)
        assert lib.slb_design_q15_tc_block(ti.ctypes.data, tq.ctypes.data, w.ctypes.data, out.ctypes.data) == 0
        for rail, taps in ((0, ti), (1, tq)):
            x = np.ascontiguousarray(w[:, rail])
            # arm_fir_q15 takes pCoeffs time-reversed (oracle/chains.inc.c does the same); pState: ntaps + block - 1 (+1), zero history
            y, _ = port.fir_q15(np.ascontiguousarray(taps[::-1]), np.zeros(64 + 16, np.int16), x, 16)
            assert np.array_equal(out[48 * rail:48 * rail + 48], y[64:].astype(np.int32)), (trial, rail)
    bad = np.zeros(64, np.int16); bad[5] = 32700
    assert lib.slb_design_q15_tc_block(bad.ctypes.data, bad.ctypes.data, np.zeros((112, 2), np.int16).ctypes.data, np.zeros(96, np.int32).ctypes.data) != 0
