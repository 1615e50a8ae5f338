"""GPU parity of the FM detector of the RX-SSB-f32 chain (FT-817 mode byte 0x08, Core/Inc/rxtx_if.h:40 -> DSP_Set_Mode,
Core/Src/rxtx_if.c:647), called through the C ABI, against the oracle and the committed golden vectors.

The detector is OURS like every composition here (the firmware's DSP_Set_Mode is empty, dsp_if.c:367-370): a limiter-discriminator
that needs no arctangent (CMSIS-DSP V1.5.3 has none), composed from arm_cmplx_conj_f32 / arm_cmplx_mult_cmplx_f32 /
arm_cmplx_mag_f32 in oracle/chains.inc.c. On the GPU it runs on the complex-detector tensor-core kernel (sl_rx_am_tc.cu)."""
import os

import numpy as np
import pytest

import selenite_lite_b200 as slb
from test_golden import GOLD, fm_audio_ok, fm_int16_ok
from test_gpu_rx_ssb_f32 import run_gpu, check_int16, audio_tolerance

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["centre", "offset"])
def test_golden_vectors(name):
    g = np.load(os.path.join(GOLD, "rx_fm_f32.npz"))
    d = slb.DspIf(1, chain=slb.CHAIN_RX_SSB_F32)
    d.DSP_Set_Mode(slb.MODE_FM)
    assert np.array_equal(d.mask(slb.MODE_FM), g["fm_mask"])
    y, audio, gain = run_gpu(d, g["fm_%s_in" % name][None])
    assert fm_audio_ok(audio[0], g["fm_%s_audio" % name])
    assert np.allclose(gain[0], g["fm_%s_gain" % name], rtol=2e-5)
    assert fm_int16_ok(y[0], g["fm_%s_out" % name])


@pytest.mark.parametrize("channels,frames", [(1, 384), (3, 768), (9, 1920), (20, 4608)])
def test_vs_oracle_ragged_shapes_mixed_modes(best_oracle, channels, frames):
    """FM channels next to SSB and AM channels of one context (three kernels serve one call), ragged sizes, several groups."""
    modes = [slb.MODE_FM, slb.MODE_USB, slb.MODE_FM, slb.MODE_AM, slb.MODE_LSB]
    x = slb.synth_iq(channels, frames)
    d = slb.DspIf(channels, chain=slb.CHAIN_RX_SSB_F32)
    for c in range(channels):
        m = modes[c % len(modes)]
        d.DSP_Set_Mode(m, channel=c)
        if m == slb.MODE_FM:
            x[c] = slb.synth_fm(1, frames, first_channel=c, carrier_hz=200.0 * (c % 5))[0]
        if m == slb.MODE_LSB:
            x[c, :, 1] = -x[c, :, 1]
    y, audio, gain = run_gpu(d, x)
    for c in range(channels):
        m = modes[c % len(modes)]
        exp, a, g_, _ = best_oracle.rx_ssb_f32(d.oracle_params(m), x[c])
        if m == slb.MODE_FM:
            assert fm_audio_ok(audio[c], a), c
            assert np.allclose(gain[c], g_, rtol=2e-5)
            assert fm_int16_ok(y[c], exp), c
        else:
            assert np.all(np.abs(audio[c] - a) <= audio_tolerance(a) + 1e-9), c
            check_int16(y[c], exp)


def test_state_carries_across_calls_and_checkpoint(best_oracle):
    """The last baseband sample is carried like the biquad state: a stream cut into calls (at supertile and at hop boundaries),
    saved and restored in between, gives the result of the uncut stream."""
    C, T = 10, 1536 * 4
    x = slb.synth_fm(C, T)
    whole = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32); whole.DSP_Set_Mode(slb.MODE_FM)
    y_whole, _, _ = run_gpu(whole, x, want_audio=False)
    cut = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32); cut.DSP_Set_Mode(slb.MODE_FM)
    parts = []
    for a, b in ((0, 1536), (1536, 1536 + 768), (1536 + 768, 4608), (4608, T)):
        parts.append(run_gpu(cut, np.ascontiguousarray(x[:, a:b]), want_audio=False)[0])
        if b == 4608:
            snap = cut.state_save()
            cut = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32); cut.DSP_Set_Mode(slb.MODE_FM); cut.state_load(snap)
    assert np.array_equal(np.concatenate(parts, 1), y_whole)
    exp, _ = best_oracle.rx_ssb_f32_batch(whole.oracle_params(slb.MODE_FM), x)
    for c in range(C):
        assert fm_int16_ok(y_whole[c], exp[c]), c


def test_fm_stays_on_its_kernel_when_the_fft_path_is_forced(best_oracle):
    """slb_set_rx_path (SLB_RX_PATH_FFT) moves SSB channels to the FFT kernel; FM has no FFT-kernel detector and keeps its own."""
    C, T = 4, 1536 * 2
    x = slb.synth_iq(C, T)
    x[1] = slb.synth_fm(1, T)[0]; x[3] = slb.synth_fm(1, T, first_channel=3)[0]
    d = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32); d.set_rx_path(slb.RX_PATH_FFT)
    for c in (1, 3):
        d.DSP_Set_Mode(slb.MODE_FM, channel=c)
    y, audio, gain = run_gpu(d, x)
    for c in range(C):
        m = slb.MODE_FM if c in (1, 3) else slb.MODE_USB
        exp, a, g_, _ = best_oracle.rx_ssb_f32(d.oracle_params(m), x[c])
        assert (fm_int16_ok if m == slb.MODE_FM else (lambda u, v: (check_int16(u, v), True)[1]))(y[c], exp), c


def test_fm_is_refused_where_there_is_no_discriminator():
    for chain in (slb.CHAIN_TX_SSB_F32, slb.CHAIN_RX_SSB_Q15):
        d = slb.DspIf(2, chain=chain)
        with pytest.raises(slb.SeleniteError):
            d.DSP_Set_Mode(slb.MODE_FM)
    d = slb.DspIf(2, chain=slb.CHAIN_RX_SSB_F32)
    bad = np.zeros(512, np.complex64); bad[3] = 1.0                  # a single bin: impulse response of 512 taps, no 129-tap FIR
    with pytest.raises(slb.SeleniteError):
        d.set_mask(slb.MODE_FM, bad)
    d.DSP_Set_Mode(slb.MODE_FM)                                      # the default mask is still in place


@pytest.mark.parametrize("amp", [0.25, 0.03, 0.004])
def test_weak_carriers_and_the_squelch_floor(best_oracle, amp):
    """Carriers above, at and below the limiter floor (2^-5 of full scale in baseband amplitude): the same bars on both sides of it."""
    C, T = 6, 1536 * 2
    x = slb.synth_fm(C, T, amp=amp, sigma=0.0005)
    d = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32); d.DSP_Set_Mode(slb.MODE_FM)
    y, audio, gain = run_gpu(d, x)
    for c in range(C):
        exp, a, g_, _ = best_oracle.rx_ssb_f32(d.oracle_params(slb.MODE_FM), x[c])
        assert fm_audio_ok(audio[c], a), c
        assert np.allclose(gain[c], g_, rtol=2e-5)
        assert fm_int16_ok(y[c], exp), c
