"""GPU parity of the fused RX-SSB-f32 chain, called through the C ABI, against the oracle (the reference build when
it is present on this box, else the port) and the committed golden vectors.

Tolerances (north_star): float stages within 1e-5 relative per sample — made well-defined at zero crossings as
1e-5 * max(|ref|, block rms) — and 0.1 dB output SNR. The int16 result is a truncating quantisation of the float
chain (arm_float_to_q15.c:147), so a 1e-7 float difference can flip at most one LSB on a small fraction of samples;
the test allows |d| <= 1 LSB on < 2 % of samples and nothing else."""
import os

import numpy as np
import pytest

import selenite_lite_b200 as slb
from test_golden import GOLD, audio_tolerance, rx_params

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def run_gpu(d, x, want_audio=True):
    C, T = x.shape[0], x.shape[1]
    xd = torch.from_numpy(x).cuda()
    audio = torch.zeros((C, T), dtype=torch.float32, device="cuda") if want_audio else None
    gain = torch.zeros((C, T // 48), dtype=torch.float32, device="cuda") if want_audio else None
    d.set_debug_taps(audio, gain)
    y = d.rx_process(xd)
    torch.cuda.synchronize()
    d.set_debug_taps(None, None)
    return y.cpu().numpy(), (audio.cpu().numpy() if want_audio else None), (gain.cpu().numpy() if want_audio else None)


def check_int16(y, exp):
    d = np.abs(y.astype(np.int32) - exp.astype(np.int32))
    assert d.max() <= 1, "int16 output differs by %d LSB" % d.max()
    assert np.mean(d > 0) < 0.02, "%.2f %% of samples differ" % (100 * np.mean(d > 0))
    assert np.array_equal(y[..., 0], y[..., 1])          # stereo endpoint carries L = R


def snr_db(y, f0, fs=48000):
    """Tone power over residual power after projecting out the tone at f0."""
    s = y.astype(np.float64); n = np.arange(s.size)
    basis = np.stack([np.cos(2 * np.pi * f0 * n / fs), np.sin(2 * np.pi * f0 * n / fs)], 1)
    coef, *_ = np.linalg.lstsq(basis, s, rcond=None)
    tone = basis @ coef
    return 10 * np.log10(np.sum(tone ** 2) / np.sum((s - tone) ** 2))


@pytest.mark.parametrize("name,mode", [("usb", slb.MODE_USB), ("lsb", slb.MODE_LSB)])
def test_golden_vectors(name, mode):
    g = np.load(os.path.join(GOLD, "rx_ssb_f32.npz"))
    d = slb.DspIf(1, chain=slb.CHAIN_RX_SSB_F32)
    d.DSP_Set_Mode(mode)
    assert np.array_equal(d.mask(mode), g["rx_%s_mask" % name])          # the design the golden run used is the shipped one
    y, audio, gain = run_gpu(d, g["rx_%s_in" % name][None])
    ref_audio = g["rx_%s_audio" % name]
    err = np.abs(audio[0] - ref_audio); tol = audio_tolerance(ref_audio)
    assert np.all(err <= tol + 1e-9), "worst audio error %.2f x tolerance" % np.max(err / (tol + 1e-9))
    assert np.allclose(gain[0], g["rx_%s_gain" % name], rtol=2e-5)
    check_int16(y[0], g["rx_%s_out" % name])


@pytest.mark.parametrize("channels,frames", [(1, 384), (3, 768), (5, 1536), (7, 1920), (33, 4608), (130, 3072)])
def test_vs_oracle_ragged_shapes(best_oracle, channels, frames):
    """1..4-hop tiles, partial last tiles, more channels than fit one wave of CTAs, mixed modes per channel."""
    x = slb.synth_iq(channels, frames)
    d = slb.DspIf(channels, chain=slb.CHAIN_RX_SSB_F32)
    modes = [slb.MODE_USB, slb.MODE_LSB, slb.MODE_CW, slb.MODE_CWR, slb.MODE_DIG]
    for c in range(channels):
        d.DSP_Set_Mode(modes[c % len(modes)], channel=c)
        if modes[c % len(modes)] in (slb.MODE_CW, slb.MODE_CWR):
            # the 500 Hz CW filters get their tone inside the pass band (the per-sample 1e-5 is relative to the OUTPUT; a tone
            # the filter rejects is the subject of test_output_dominated_by_a_rejected_tone below)
            x[c] = slb.synth_iq(1, frames, f0=500.0 + 0.15 * slb.channel_tone_hz(c), first_channel=c)[0]
        if modes[c % len(modes)] in (slb.MODE_LSB, slb.MODE_CWR):
            x[c, :, 1] = -x[c, :, 1]                                      # put the tone on the lower sideband
    y, audio, gain = run_gpu(d, x)
    for c in range(channels):
        exp, a, g_, _ = best_oracle.rx_ssb_f32(d.oracle_params(modes[c % len(modes)]), x[c])
        err = np.abs(audio[c] - a); tol = audio_tolerance(a)
        assert np.all(err <= tol + 1e-9), (c, float(np.max(err / (tol + 1e-9))))
        check_int16(y[c], exp)


REJECTED = [(slb.MODE_CWR, 2350.0, -1), (slb.MODE_CW, 2350.0, 1), (slb.MODE_USB, 1500.0, -1), (slb.MODE_LSB, 1500.0, 1), (slb.MODE_DIG, 2000.0, -1), (slb.MODE_USB, 10000.0, 1)]


@pytest.mark.parametrize("path", [slb.RX_PATH_AUTO, slb.RX_PATH_FFT])
@pytest.mark.parametrize("mode,f0,sideband", REJECTED)
def test_output_dominated_by_a_rejected_tone(best_oracle, path, mode, f0, sideband):
    """A 0.25 FS tone the mode's filter REJECTS (the other sideband, or far outside the pass band): the output — noise in the pass band
    plus leakage — is 36 .. 42 dB below the input, so 1e-5 of the output is ~1e-7 of the input, the level of float32 rounding in any
    implementation of this chain. Both kernels are held to the plain bar, 1e-5 of the output: the FFT kernel (float32 like the oracle)
    sits at ~0.6 of it, the tensor-core kernel — an exact integer contraction with a 32-bit map — at what is left of the oracle's own
    float32 noise (round 1's 24-bit map needed an extra 2.5e-7 of the input here; the fourth map digit removed that term)."""
    frames = 1536 * 3
    x = slb.synth_iq(2, frames, f0=f0, sideband=sideband)
    d = slb.DspIf(2, chain=slb.CHAIN_RX_SSB_F32); d.set_rx_path(path)
    d.DSP_Set_Mode(mode)
    y, audio, gain = run_gpu(d, x)
    for c in range(2):
        exp, a, g_, _ = best_oracle.rx_ssb_f32(d.oracle_params(mode), x[c])
        rms_in = np.sqrt(np.mean((x[c].astype(np.float64) / 32768.0) ** 2) * 2)
        assert np.sqrt(np.mean(a.astype(np.float64) ** 2)) < 0.03 * rms_in           # the premise: a rejected tone
        err = np.abs(audio[c] - a); tol = audio_tolerance(a)
        assert np.all(err <= tol + 1e-12), float(np.max(err / tol))
        check_int16(y[c], exp)


def test_config1_one_second_snr(best_oracle):
    """BASELINE config 1: single channel, 48 kHz, 1 s tone + noise (48000 frames = 125 hops), SNR within 0.1 dB."""
    x = slb.synth_iq(1, 48000, f0=1000.0)
    d = slb.DspIf(1, chain=slb.CHAIN_RX_SSB_F32)
    y, audio, _ = run_gpu(d, x)
    exp, a, _, _ = best_oracle.rx_ssb_f32(d.oracle_params(), x[0])
    check_int16(y[0], exp)
    s_gpu, s_ref = snr_db(y[0, 4800:, 0], 1000.0), snr_db(exp[4800:, 0], 1000.0)
    assert s_ref > 20.0 and abs(s_gpu - s_ref) < 0.1, (s_gpu, s_ref)
    assert np.all(np.abs(audio[0] - a) <= audio_tolerance(a) + 1e-9)


def test_state_carries_across_calls(best_oracle):
    """The firmware processes a stream in blocks; results must not depend on how the stream is cut into calls."""
    C, T = 4, 1536 * 6
    x = slb.synth_iq(C, T)
    whole = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32)
    y_whole, _, _ = run_gpu(whole, x, want_audio=False)
    # tile-aligned cuts walk exactly the same arithmetic: bit-identical
    cut = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32)
    parts = [run_gpu(cut, np.ascontiguousarray(x[:, a:b]), want_audio=False)[0] for a, b in ((0, 1536), (1536, 4608), (4608, T))]
    assert np.array_equal(np.concatenate(parts, 1), y_whole)
    # hop-sized calls (the 8 x 48-frame firmware super-block) change the scan grouping only: same within 1 LSB
    hop = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32)
    parts = [run_gpu(hop, np.ascontiguousarray(x[:, a:a + 384]), want_audio=False)[0] for a in range(0, T, 384)]
    y_hop = np.concatenate(parts, 1)
    exp, _ = best_oracle.rx_ssb_f32_batch(hop.oracle_params(), x)
    check_int16(y_hop, exp)


def test_checkpoint_round_trip():
    C, T = 3, 1536 * 2
    x = slb.synth_iq(C, 2 * T)
    a = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32)
    run_gpu(a, np.ascontiguousarray(x[:, :T]), want_audio=False)
    snap = a.state_save()
    y1 = run_gpu(a, np.ascontiguousarray(x[:, T:]), want_audio=False)[0]
    b = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32)
    b.state_load(snap)
    y2 = run_gpu(b, np.ascontiguousarray(x[:, T:]), want_audio=False)[0]
    assert np.array_equal(y1, y2)


def test_host_bulk_path_equals_device_path():
    C, T = 70, 1536 * 4
    x = slb.synth_iq(C, T)
    y_dev = run_gpu(slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32), x, want_audio=False)[0]
    y_host = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32).rx_process(x)
    assert np.array_equal(y_dev, y_host)
    xp = torch.from_numpy(x).pin_memory(); yp = torch.empty_like(xp).pin_memory()
    slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32).rx_process_pinned(xp, yp)
    assert np.array_equal(yp.numpy(), y_dev)


def test_shards_equal_whole_and_full_size_properties(best_oracle):
    """BASELINE config 2 width (1024 channels): every shard of a 2/4/8-way split reproduces the 1-GPU result byte for
    byte (channels carry no cross-channel state), and a sample of channels is checked against the oracle."""
    C, T = 1024, 1536 * 5
    rng = np.random.Generator(np.random.PCG64(3))
    base = slb.synth_iq(16, T)
    x = np.ascontiguousarray(base[rng.integers(0, 16, C)] // np.int16(1) + rng.integers(-300, 300, (C, T, 2)).astype(np.int16))
    whole = run_gpu(slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32), x, want_audio=False)[0]
    assert np.array_equal(whole[..., 0], whole[..., 1])
    for world in (2, 8):
        for rank in (0, world - 1):
            lo, hi = slb.shard.shard_range(C, rank, world)
            part = run_gpu(slb.DspIf(hi - lo, chain=slb.CHAIN_RX_SSB_F32), np.ascontiguousarray(x[lo:hi]), want_audio=False)[0]
            assert np.array_equal(part, whole[lo:hi]), (world, rank)
    d = slb.DspIf(1, chain=slb.CHAIN_RX_SSB_F32)
    for c in rng.integers(0, C, 6):
        exp, _, _, _ = best_oracle.rx_ssb_f32(d.oracle_params(), x[c])
        check_int16(whole[c], exp)


def test_pass_chain_is_bit_exact_copy():
    x = slb.synth_iq(9, 1000)
    d = slb.DspIf(9, chain=slb.CHAIN_PASS)
    y = d.rx_process(torch.from_numpy(x).cuda()); torch.cuda.synchronize()
    assert np.array_equal(y.cpu().numpy(), x)
    assert np.array_equal(d.rx_process(x), x)


def test_bad_sizes_are_rejected():
    d = slb.DspIf(2, chain=slb.CHAIN_RX_SSB_F32)
    with pytest.raises(slb.SeleniteError):
        d.rx_process(torch.zeros((2, 400, 2), dtype=torch.int16, device="cuda"))
    with pytest.raises(slb.SeleniteError):
        d.DSP_Set_Mode(0x07)        # not an FT-817 mode byte (rxtx_if.h:33-43)
    assert d.kernel_launches() == 0


def am_signal(channels, frames, depth=0.6, f_mod=700.0, f_off=150.0, seed=5):
    """An AM carrier f_off away from the channel centre, modulated by a tone, plus noise: int16[channels][frames][2]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = np.arange(frames)
    out = np.empty((channels, frames, 2), np.int16)
    for c in range(channels):
        env = 0.2 * (1.0 + depth * np.cos(2 * np.pi * (f_mod + 10 * c) * n / 48000.0))
        z = env * np.exp(2j * np.pi * (f_off + 3 * c) * n / 48000.0) + 0.004 * (rng.standard_normal(frames) + 1j * rng.standard_normal(frames))
        out[c, :, 0] = np.rint(z.real * 32768); out[c, :, 1] = np.rint(z.imag * 32768)
    return out


def test_am_envelope_detector(best_oracle):
    """FT-817 mode byte 0x04 (rxtx_if.h:37): two-sided channel mask + arm_cmplx_mag_f32 instead of the real part, per
    channel, mixed with SSB channels in one launch."""
    C, T = 6, 1536 * 4
    x = am_signal(C, T)
    d = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32)
    modes = [slb.MODE_AM, slb.MODE_USB, slb.MODE_AM, slb.MODE_LSB, slb.MODE_AM, slb.MODE_AM]
    for c, m in enumerate(modes):
        d.DSP_Set_Mode(m, channel=c)
    y, audio, gain = run_gpu(d, x)
    for c, m in enumerate(modes):
        exp, a, g_, _ = best_oracle.rx_ssb_f32(d.oracle_params(m), x[c])
        tol = audio_tolerance(a)
        assert np.all(np.abs(audio[c] - a) <= tol + 1e-12), (c, float(np.max(np.abs(audio[c] - a) / (tol + 1e-12))))
        check_int16(y[c], exp)
    # the demodulated AM audio carries the modulating tone: envelope swing ~ depth around the carrier level
    a = audio[0][1536:]
    assert 0.5 < (a.max() - a.min()) / (2 * a.mean()) < 0.7


@pytest.mark.parametrize("fs", [96000])
@pytest.mark.parametrize("path", [slb.RX_PATH_AUTO, slb.RX_PATH_FFT])
def test_shipped_sample_rate(best_oracle, fs, path):
    """The firmware ships at 96 kHz (USB_DEVICE/Class/usbd_audio.h:46; geometry dsp_if.h:69-85). The float chain keeps its
    48-frame AGC block at every rate (slb_get_rx_f32_params says so and the oracle is fed from it), masks and biquads are
    designed for fs, the AGC release time is kept by scaling the decay with fs. (At 192 kHz — config 4's rate, served by the
    channelizer — the 129 taps cannot realise a 2.4 kHz pass band: the wanted tone sits on the filter's skirt, and the
    output-relative 1e-5 bar is ill-conditioned there for the FFT kernel and the tensor-core kernel alike, 1.3 x, the
    situation of test_output_dominated_by_a_rejected_tone; the TX chain is tested at both rates.)"""
    C, T = 6, 1536 * 3
    d = slb.DspIf(C, fs=fs, chain=slb.CHAIN_RX_SSB_F32); d.set_rx_path(path)
    p = d.rx_params()
    assert p.agc_block == 48 and abs(p.agc_decay - np.exp(-(48.0 / fs) / 0.3)) < 1e-7
    d.set_rx_params(p)                                                    # a get-modify-set round trip is accepted
    x = slb.synth_iq(C, T, fs=fs)
    modes = [slb.MODE_USB, slb.MODE_LSB, slb.MODE_DIG]
    for c in range(C):
        d.DSP_Set_Mode(modes[c % 3], channel=c)
        if modes[c % 3] == slb.MODE_LSB:
            x[c, :, 1] = -x[c, :, 1]
    y, audio, gain = run_gpu(d, x)
    for c in range(C):
        exp, a, g_, _ = best_oracle.rx_ssb_f32(d.oracle_params(modes[c % 3]), x[c])
        assert np.sqrt(np.mean(a.astype(np.float64) ** 2)) > 0.02         # the tone passes at this rate too (the 129 taps are a wider filter at 96 / 192 kHz)
        err = np.abs(audio[c] - a); tol = audio_tolerance(a)
        assert np.all(err <= tol + 1e-9), (c, float(np.max(err / (tol + 1e-9))))
        assert np.allclose(gain[c], g_, rtol=2e-5)
        check_int16(y[c], exp)
