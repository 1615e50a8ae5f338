"""GPU parity of the tensor-core form of the RX-SSB-f32 chain (sl_rx_ssb_tc.cu), through the C ABI.

The kernel evaluates the overlap-save filter as the exact 129-tap integer FIR it is (tcgen05.mma kind::i8), so against
the float32-FFT oracle it is held to the same bars as the FFT kernel: audio within 1e-5 * max(|ref|, frame rms), int16
output within 1 LSB on < 2 % of samples. Extra here: the two kernels against each other, the hand-over of the carried
state between them, per-channel kernel choice with mixed modes, short groups, the half supertile at the end of a stream."""
import numpy as np
import pytest

import selenite_lite_b200 as slb
from test_golden import audio_tolerance
from test_gpu_rx_ssb_f32 import check_int16, run_gpu

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def one_lsb(a, b):
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    return d.max() <= 1 and np.mean(d > 0) < 0.02


@pytest.mark.parametrize("channels,frames", [(1, 384), (7, 768), (8, 1152), (9, 3840), (150, 1536), (300, 2304)])
def test_both_kernels_meet_the_oracle_and_each_other(best_oracle, channels, frames):
    """384 = half a supertile, 1152 = one and a half; 1 / 7 / 9 channels = short groups (rows repeat a channel, unstored);
    150 / 300 = more groups than one per SM and the smaller-group rule (fewer than 8 channels per SM)."""
    x = slb.synth_iq(channels, frames)
    d_tc = slb.DspIf(channels, chain=slb.CHAIN_RX_SSB_F32)
    d_fft = slb.DspIf(channels, chain=slb.CHAIN_RX_SSB_F32); d_fft.set_rx_path(slb.RX_PATH_FFT)
    y_tc, a_tc, g_tc = run_gpu(d_tc, x)
    y_fft, a_fft, g_fft = run_gpu(d_fft, x)
    assert one_lsb(y_tc, y_fft)
    worst = 0.0
    for c in range(0, channels, max(1, channels // 6)):
        exp, a, g_, _ = best_oracle.rx_ssb_f32(d_tc.oracle_params(), x[c])
        tol = audio_tolerance(a)
        e_tc = float(np.max(np.abs(a_tc[c] - a) / tol)); e_fft = float(np.max(np.abs(a_fft[c] - a) / tol))
        assert e_tc <= 1.0 and e_fft <= 1.0, (c, e_tc, e_fft)
        worst = max(worst, e_tc)
        check_int16(y_tc[c], exp)
        assert np.allclose(g_tc[c], g_, rtol=2e-5)
    assert worst < 0.5        # the exact FIR sits well inside the tolerance (the oracle's own float32 FFT noise dominates)


def test_state_hands_over_between_the_two_kernels(best_oracle):
    """One stream cut into four calls served alternately by the tensor-core and the FFT kernel: raw tail, biquad state,
    AGC envelope and the FFT kernel's hand-over counter all cross the switch (DESIGN.md §3.2 carried state)."""
    C, T = 12, 1536 * 8
    x = slb.synth_iq(C, T)
    d = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32)
    cuts = [0, 1536 * 2, 1536 * 2 + 384, 1536 * 5, T]
    xd = torch.from_numpy(x).cuda(); parts = []
    for i in range(4):
        d.set_rx_path(slb.RX_PATH_FFT if i % 2 else slb.RX_PATH_AUTO)
        parts.append(d.rx_process(xd[:, cuts[i]:cuts[i + 1]].contiguous()).cpu().numpy())
    y = np.concatenate(parts, axis=1)
    for c in (0, 5, 11):
        exp, _, _, _ = best_oracle.rx_ssb_f32(d.oracle_params(), x[c])
        check_int16(y[c], exp)
    # and a stream cut into calls on the tensor-core kernel alone equals the uncut stream bit for bit
    d1 = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32); d2 = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32)
    whole = d1.rx_process(xd).cpu().numpy()
    cut = np.concatenate([d2.rx_process(xd[:, a:b].contiguous()).cpu().numpy() for a, b in ((0, 768 * 3), (768 * 3, 768 * 10), (768 * 10, T))], axis=1)
    assert np.array_equal(whole, cut)


def test_kernel_choice_is_per_channel(best_oracle):
    """Modes scattered over the channels: every SSB / CW / DIG channel is served by the tensor-core kernel in groups of its
    own mask, AM channels by the FFT kernel, all in one call; a channel's result does not depend on its neighbours."""
    C, T = 41, 768 * 5
    modes = [slb.MODE_USB, slb.MODE_AM, slb.MODE_LSB, slb.MODE_CW, slb.MODE_USB, slb.MODE_DIG, slb.MODE_CWR, slb.MODE_PKT]
    # most channels get a tone inside the pass band of their own mode, every fifth one a tone its mode REJECTS (the other sideband):
    # the same plain 1e-5 bar holds for both (test_output_dominated_by_a_rejected_tone)
    f_in = {slb.MODE_USB: 1000.0, slb.MODE_LSB: -1200.0, slb.MODE_CW: 700.0, slb.MODE_CWR: -650.0, slb.MODE_DIG: 2000.0, slb.MODE_PKT: 1500.0, slb.MODE_AM: 150.0}
    def tone_of(c):
        f = f_in[modes[c % len(modes)]]
        if c % 5 == 4 and modes[c % len(modes)] != slb.MODE_AM:
            f = -f                                                        # out of band: the opposite sideband
        return f
    x = np.concatenate([slb.synth_iq(1, T, f0=abs(tone_of(c)) + 3 * c, sideband=1 if tone_of(c) > 0 else -1, first_channel=c) for c in range(C)])
    d = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32)
    for c in range(C):
        d.DSP_Set_Mode(modes[c % len(modes)], channel=c)
    y, audio, gain = run_gpu(d, x)
    for c in range(C):
        m = modes[c % len(modes)]
        exp, a, g_, _ = best_oracle.rx_ssb_f32(d.oracle_params(m), x[c])
        tol = audio_tolerance(a)
        assert np.all(np.abs(audio[c] - a) <= tol + 1e-12), (c, m)
        check_int16(y[c], exp)
    # the same channels alone, one context per mode: bit-identical to the mixed batch
    for m in (slb.MODE_USB, slb.MODE_CW):
        idx = [c for c in range(C) if modes[c % len(modes)] == m]
        ds = slb.DspIf(len(idx), chain=slb.CHAIN_RX_SSB_F32); ds.DSP_Set_Mode(m)
        ys = run_gpu(ds, np.ascontiguousarray(x[idx]), want_audio=False)[0]
        assert np.array_equal(ys, y[idx]), m


def test_weak_and_full_scale_signals(best_oracle):
    """The integer FIR is exact, so its error does not depend on the signal level: a -60 dBFS stream and a stream that
    touches both int16 rails (the byte split's extremes) meet the same relative tolerance."""
    T = 768 * 6
    base = slb.synth_iq(2, T).astype(np.int32)
    weak = (base // 250).astype(np.int16)
    loud = np.clip(base * 4, -32768, 32767).astype(np.int16)
    loud[0, 100] = (-32768, 32767); loud[1, 101] = (32767, -32768)
    for x in (weak, loud):
        d = slb.DspIf(2, chain=slb.CHAIN_RX_SSB_F32)
        y, audio, gain = run_gpu(d, x)
        for c in range(2):
            exp, a, _, _ = best_oracle.rx_ssb_f32(d.oracle_params(), x[c])
            assert np.all(np.abs(audio[c] - a) <= audio_tolerance(a) + 1e-12)
            check_int16(y[c], exp)


def test_worst_case_inputs_for_the_integer_accumulators(best_oracle):
    """Inputs that drive the int32 accumulators of the digit products as far as int16 data can: both rails held at -32768 (every
    high byte -128, every low byte 0), at +32767 (high 127, low 255: the unsigned low-byte operand at its maximum), a full-scale
    complex tone inside the pass-band (the filter's gain on top), and full-scale alternation at the band edge. Both kernels against
    the oracle at the plain bars, all default masks."""
    T = 768 * 4
    n = np.arange(T)
    pats = []
    pats.append(np.full((T, 2), -32768, np.int16))
    pats.append(np.full((T, 2), 32767, np.int16))
    tone = np.exp(2j * np.pi * 1000.0 * n / 48000.0)
    pats.append(np.stack([np.clip(np.rint(tone.real * 32767), -32768, 32767), np.clip(np.rint(tone.imag * 32767), -32768, 32767)], axis=1).astype(np.int16))
    sq = np.where((n // 8) % 2 == 0, 32767, -32768).astype(np.int16)              # 3 kHz square wave, rail to rail, on both rails
    pats.append(np.stack([sq, np.roll(sq, 4)], axis=1))
    x = np.stack(pats)
    C = len(pats)
    for mode in (slb.MODE_USB, slb.MODE_LSB, slb.MODE_CW, slb.MODE_DIG):
        for path in (slb.RX_PATH_AUTO, slb.RX_PATH_FFT):
            d = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32); d.set_rx_path(path); d.DSP_Set_Mode(mode)
            y, audio, gain = run_gpu(d, x)
            for c in range(C):
                exp, a, _, _ = best_oracle.rx_ssb_f32(d.oracle_params(mode), x[c])
                if float(np.max(np.abs(a[384:]))) < 1e-4:                             # (after the filter has filled)
                    # a full-scale input the mode's filter rejects by > 80 dB (DC; the tone in the other sideband; the square wave outside a
                    # narrow filter): what is left of it in the oracle is its own float32 rounding noise, and a bar
                    # relative to THAT output means nothing (a 512-point float32 FFT of a full-scale signal carries ~ log2 (512) * 2^-24 =
                    # 5e-7 of it; the integer FIR's own DC leak is the 32-bit map's, 3.6e-9). Held to -120 dB of the input instead.
                    assert np.max(np.abs(audio[c] - a)) <= 1e-6, (mode, path, c, float(np.max(np.abs(audio[c] - a))), float(np.max(np.abs(a))))
                else:
                    assert np.all(np.abs(audio[c] - a) <= audio_tolerance(a) + 1e-12), (mode, path, c, float(np.max(np.abs(audio[c] - a) / (audio_tolerance(a) + 1e-12))))
                # (constant-envelope inputs park the AGC output on the target, 0.25 * 32768 = an integer: the truncating pack then turns
                # float differences of 1e-7 into +-1 LSB on more samples than the 2 % of a noisy signal — the LSB bound itself holds)
                dd = np.abs(y[c].astype(np.int32) - exp.astype(np.int32))
                assert dd.max() <= 1 and np.mean(dd > 0) < 0.10, (mode, path, c, int(dd.max()), float(np.mean(dd > 0)))


def test_caller_mask_that_is_no_fir_stays_on_the_fft_kernel(best_oracle):
    """slb_set_mask with a spectrum that is not the DFT of a 129-tap filter: the channel keeps the FFT kernel (the oracle's
    circular convolution is reproduced, not approximated by a truncated FIR)."""
    C, T = 9, 1536 * 3
    x = slb.synth_iq(C, T)
    d = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32)
    rng = np.random.Generator(np.random.PCG64(11))
    mask = d.mask(slb.MODE_USB).copy()
    mask *= (1.0 + 0.05 * rng.standard_normal(mask.shape)).astype(np.float32)     # ragged pass-band: impulse response fills all 512 taps
    d.set_mask(slb.MODE_USB, mask)
    y, audio, gain = run_gpu(d, x)
    prm = d.oracle_params()
    for c in (0, 8):
        exp, a, _, _ = best_oracle.rx_ssb_f32(prm, x[c])
        assert np.all(np.abs(audio[c] - a) <= audio_tolerance(a) + 1e-12)
        check_int16(y[c], exp)


def test_several_groups_per_cta_and_mask_reloads(best_oracle):
    """More channel groups than CTAs (what 8192 channels per GPU, BASELINE config 5, look like): a CTA walks several groups
    back to back — pipelines and carry barriers run on across the group boundary, the carried state is re-read at every
    group start — and reloads the operand planes when the next group has another mask. Forced here with a 3-CTA grid
    (SELENITE_B200_TC_GRID, a profiling knob of the launcher) so that a few channels suffice; also 2048 channels on the
    full grid against single-mode contexts."""
    import os
    C, T = 44, 768 * 3 + 384
    modes = [slb.MODE_USB, slb.MODE_LSB, slb.MODE_DIG, slb.MODE_USB, slb.MODE_CW]
    f_in = {slb.MODE_USB: 1000.0, slb.MODE_LSB: -1200.0, slb.MODE_CW: 700.0, slb.MODE_DIG: 2000.0}
    x = np.concatenate([slb.synth_iq(1, T, f0=abs(f_in[modes[c % 5]]) + 2 * c, sideband=1 if f_in[modes[c % 5]] > 0 else -1, first_channel=c) for c in range(C)])
    d = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32)
    for c in range(C):
        d.DSP_Set_Mode(modes[c % 5], channel=c)
    full = run_gpu(d, x, want_audio=False)[0]
    os.environ["SELENITE_B200_TC_GRID"] = "3"
    try:
        d3 = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32)
        for c in range(C):
            d3.DSP_Set_Mode(modes[c % 5], channel=c)
        y3, audio, gain = run_gpu(d3, x)
        # two calls: the state written at the end of every group is picked up again
        d3b = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32)
        for c in range(C):
            d3b.DSP_Set_Mode(modes[c % 5], channel=c)
        xd = torch.from_numpy(x).cuda()
        y3b = np.concatenate([d3b.rx_process(xd[:, :768].contiguous()).cpu().numpy(), d3b.rx_process(xd[:, 768:].contiguous()).cpu().numpy()], axis=1)
    finally:
        del os.environ["SELENITE_B200_TC_GRID"]
    assert np.array_equal(y3, full) and np.array_equal(y3b, full)
    for c in range(0, C, 5):
        exp, a, _, _ = best_oracle.rx_ssb_f32(d.oracle_params(modes[c % 5]), x[c])
        assert np.all(np.abs(audio[c] - a) <= audio_tolerance(a) + 1e-12), c
        check_int16(full[c], exp)
    # full grid, more groups than SMs: 2048 channels = 256 groups of 8, two masks interleaved in blocks of 100 channels
    C2, T2 = 2048, 768 * 2
    base = slb.synth_iq(8, T2)
    x2 = np.ascontiguousarray(base[np.arange(C2) % 8])
    lsb = (np.arange(C2) // 100) % 2 == 1
    x2[lsb, :, 1] = -x2[lsb, :, 1]
    d2 = slb.DspIf(C2, chain=slb.CHAIN_RX_SSB_F32)
    for c in np.nonzero(lsb)[0]:
        d2.DSP_Set_Mode(slb.MODE_LSB, channel=int(c))
    y2 = run_gpu(d2, x2, want_audio=False)[0]
    for m, sel in ((slb.MODE_USB, ~lsb), (slb.MODE_LSB, lsb)):
        ds = slb.DspIf(8, chain=slb.CHAIN_RX_SSB_F32); ds.DSP_Set_Mode(m)
        xs = base.copy()
        if m == slb.MODE_LSB:
            xs[:, :, 1] = -xs[:, :, 1]
        ys = run_gpu(ds, xs, want_audio=False)[0]
        idx = np.nonzero(sel)[0]
        assert np.array_equal(y2[idx], ys[idx % 8]), m


def test_host_path_cut_into_many_time_slices(best_oracle):
    """slb_rx_process_host cuts the batch in time (8 MiB tiles by default); with the slice forced down to 1536 frames a
    stream of 7 slices (the last one short) must equal the device path bit for bit — carried state across slices, strided
    copies, three staging slots in flight — for the tensor-core kernel and, with AM channels mixed in, the FFT kernel."""
    import os
    C, T = 20, 1536 * 6 + 384
    x = slb.synth_iq(C, T)
    def make():
        d = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32)
        for c in (3, 11):
            d.DSP_Set_Mode(slb.MODE_AM, channel=c)
        return d
    y_dev = run_gpu(make(), x, want_audio=False)[0]
    os.environ["SELENITE_B200_SLICE_BYTES"] = str(C * 4 * 1536)
    try:
        y_host = make().rx_process(x)
        xp = torch.from_numpy(x).pin_memory(); yp = torch.empty_like(xp).pin_memory()
        make().rx_process_pinned(xp, yp)
    finally:
        del os.environ["SELENITE_B200_SLICE_BYTES"]
    assert np.array_equal(y_host, y_dev) and np.array_equal(yp.numpy(), y_dev)
    exp, _, _, _ = best_oracle.rx_ssb_f32(make().oracle_params(), x[0])
    check_int16(y_host[0], exp)
    # wide batches are cut into channel blocks as well (config 5: blocks of 1024 channels): forced down to blocks of 7 channels
    # (three blocks, the last of 6; AM channels in two of them) x 1536-frame slices = 21 tiles, state per channel across them
    os.environ["SELENITE_B200_SLICE_BYTES"] = str(7 * 4 * 1536); os.environ["SELENITE_B200_SLICE_CHANNELS"] = "7"
    try:
        d = make()
        y_tiles = d.rx_process(x)
        y_next = d.rx_process(x)                                                  # and the context carries on from there
    finally:
        del os.environ["SELENITE_B200_SLICE_BYTES"]; del os.environ["SELENITE_B200_SLICE_CHANNELS"]
    assert np.array_equal(y_tiles, y_dev)
    d = make(); xd = torch.from_numpy(x).cuda(); d.rx_process(xd)
    assert np.array_equal(y_next, d.rx_process(xd).cpu().numpy())


def test_am_channels_on_the_tensor_cores(best_oracle):
    """AM (FT-817 mode byte 0x04) has its own tensor-core kernel (sl_rx_am_tc.cu: both rails of the channel filter by
    tcgen05.mma, envelope + biquad in the epilogue; opt-in with SELENITE_B200_AM_PATH=tc because the FFT kernel is faster for
    AM): an all-AM batch is one launch, meets the oracle, agrees with the FFT
    kernel, hands its state over to it and back, and several groups per CTA / host slices reproduce it bit for bit."""
    import os
    from test_gpu_rx_ssb_f32 import am_signal
    C, T = 21, 768 * 4 + 384
    x = am_signal(C, T)

    def make(path=slb.RX_PATH_AUTO):
        # the AM tensor-core kernel is opt-in (it measures slower than the FFT kernel): the knob is read when a context is created
        os.environ["SELENITE_B200_AM_PATH"] = "tc"
        try:
            d = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32)
        finally:
            del os.environ["SELENITE_B200_AM_PATH"]
        d.DSP_Set_Mode(slb.MODE_AM); d.set_rx_path(path)
        return d
    d = make(); n0 = d.kernel_launches()
    y, audio, gain = run_gpu(d, x)
    assert d.kernel_launches() - n0 == 1
    y_fft, a_fft, _ = run_gpu(make(slb.RX_PATH_FFT), x)
    assert one_lsb(y, y_fft)
    prm = d.oracle_params(slb.MODE_AM)
    for c in range(0, C, 4):
        exp, a, g_, _ = best_oracle.rx_ssb_f32(prm, x[c])
        tol = audio_tolerance(a)
        assert np.all(np.abs(audio[c] - a) <= tol + 1e-12), (c, float(np.max(np.abs(audio[c] - a) / tol)))
        assert np.allclose(gain[c], g_, rtol=2e-5)
        check_int16(y[c], exp)
    d2 = make(); xd = torch.from_numpy(x).cuda(); parts = []
    cuts = [0, 768, 768 + 384, 768 * 3, T]
    for i in range(4):
        d2.set_rx_path(slb.RX_PATH_FFT if i % 2 else slb.RX_PATH_AUTO)
        parts.append(d2.rx_process(xd[:, cuts[i]:cuts[i + 1]].contiguous()).cpu().numpy())
    ya = np.concatenate(parts, axis=1)
    for c in (0, 20):
        exp, _, _, _ = best_oracle.rx_ssb_f32(prm, x[c])
        check_int16(ya[c], exp)
    os.environ["SELENITE_B200_TC_GRID"] = "2"; os.environ["SELENITE_B200_SLICE_BYTES"] = str(C * 4 * 1536)
    try:
        y2 = run_gpu(make(), x, want_audio=False)[0]
        y_host = make().rx_process(x)
    finally:
        del os.environ["SELENITE_B200_TC_GRID"]; del os.environ["SELENITE_B200_SLICE_BYTES"]
    assert np.array_equal(y2, y) and np.array_equal(y_host, y)
