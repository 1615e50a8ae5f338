"""GPU parity of the fused TX-SSB-f32 chain (BASELINE config 3: mic audio -> SSB modulation -> I/Q), called through the
C ABI, against the oracle and the committed golden vectors. Tolerances as for RX (tests/test_gpu_rx_ssb_f32.py):
the pre-ALC complex baseband within 1e-5 * max(|z[n]|, super-block rms), the int16 I/Q within 1 LSB on < 2 % of samples."""
import os

import numpy as np
import pytest

import selenite_lite_b200 as slb
from test_golden import GOLD, iq_tolerance

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def run_gpu(d, x, want_dbg=True):
    C, T = x.shape[0], x.shape[1]
    xd = torch.from_numpy(x).cuda()
    iq = torch.zeros((C, T, 2), dtype=torch.float32, device="cuda") if want_dbg else None
    gain = torch.zeros((C, T // 48), dtype=torch.float32, device="cuda") if want_dbg else None
    d.set_debug_taps(iq, gain)
    y = d.tx_process(xd)
    torch.cuda.synchronize()
    d.set_debug_taps(None, None)
    return y.cpu().numpy(), (iq.cpu().numpy() if want_dbg else None), (gain.cpu().numpy() if want_dbg else None)


def check_int16(y, exp):
    d = np.abs(y.astype(np.int32) - exp.astype(np.int32))
    assert d.max() <= 1, "int16 output differs by %d LSB" % d.max()
    assert np.mean(d > 0) < 0.02, "%.2f %% of samples differ" % (100 * np.mean(d > 0))


@pytest.mark.parametrize("name,mode", [("usb", slb.MODE_USB), ("lsb", slb.MODE_LSB)])
def test_golden_vectors(name, mode):
    g = np.load(os.path.join(GOLD, "tx_ssb_f32.npz"))
    d = slb.DspIf(1, chain=slb.CHAIN_TX_SSB_F32)
    d.DSP_Set_Mode(mode)
    assert np.array_equal(d.mask(mode), g["tx_%s_mask" % name])
    y, iq, gain = run_gpu(d, g["tx_%s_in" % name][None])
    ref_iq = g["tx_%s_iq" % name]
    err = np.abs(iq[0] - ref_iq); tol = iq_tolerance(ref_iq)
    assert np.all(err <= tol + 1e-9), "worst baseband error %.2f x tolerance" % np.max(err / (tol + 1e-9))
    assert np.allclose(gain[0], g["tx_%s_gain" % name], rtol=2e-5)
    check_int16(y[0], g["tx_%s_out" % name])


@pytest.mark.parametrize("channels,frames", [(1, 384), (3, 768), (7, 1920), (33, 4608), (130, 3072)])
def test_vs_oracle_ragged_shapes(best_oracle, channels, frames):
    x = slb.synth_mic(channels, frames)
    d = slb.DspIf(channels, chain=slb.CHAIN_TX_SSB_F32)
    modes = [slb.MODE_USB, slb.MODE_LSB, slb.MODE_CW, slb.MODE_DIG]
    for c in range(channels):
        d.DSP_Set_Mode(modes[c % len(modes)], channel=c)
    y, iq, gain = run_gpu(d, x)
    for c in range(channels):
        exp, z, g_, _ = best_oracle.tx_ssb_f32(d.oracle_params(modes[c % len(modes)]), x[c])
        err = np.abs(iq[c] - z); tol = iq_tolerance(z)
        assert np.all(err <= tol + 1e-9), (c, float(np.max(err / (tol + 1e-9))))
        assert np.allclose(gain[c], g_, rtol=2e-5)
        check_int16(y[c], exp)


def test_worst_case_inputs_for_the_integer_accumulators(best_oracle):
    """Mic streams that drive the int32 accumulators of the digit products as far as int16 data can — the rail at -32768, at +32767,
    a full-scale tone in the pass-band, a rail-to-rail square wave — through both TX kernels, all default masks. A full-scale input
    the mode's filter rejects by > 80 dB (DC) leaves only the oracle's own float32 rounding noise: held to -120 dB of the input there."""
    T = 768 * 4
    n = np.arange(T)
    pats = [np.full(T, -32768, np.int16), np.full(T, 32767, np.int16),
            np.clip(np.rint(np.sin(2 * np.pi * 1200.0 * n / 48000.0) * 32767), -32768, 32767).astype(np.int16),
            np.where((n // 16) % 2 == 0, 32767, -32768).astype(np.int16)]
    x = np.stack([np.stack([p_, p_], axis=1) for p_ in pats])
    C = len(pats)
    for mode in (slb.MODE_USB, slb.MODE_LSB, slb.MODE_CW, slb.MODE_DIG):
        for path in (slb.RX_PATH_AUTO, slb.RX_PATH_FFT):
            d = slb.DspIf(C, chain=slb.CHAIN_TX_SSB_F32); d.set_rx_path(path); d.DSP_Set_Mode(mode)
            y, iq, gain = run_gpu(d, x)
            for c in range(C):
                exp, z, g_, _ = best_oracle.tx_ssb_f32(d.oracle_params(mode), x[c])
                err = np.abs(iq[c] - z)
                if float(np.max(np.abs(z[384:]))) < 1e-4:
                    assert np.max(err) <= 1e-6, (mode, path, c, float(np.max(err)))
                else:
                    tol = iq_tolerance(z)
                    assert np.all(err <= tol + 1e-9), (mode, path, c, float(np.max(err / (tol + 1e-9))))
                dd = np.abs(y[c].astype(np.int32) - exp.astype(np.int32))
                assert dd.max() <= 1 and np.mean(dd > 0) < 0.10, (mode, path, c, int(dd.max()), float(np.mean(dd > 0)))


def test_only_the_left_channel_is_used(best_oracle):
    """The codec routes the microphone to both ADC channels in TX (codec_if.c:304-306); the chain reads L."""
    x = slb.synth_mic(2, 1536)
    x2 = x.copy(); x2[:, :, 1] = 12345
    a = run_gpu(slb.DspIf(2, chain=slb.CHAIN_TX_SSB_F32), x, want_dbg=False)[0]
    b = run_gpu(slb.DspIf(2, chain=slb.CHAIN_TX_SSB_F32), x2, want_dbg=False)[0]
    assert np.array_equal(a, b)


def test_config3_width_state_carry_and_shards(best_oracle):
    """BASELINE config 3 width: 1024 mic channels. Calls cut at tile boundaries are bit-identical to one call; a shard
    equals the same rows of the whole; a sample of channels is checked against the oracle."""
    C, T = 1024, 1536 * 4
    base = slb.synth_mic(16, T)
    rng = np.random.Generator(np.random.PCG64(5))
    x = np.ascontiguousarray(base[rng.integers(0, 16, C)])
    x[:, :, 0] += rng.integers(-200, 200, (C, T)).astype(np.int16); x[:, :, 1] = x[:, :, 0]
    whole = run_gpu(slb.DspIf(C, chain=slb.CHAIN_TX_SSB_F32), x, want_dbg=False)[0]
    cut = slb.DspIf(C, chain=slb.CHAIN_TX_SSB_F32)
    parts = [run_gpu(cut, np.ascontiguousarray(x[:, a:b]), want_dbg=False)[0] for a, b in ((0, 1536), (1536, T))]
    assert np.array_equal(np.concatenate(parts, 1), whole)
    lo, hi = slb.shard.shard_range(C, 3, 8)
    part = run_gpu(slb.DspIf(hi - lo, chain=slb.CHAIN_TX_SSB_F32), np.ascontiguousarray(x[lo:hi]), want_dbg=False)[0]
    assert np.array_equal(part, whole[lo:hi])
    d = slb.DspIf(1, chain=slb.CHAIN_TX_SSB_F32)
    for c in rng.integers(0, C, 5):
        exp, _, _, _ = best_oracle.tx_ssb_f32(d.oracle_params(), x[c])
        check_int16(whole[c], exp)


def test_host_path_and_1ms_cadence(best_oracle):
    """slb_tx_process_host equals the device path; behind the firmware API (DSP_In_Buff_Write of 48-frame blocks in a
    TX-chain context) the ring delivers the modulated stream delayed by one 384-frame super-block."""
    C, T = 5, 384 * 6
    x = slb.synth_mic(C, T)
    y_dev = run_gpu(slb.DspIf(C, chain=slb.CHAIN_TX_SSB_F32), x, want_dbg=False)[0]
    assert np.array_equal(slb.DspIf(C, chain=slb.CHAIN_TX_SSB_F32).tx_process(x), y_dev)
    d = slb.DspIf(C, chain=slb.CHAIN_TX_SSB_F32)
    d.DSP_Init(); d.DSP_Set_TX()
    got = []
    for b in range(T // 48):
        d.DSP_In_Buff_Write(np.ascontiguousarray(x[:, 48 * b:48 * (b + 1)]).reshape(C, 96))
        got.append(d.DSP_In_Buff_Read(192).reshape(C, 48, 2))
    got = np.concatenate(got, 1)
    # find where the modulated stream starts in the ring output (ring delay + one super-block of chain latency)
    flat = got[0].reshape(-1, 2); ref = y_dev[0]
    starts = [s for s in range(300, 800) if np.array_equal(flat[s:s + 600], ref[:600])]
    assert starts, "modulated stream not found behind the ring"
    with pytest.raises(slb.SeleniteError):
        d.rx_process(torch.zeros((C, 384, 2), dtype=torch.int16, device="cuda"))    # wrong direction for this context


def test_tensor_core_and_fft_kernels_agree_and_hand_over(best_oracle):
    """TX on the tensor cores (sl_tx_ssb_tc.cu, the default) against the FFT kernel (slb_set_rx_path FFT) on the same input,
    a stream cut into calls served alternately by the two (raw tail, ALC envelope and the hand-over counter cross the
    switch), several channel groups per CTA with mask reloads (forced 3-CTA grid), and the time-sliced host path."""
    C, T = 44, 1536 * 4 + 384
    x = slb.synth_mic(C, T)
    modes = [slb.MODE_USB, slb.MODE_LSB, slb.MODE_DIG, slb.MODE_USB, slb.MODE_CW]

    def make(path=slb.RX_PATH_AUTO):
        d = slb.DspIf(C, chain=slb.CHAIN_TX_SSB_F32); d.set_rx_path(path)
        for c in range(C):
            d.DSP_Set_Mode(modes[c % 5], channel=c)
        return d
    y_tc, iq_tc, g_tc = run_gpu(make(), x)
    y_fft, iq_fft, g_fft = run_gpu(make(slb.RX_PATH_FFT), x)
    dd = np.abs(y_tc.astype(np.int32) - y_fft.astype(np.int32))
    assert dd.max() <= 1 and np.mean(dd > 0) < 0.02
    for c in range(0, C, 7):
        exp, z, g_, _ = best_oracle.tx_ssb_f32(make().oracle_params(modes[c % 5]), x[c])
        tol = iq_tolerance(z)
        assert np.all(np.abs(iq_tc[c] - z) <= tol + 1e-9) and np.all(np.abs(iq_fft[c] - z) <= tol + 1e-9), c
        assert np.allclose(g_tc[c], g_, rtol=2e-5)
        check_int16(y_tc[c], exp)
    # alternate kernels between calls
    d = make(); xd = torch.from_numpy(x).cuda(); parts = []
    cuts = [0, 1536, 1536 + 384, 1536 * 3, T]
    for i in range(4):
        d.set_rx_path(slb.RX_PATH_FFT if i % 2 else slb.RX_PATH_AUTO)
        parts.append(d.tx_process(xd[:, cuts[i]:cuts[i + 1]].contiguous()).cpu().numpy())
    ya = np.concatenate(parts, axis=1)
    for c in (0, 13, 43):
        exp, _, _, _ = best_oracle.tx_ssb_f32(make().oracle_params(modes[c % 5]), x[c])
        check_int16(ya[c], exp)
    # several groups per CTA, mask reloads; and the host path cut into time slices
    os.environ["SELENITE_B200_TC_GRID"] = "3"; os.environ["SELENITE_B200_SLICE_BYTES"] = str(C * 4 * 1536)
    try:
        y3 = run_gpu(make(), x, want_dbg=False)[0]
        y_host = make().tx_process(x)
    finally:
        del os.environ["SELENITE_B200_TC_GRID"]; del os.environ["SELENITE_B200_SLICE_BYTES"]
    assert np.array_equal(y3, y_tc) and np.array_equal(y_host, y_tc)
    # channel blocks x time slices (the cut of wide batches): blocks of 16 channels (16 + 16 + 12)
    os.environ["SELENITE_B200_SLICE_BYTES"] = str(16 * 4 * 1536); os.environ["SELENITE_B200_SLICE_CHANNELS"] = "16"
    try:
        y_tiles = make().tx_process(x)
    finally:
        del os.environ["SELENITE_B200_SLICE_BYTES"]; del os.environ["SELENITE_B200_SLICE_CHANNELS"]
    assert np.array_equal(y_tiles, y_tc)


@pytest.mark.parametrize("fs", [96000, 192000])
def test_shipped_sample_rate(best_oracle, fs):
    """96 kHz is what the firmware ships with (usbd_audio.h:46); the TX chain keeps its 48-frame ALC block at every rate."""
    C, T = 5, 1536 * 2
    d = slb.DspIf(C, fs=fs, chain=slb.CHAIN_TX_SSB_F32)
    p = d.tx_params()
    assert p.alc_block == 48 and abs(p.alc_decay - np.exp(-(48.0 / fs) / 0.1)) < 1e-7
    d.set_tx_params(p)
    x = slb.synth_mic(C, T, fs=fs)
    modes = [slb.MODE_USB, slb.MODE_LSB]
    for c in range(C):
        d.DSP_Set_Mode(modes[c % 2], channel=c)
    y, iq, gain = run_gpu(d, x)
    for c in range(C):
        exp, z, g_, _ = best_oracle.tx_ssb_f32(d.oracle_params(modes[c % 2]), x[c])
        err = np.abs(iq[c] - z); tol = iq_tolerance(z)
        assert np.all(err <= tol + 1e-9), (c, float(np.max(err / (tol + 1e-9))))
        assert np.allclose(gain[c], g_, rtol=2e-5)
        check_int16(y[c], exp)
