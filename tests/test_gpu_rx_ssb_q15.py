"""GPU parity of the fused RX-SSB-q15 chain (the all-integer receive chain: fir_q15 Hilbert pair on the integer tensor
cores -> saturating add/sub -> q15 AGC), called through the C ABI, against the oracle and the committed golden vectors.
Every stage is integer arithmetic, so every comparison is BIT-EXACT: pre-AGC audio, per-block gain word and output."""
import os

import numpy as np
import pytest

import selenite_lite_b200 as slb
from test_golden import GOLD, q15_params

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def run_gpu(d, x, want_dbg=True):
    C, T = x.shape[0], x.shape[1]
    xd = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    audio = torch.zeros((C, T), dtype=torch.int16, device="cuda") if want_dbg else None
    gain = torch.zeros((C, T // 48), dtype=torch.int32, device="cuda") if want_dbg else None
    d.set_q15_debug_taps(audio, gain)
    y = d.rx_process(xd)
    torch.cuda.synchronize()
    d.set_q15_debug_taps(None, None)
    return y.cpu().numpy(), (audio.cpu().numpy() if want_dbg else None), (gain.cpu().numpy().view(np.uint32) if want_dbg else None)


@pytest.mark.parametrize("name,mode", [("usb", slb.MODE_USB), ("lsb", slb.MODE_LSB), ("sat", slb.MODE_USB)])
def test_golden_vectors(name, mode):
    g = np.load(os.path.join(GOLD, "rx_ssb_q15.npz"))
    d = slb.DspIf(1, chain=slb.CHAIN_RX_SSB_Q15)
    d.DSP_Set_Mode(mode)
    p = d.rx_q15_params()
    assert np.array_equal(np.array(p.taps_i[:], np.int16), g["q15_taps_i"]) and np.array_equal(np.array(p.taps_q[:], np.int16), g["q15_taps_q"])
    y, audio, gain = run_gpu(d, g["q15_%s_in" % name][None])
    assert np.array_equal(audio[0], g["q15_%s_audio" % name])
    assert np.array_equal(gain[0], g["q15_%s_gain" % name])
    assert np.array_equal(y[0], g["q15_%s_out" % name])


@pytest.mark.parametrize("channels,frames", [(1, 48), (3, 96), (8, 1488), (9, 4800), (33, 9600), (130, 3072)])
def test_vs_oracle_ragged_shapes(best_oracle, channels, frames):
    x = slb.synth_iq(channels, frames)
    d = slb.DspIf(channels, chain=slb.CHAIN_RX_SSB_Q15)
    modes = [slb.MODE_USB, slb.MODE_LSB, slb.MODE_CW, slb.MODE_CWR, slb.MODE_DIG]
    for c in range(channels):
        d.DSP_Set_Mode(modes[c % len(modes)], channel=c)
    y, audio, gain = run_gpu(d, x)
    for c in range(channels):
        exp, a, g_, _ = best_oracle.rx_ssb_q15(d.oracle_params(modes[c % len(modes)]), x[c])
        assert np.array_equal(audio[c], a), (c, int(np.argmax(audio[c] != a)))
        assert np.array_equal(gain[c], g_), c
        assert np.array_equal(y[c], exp), c


def test_every_saturation_point(best_oracle, rng):
    """Hot taps (4x, clipped) and full-scale input: the FIR's __SSAT, arm_add_q15's saturation, arm_abs_q15(-32768) and
    arm_scale_q15's saturation all fire; the split-byte tensor-core FIR must still equal the 64-bit accumulator."""
    C, T = 16, 4800
    d = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_Q15)
    p = d.rx_q15_params()
    for k in range(64):
        p.taps_i[k] = int(np.clip(4 * p.taps_i[k], -32768, 32767)); p.taps_q[k] = int(np.clip(4 * p.taps_q[k], -32768, 32767))
    p.taps_i[0] = -32768; p.taps_q[63] = 32767; p.taps_i[31] = 32767; p.taps_q[32] = -32768
    p.agc_window = 32; p.agc_floor = 1; p.agc_gmax_q15 = 255 << 15
    d.set_rx_q15_params(p)
    x = rng.integers(-32768, 32768, (C, T, 2)).astype(np.int16)
    x[0, 100:300] = -32768; x[1, 100:300] = 32767; x[2, 100:300, 0] = -32768; x[2, 100:300, 1] = 32767; x[3, 1000:2000] = 0
    x[4] = (x[4] // 4096).astype(np.int16)                   # a quiet channel: the gain limit and the floor come into play
    y, audio, gain = run_gpu(d, x)
    sat_seen = 0
    for c in range(C):
        exp, a, g_, _ = best_oracle.rx_ssb_q15(d.oracle_params(slb.MODE_USB), x[c])
        assert np.array_equal(audio[c], a), c
        assert np.array_equal(gain[c], g_), c
        assert np.array_equal(y[c], exp), c
        sat_seen += int(np.sum(np.abs(a.astype(np.int32)) >= 32767))
    assert sat_seen > 100


def test_streaming_state_across_calls(best_oracle):
    """Two bulk calls == one: the carried raw tail and peak window are the whole state. Long enough (and enough channels
    few) that the launch is cut into several time segments, each re-deriving its peak window."""
    C, T = 2, 48 * 2000
    x = slb.synth_iq(C, T)
    d = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_Q15)
    y1, _, _ = run_gpu(d, x[:, :48 * 700], want_dbg=False)
    y2, _, _ = run_gpu(d, x[:, 48 * 700:], want_dbg=False)
    whole = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_Q15)
    yw, _, _ = run_gpu(whole, x, want_dbg=False)
    assert np.array_equal(np.concatenate([y1, y2], 1), yw)
    exp = best_oracle.rx_ssb_q15(d.oracle_params(slb.MODE_USB), x[1])[0]
    assert np.array_equal(yw[1], exp)


def test_host_path_checkpoint_and_ring(best_oracle):
    C = 5
    x = slb.synth_iq(C, 48 * 40)
    d = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_Q15)
    y = d.rx_process(x[:, :48 * 20])                          # numpy -> slb_rx_process_host
    snap = d.state_save()
    y2 = d.rx_process(x[:, 48 * 20:])
    exp = np.stack([best_oracle.rx_ssb_q15(d.oracle_params(slb.MODE_USB), x[c])[0] for c in range(C)])
    assert np.array_equal(np.concatenate([y, y2], 1), exp)
    e = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_Q15)
    e.state_load(snap)
    assert np.array_equal(e.rx_process(x[:, 48 * 20:]), y2)
    # behind the firmware's 1 ms cadence: each 48-frame block is demodulated in the call that delivers it, then rides the ring
    f = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_Q15)
    p = slb.DspIf(C, chain=slb.CHAIN_PASS)
    for b in range(40):
        f.DSP_In_Buff_Write(x[:, 48 * b:48 * (b + 1)].reshape(C, -1))
        p.DSP_In_Buff_Write(exp[:, 48 * b:48 * (b + 1)].reshape(C, -1))
        assert np.array_equal(f.DSP_In_Buff_Read(192), p.DSP_In_Buff_Read(192))


def test_full_size_properties():
    """BASELINE configs[1] width (1024 channels): shards equal the whole, and the output is L = R everywhere."""
    C, T = 1024, 48 * 500
    x = slb.synth_iq(8, T)
    x = np.ascontiguousarray(np.tile(x, (C // 8, 1, 1)))
    d = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_Q15)
    y, _, _ = run_gpu(d, x, want_dbg=False)
    assert np.array_equal(y[..., 0], y[..., 1])
    assert np.array_equal(y[:8], y[8:16]) and np.array_equal(y[:8], y[-8:])
    h = slb.DspIf(C // 2, chain=slb.CHAIN_RX_SSB_Q15)
    yh, _, _ = run_gpu(h, x[C // 2:], want_dbg=False)
    assert np.array_equal(yh, y[C // 2:])


def test_tcgen05_kernel_equals_the_mma_sync_kernel_bit_for_bit():
    """The chain has two kernels: sl_rx_q15_tc.cu (tcgen05, the default when the taps and the AGC window allow it) and the
    mma.sync kernel of sl_rx_ssb_q15.cu (SELENITE_B200_Q15_PATH=legacy). Same output, audio, gain words AND carried state
    (raw tail, peak window by age) on ragged shapes: one block, a half supertile, several groups per CTA, both sidebands;
    and a stream cut into calls that alternate between the two."""
    import os
    import torch

    def run(C, x, legacy, cuts=None, grid=None):
        if legacy:
            os.environ["SELENITE_B200_Q15_PATH"] = "legacy"
        if grid:
            os.environ["SELENITE_B200_TC_GRID"] = str(grid)
        try:
            d = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_Q15)
            for c in range(1, C, 3):
                d.DSP_Set_Mode(slb.MODE_LSB, channel=c)
            T = x.shape[1]
            audio = torch.zeros((C, T), dtype=torch.int16, device="cuda"); gain = torch.zeros((C, T // 48), dtype=torch.int32, device="cuda")
            d.set_q15_debug_taps(audio, gain)
            xd = torch.from_numpy(x).cuda()
            if cuts is None:
                y = d.rx_process(xd).cpu().numpy()
            else:
                parts = []
                for i in range(len(cuts) - 1):
                    if (i % 2 == 1) != legacy:
                        os.environ["SELENITE_B200_Q15_PATH"] = "legacy"
                    else:
                        os.environ.pop("SELENITE_B200_Q15_PATH", None)
                    parts.append(d.rx_process(xd[:, cuts[i]:cuts[i + 1]].contiguous()).cpu().numpy())
                y = np.concatenate(parts, axis=1)
            torch.cuda.synchronize()
            return y, audio.cpu().numpy(), gain.cpu().numpy(), bytes(d.state_save())
        finally:
            os.environ.pop("SELENITE_B200_Q15_PATH", None); os.environ.pop("SELENITE_B200_TC_GRID", None)

    rng = np.random.Generator(np.random.PCG64(21))
    for C, T, grid in ((1, 48, None), (5, 48 * 9, None), (19, 768 * 2 + 48 * 5, 2), (300, 768 * 3, None)):
        x = slb.synth_iq(C, T)
        x[:, ::97] = rng.integers(-32768, 32768, x[:, ::97].shape)                  # rail-to-rail samples in between
        a = run(C, x, legacy=False, grid=grid); b = run(C, x, legacy=True)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]), (C, T)
        assert a[3] == b[3], (C, T)                                               # tail + peak window + sidebands
    C, T = 11, 768 * 2 + 48 * 7
    x = slb.synth_iq(C, T)
    whole = run(C, x, legacy=True)
    cut = run(C, x, legacy=False, cuts=[0, 48, 48 * 14, 768 + 48 * 3, T])
    assert np.array_equal(whole[0], cut[0]) and whole[3] == cut[3]


def test_config5_width_many_groups_per_cta_under_memory_load():
    """8192 channels (config 5's shard per GPU) x 2 s: 1024 groups on 148 CTAs = seven groups per CTA with the memory system saturated.
    The tcgen05 kernel once hung here (a set's wait on the two-slot peak hand-over barrier could be overtaken by the other set whenever
    the load at a group's start outlasted a supertile's MMAs; tests at small widths never saw it). Output and carried state must equal
    the mma.sync kernel's, compared on the device; a broken hand-over now traps instead of hanging (sl_tc_common.cuh)."""
    import os
    import torch
    C, T = 8192, 96000
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    x = torch.randint(-20000, 20000, (C, T, 2), dtype=torch.int16, device="cuda", generator=g)
    outs = []
    for legacy in (False, True):
        if legacy:
            os.environ["SELENITE_B200_Q15_PATH"] = "legacy"
        try:
            d = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_Q15)
            for rep in range(3 if not legacy else 1):                              # the race needed a few launches
                y = d.rx_process(x)
            torch.cuda.synchronize()
            if not legacy:                                                         # same stream position for the comparison
                d = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_Q15); y = d.rx_process(x); torch.cuda.synchronize()
            outs.append((y, bytes(d.state_save())))
        finally:
            os.environ.pop("SELENITE_B200_Q15_PATH", None)
    assert torch.equal(outs[0][0], outs[1][0])
    assert outs[0][1] == outs[1][1]


def test_host_path_cut_into_time_slices_is_bit_exact():
    """slb_rx_process_host for the q15 chain cuts the batch in time like the f32 chains: forced down to 1536-frame slices, a stream
    of 6 slices and a 5-block remainder equals the device path bit for bit, state included."""
    import os
    import torch
    C, T = 13, 1536 * 6 + 48 * 5
    x = slb.synth_iq(C, T)
    d_dev = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_Q15)
    y_dev = d_dev.rx_process(torch.from_numpy(x).cuda()).cpu().numpy()
    os.environ["SELENITE_B200_SLICE_BYTES"] = str(C * 4 * 1536)
    try:
        d_host = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_Q15)
        y_host = d_host.rx_process(x)
    finally:
        del os.environ["SELENITE_B200_SLICE_BYTES"]
    assert np.array_equal(y_host, y_dev)
    assert bytes(d_host.state_save()) == bytes(d_dev.state_save())
    # channel blocks x time slices (the cut of wide batches): blocks of 5 channels (5 + 5 + 3)
    os.environ["SELENITE_B200_SLICE_BYTES"] = str(5 * 4 * 1536); os.environ["SELENITE_B200_SLICE_CHANNELS"] = "5"
    try:
        d_tiles = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_Q15)
        y_tiles = d_tiles.rx_process(x)
    finally:
        del os.environ["SELENITE_B200_SLICE_BYTES"]; del os.environ["SELENITE_B200_SLICE_CHANNELS"]
    assert np.array_equal(y_tiles, y_dev)
    assert bytes(d_tiles.state_save()) == bytes(d_dev.state_save())


def _q15_audio_biquad():
    """Two df1 sections in q15, postShift 1: a 150 Hz high-pass and a 2.6 kHz low-pass (arm_biquad_cascade_df1_init_q15 layout
    {b0, 0, b1, b2, a1, a2}, feedback signs as CMSIS stores them)."""
    from scipy import signal
    secs = []
    for kind, fc in (("highpass", 150.0), ("lowpass", 2600.0)):
        b, a = signal.butter(2, fc, kind, fs=48000)
        c = np.round(np.array([b[0], b[1], b[2], -a[1], -a[2]]) / 2 * 32768).astype(np.int64)
        secs += [c[0], 0, c[1], c[2], c[3], c[4]]
    return np.clip(np.array(secs), -32768, 32767).astype(np.int16), 1


@pytest.mark.parametrize("legacy", [False, True])
def test_optional_integer_biquad_stage(best_oracle, rng, legacy, monkeypatch):
    """SURVEY Appendix B's arm_biquad_cascade_df1_q15 between mixer and AGC, switched on through the chain parameters: the chain then
    runs as three kernels (fused FIR + mixer, the channel-parallel integer biquad, AGC + scale) and stays BIT-EXACT — filtered audio,
    gain words, output, over calls of ragged length (carried filter state and peak window), through the 1 ms firmware API and a
    checkpoint, for both FIR kernels; hot input makes the biquad's own saturation fire."""
    if legacy:
        monkeypatch.setenv("SELENITE_B200_Q15_PATH", "legacy")
    C, T = 11, 48 * 90
    x = slb.synth_iq(C, T)
    x[3] = np.clip(x[3].astype(np.int32) * 5, -32768, 32767).astype(np.int16)          # saturating channel
    coeffs, ps = _q15_audio_biquad()
    d = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_Q15)
    p = d.rx_q15_params(); p.bq_stages, p.bq_postshift = 2, ps
    for k in range(12):
        p.bq_coeffs[k] = int(coeffs[k])
    d.set_rx_q15_params(p)
    modes = [slb.MODE_USB, slb.MODE_LSB, slb.MODE_CW]
    for c in range(C):
        d.DSP_Set_Mode(modes[c % 3], channel=c)
    outs, auds, gains = [], [], []
    for a, b in ((0, 48 * 7), (48 * 7, 48 * 40), (48 * 40, T)):                      # ragged cuts: filter state and peak window carry
        y, au, g = run_gpu(d, np.ascontiguousarray(x[:, a:b])); outs.append(y); auds.append(au); gains.append(g)
    y, au, g = np.concatenate(outs, 1), np.concatenate(auds, 1), np.concatenate(gains, 1)
    for c in range(C):
        prm = d.oracle_params(modes[c % 3])
        assert prm["bq_stages"] == 2
        exp, a_, g_, _ = best_oracle.rx_ssb_q15(prm, x[c])
        assert np.array_equal(au[c], a_), (c, int(np.argmax(au[c] != a_)))
        assert np.array_equal(g[c], g_), c
        assert np.array_equal(y[c], exp), c
    # the filter does something: 50 Hz hum that the FIR pair lets through is attenuated in the filtered audio
    plain = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_Q15)
    _, au_plain, _ = run_gpu(plain, x)
    assert not np.array_equal(au_plain[0], au[0])
    # 1 ms API + checkpoint: the same stream block by block, saved and restored half way
    e = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_Q15); e.set_rx_q15_params(p)
    for c in range(C):
        e.DSP_Set_Mode(modes[c % 3], channel=c)
    ring_in = []
    for t in range(40):
        if t == 20:
            snap = e.state_save(); e = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_Q15); e.set_rx_q15_params(p); e.state_load(snap)
        e.DSP_In_Buff_Write(x[:, 48 * t:48 * t + 48].reshape(C, -1)); ring_in.append(e.DSP_In_Buff_Read(192))
    f = slb.DspIf(C, chain=slb.CHAIN_PASS)                                            # the ring alone, fed with the oracle chain's output
    for t in range(40):
        f.DSP_In_Buff_Write(y[:, 48 * t:48 * t + 48].reshape(C, -1))
        assert np.array_equal(f.DSP_In_Buff_Read(192), ring_in[t]), t
