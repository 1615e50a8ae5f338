"""The chains at config 5's width (8192 channels per GPU: 1024 channel groups on 148 persistent CTAs, the memory system saturated).
Parity proper lives in the per-chain test files at sizes the oracle finishes in seconds; this file is about what only shows at scale:
hand-over protocols between the roles of a CTA that can be overtaken when a load is slow (the RX-SSB-q15 kernel once hung here,
tests/test_gpu_rx_ssb_q15.py::test_config5_width_many_groups_per_cta_under_memory_load), 32-bit index overflow, groups of mixed
masks dealt to many CTAs. Each case compares two independent kernels of the library on the device (tensor-core kernel against the
FFT kernel) or the wide batch against a narrow one, and runs under a hard timeout."""
import numpy as np
import pytest

import selenite_lite_b200 as slb

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _int16_close(a, b):
    d = (a.to(torch.int32) - b.to(torch.int32)).abs()
    return int(d.max()), float((d > 0).float().mean())


@pytest.mark.timeout(240, method="thread")
def test_rx_ssb_f32_8192_channels_tensor_core_against_fft_kernel():
    C, T = 8192, 96000
    g = torch.Generator(device="cuda"); g.manual_seed(11)
    x = torch.randint(-12000, 12000, (C, T, 2), dtype=torch.int16, device="cuda", generator=g)
    modes = [slb.MODE_USB, slb.MODE_LSB, slb.MODE_CW, slb.MODE_DIG, slb.MODE_USB, slb.MODE_AM, slb.MODE_FM]
    outs = []
    for path in (slb.RX_PATH_AUTO, slb.RX_PATH_FFT):
        d = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32); d.set_rx_path(path)
        for m in range(1, len(modes)):
            for c in range(m, C, 64 * len(modes)):                                 # a sprinkle of every mode: mixed groups, mask reloads
                d.DSP_Set_Mode(modes[m], channel=c)
        y = None
        for rep in range(3):                                                        # a race needs a few launches
            y = d.rx_process(x, y)
        torch.cuda.synchronize()
        d = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32); d.set_rx_path(path)
        for m in range(1, len(modes)):
            for c in range(m, C, 64 * len(modes)):
                d.DSP_Set_Mode(modes[m], channel=c)
        outs.append(d.rx_process(x)); torch.cuda.synchronize()
    fm = torch.zeros(C, dtype=torch.bool, device="cuda"); fm[6::64 * len(modes)] = True       # FM has one kernel: the same on both paths
    mx, frac = _int16_close(outs[0][~fm], outs[1][~fm])
    assert mx <= 1 and frac < 0.02, (mx, frac)
    assert torch.equal(outs[0][fm], outs[1][fm])
    assert int(outs[0].abs().max()) > 1000                                          # the chains did run


@pytest.mark.timeout(240, method="thread")
def test_tx_ssb_f32_8192_channels_tensor_core_against_fft_kernel():
    C, T = 8192, 96000
    g = torch.Generator(device="cuda"); g.manual_seed(12)
    m = torch.randint(-8000, 8000, (C, T, 1), dtype=torch.int16, device="cuda", generator=g).expand(C, T, 2).contiguous()
    outs = []
    for path in (slb.RX_PATH_AUTO, slb.RX_PATH_FFT):
        d = slb.DspIf(C, chain=slb.CHAIN_TX_SSB_F32); d.set_rx_path(path)
        for c in range(1, C, 3):
            d.DSP_Set_Mode(slb.MODE_LSB, channel=c)
        for rep in range(3):
            d.tx_process(m)
        torch.cuda.synchronize()
        d = slb.DspIf(C, chain=slb.CHAIN_TX_SSB_F32); d.set_rx_path(path)
        for c in range(1, C, 3):
            d.DSP_Set_Mode(slb.MODE_LSB, channel=c)
        outs.append(d.tx_process(m)); torch.cuda.synchronize()
    mx, frac = _int16_close(outs[0], outs[1])
    assert mx <= 1 and frac < 0.02, (mx, frac)
    assert int(outs[0].abs().max()) > 1000


@pytest.mark.timeout(240, method="thread")
def test_chan64_512_streams_equal_the_same_streams_alone():
    S, T = 512, 192000 // 768 * 768
    g = torch.Generator(device="cuda"); g.manual_seed(13)
    x = torch.randint(-3000, 3000, (S, T, 2), dtype=torch.int16, device="cuda", generator=g)
    d = slb.DspIf(S, fs=192000, chain=slb.CHAIN_CHAN64_F32)
    y = torch.empty((S, 64, T // 64, 2), dtype=torch.int16, device="cuda")
    for rep in range(3):
        d.chan_process(x, y)
    torch.cuda.synchronize()
    d = slb.DspIf(S, fs=192000, chain=slb.CHAIN_CHAN64_F32); d.chan_process(x, y); torch.cuda.synchronize()
    for s0 in (0, 509):
        n = 3
        dn = slb.DspIf(n, fs=192000, chain=slb.CHAIN_CHAN64_F32)
        yn = torch.empty((n, 64, T // 64, 2), dtype=torch.int16, device="cuda")
        dn.chan_process(x[s0:s0 + n].contiguous(), yn); torch.cuda.synchronize()
        assert torch.equal(yn, y[s0:s0 + n]), s0


@pytest.mark.timeout(240, method="thread")
def test_firmware_api_at_8192_channels_equals_a_16_channel_context():
    """64 firmware milliseconds through the feeder (rings, 8-tick accumulation, chain, ring again) at 8192 channels: every channel hears
    what the same stream hears in a 16-channel context."""
    C, ticks = 8192, 64
    x16 = slb.synth_iq(16, 48 * ticks)
    x = np.ascontiguousarray(np.tile(x16, (C // 16, 1, 1)))
    d = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32)
    d.DSP_Init(); d.DSP_Set_RX()
    got, _ = d.feeder_run(adc=x)
    ref = slb.DspIf(16, chain=slb.CHAIN_RX_SSB_F32)
    ref.DSP_Init(); ref.DSP_Set_RX()
    exp, _ = ref.feeder_run(adc=x16)
    assert got.shape == x.shape
    assert np.array_equal(got[:16], exp) and np.array_equal(got[-16:], exp) and np.array_equal(got[4096:4112], exp)
    assert np.abs(exp).max() > 100


@pytest.mark.timeout(240, method="thread")
def test_host_path_of_a_wide_batch_is_cut_into_channel_blocks_and_equals_the_device_path():
    """4096 channels x 9216 frames through slb_rx_process_host with the default tile size: four blocks of 1024 channels x six slices
    of 1536 frames, against one device-resident call."""
    C, T = 4096, 1536 * 6
    g = torch.Generator(device="cuda"); g.manual_seed(14)
    xd = torch.randint(-12000, 12000, (C, T, 2), dtype=torch.int16, device="cuda", generator=g)
    for chain in (slb.CHAIN_RX_SSB_F32, slb.CHAIN_RX_SSB_Q15):
        d = slb.DspIf(C, chain=chain)
        y_dev = d.rx_process(xd); torch.cuda.synchronize()
        xp = xd.cpu().pin_memory(); yp = torch.empty_like(xp).pin_memory()
        h = slb.DspIf(C, chain=chain)
        h.rx_process_pinned(xp, yp)
        assert torch.equal(yp, y_dev.cpu()), chain
        if chain == slb.CHAIN_RX_SSB_Q15:
            assert bytes(h.state_save()) == bytes(d.state_save())                 # (the f32 header counts calls: six slices vs one)
        y_dev2 = d.rx_process(xd); torch.cuda.synchronize()                       # both contexts carry on alike
        h.rx_process_pinned(xp, yp)
        assert torch.equal(yp, y_dev2.cpu()), chain


@pytest.mark.timeout(300, method="thread")
def test_bench_shape_1024_channels_x_10_s_against_the_oracle_and_its_own_cuts(best_oracle):
    """The bench line's workload at full size (configs[1]: 1024 channels x 10 s = 480 000 frames per channel, one launch). Four
    channels spread over the batch are run through the oracle for the whole 10 s (int16 within 1 LSB on < 2 % of samples: no drift of
    the AGC / biquad state over 625 supertiles); channels that carry the same stream give the same bytes wherever they sit in the
    batch; and the stream cut into two calls at a 384-frame boundary that is NOT a supertile boundary equals the one call."""
    from test_gpu_rx_ssb_f32 import check_int16
    C, T = 1024, 480000
    base = slb.synth_iq(8, T)                                                       # eight different streams (tone + noise, PCG64 per channel)
    xd = torch.from_numpy(base).cuda().repeat(C // 8, 1, 1).contiguous()            # channel c carries stream c % 8
    d = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32)
    for c in range(1, C, 2):
        d.DSP_Set_Mode(slb.MODE_LSB, channel=c)                                    # odd streams demodulate the lower sideband
    y = d.rx_process(xd); torch.cuda.synchronize()
    y8 = y[:8]
    assert torch.equal(y.reshape(C // 8, 8, T, 2), y8.unsqueeze(0).expand(C // 8, 8, T, 2))
    yh = y8.cpu().numpy()
    for c in (0, 3, 5, 6):
        exp, _, _, _ = best_oracle.rx_ssb_f32(d.oracle_params(slb.MODE_LSB if c % 2 else slb.MODE_USB), base[c])
        check_int16(yh[c], exp)
    d2 = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_F32)
    for c in range(1, C, 2):
        d2.DSP_Set_Mode(slb.MODE_LSB, channel=c)
    cut = 384 * 601                                                                 # 300.5 supertiles
    ya = d2.rx_process(xd[:, :cut].contiguous()); yb = d2.rx_process(xd[:, cut:].contiguous()); torch.cuda.synchronize()
    assert torch.equal(ya, y[:, :cut]) and torch.equal(yb, y[:, cut:])


@pytest.mark.timeout(300, method="thread")
def test_config3_shape_1024_mic_channels_x_10_s_against_the_oracle(best_oracle):
    """TX-SSB-f32 at full size: 1024 mic channels x 10 s, four channels against the oracle for the whole stream, placement invariance."""
    from test_gpu_tx_ssb_f32 import check_int16
    C, T = 1024, 480000
    base = slb.synth_mic(8, T)
    xd = torch.from_numpy(base).cuda().repeat(C // 8, 1, 1).contiguous()
    d = slb.DspIf(C, chain=slb.CHAIN_TX_SSB_F32)
    for c in range(1, C, 2):
        d.DSP_Set_Mode(slb.MODE_LSB, channel=c)
    y = d.tx_process(xd); torch.cuda.synchronize()
    assert torch.equal(y.reshape(C // 8, 8, T, 2), y[:8].unsqueeze(0).expand(C // 8, 8, T, 2))
    yh = y[:8].cpu().numpy()
    for c in (0, 3, 5, 6):
        exp, _, _, _ = best_oracle.tx_ssb_f32(d.oracle_params(slb.MODE_LSB if c % 2 else slb.MODE_USB), base[c])
        check_int16(yh[c], exp)


@pytest.mark.timeout(300, method="thread")
def test_config4_shape_64_streams_x_10_s_against_the_oracle(best_oracle):
    """CHAN-64-f32 at full size: 64 wideband streams x 10 s at 192 kHz (625 tiles per stream chained by the look-back), two streams
    against the oracle for the whole 10 s, placement invariance."""
    from test_gpu_chan64_f32 import check_int16
    S, T = 64, 192000 * 10 // 768 * 768
    base = slb.synth_wideband(4, T)
    xd = torch.from_numpy(base).cuda().repeat(S // 4, 1, 1).contiguous()
    d = slb.DspIf(S, fs=192000, chain=slb.CHAIN_CHAN64_F32)
    y = d.chan_process(xd); torch.cuda.synchronize()
    assert torch.equal(y.reshape(S // 4, 4, 64, T // 64, 2), y[:4].unsqueeze(0).expand(S // 4, 4, 64, T // 64, 2))
    yh = y[:4].cpu().numpy()
    for s in (1, 2):
        exp, _, _, _ = best_oracle.chan_f32(d.oracle_params(), base[s])
        check_int16(yh[s], exp)


@pytest.mark.timeout(300, method="thread")
def test_q15_chain_1024_channels_x_10_s_bit_exact(best_oracle):
    """RX-SSB-q15 at the bench width and length: bit-exact against the oracle on four channels over the whole 10 s, placement invariance."""
    C, T = 1024, 480000
    base = slb.synth_iq(8, T)
    xd = torch.from_numpy(base).cuda().repeat(C // 8, 1, 1).contiguous()
    d = slb.DspIf(C, chain=slb.CHAIN_RX_SSB_Q15)
    y = d.rx_process(xd); torch.cuda.synchronize()
    assert torch.equal(y.reshape(C // 8, 8, T, 2), y[:8].unsqueeze(0).expand(C // 8, 8, T, 2))
    yh = y[:8].cpu().numpy()
    for c in (0, 2, 5, 7):
        exp = best_oracle.rx_ssb_q15(d.oracle_params(slb.MODE_USB), base[c])[0]
        assert np.array_equal(yh[c], exp), c
