"""GPU parity of the firmware ring behind the drop-in API: the batched entry points and the single-channel functions
with the firmware's own names, against the unmodified dsp_if.c (when built) and its port. int16: bit-exact."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib
import selenite_lite_b200 as slb
from selenite_lite_b200 import _lib

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def oracle_ring(fs):
    try:
        r = oracle_lib.RefRing(fs)
        return r, (r.in_write, lambda nb: r.in_read(nb), r.out_write, r.out_read, lambda w: r.ptrs(w), r.out_mute)
    except (FileNotFoundError, OSError):
        a, b = oracle_lib.PortRing(fs), oracle_lib.PortRing(fs)
        return None, (a.in_write, a.in_read, b.out_write, b.out_read, lambda w: (b if w else a).ptrs(), b.out_mute)


@pytest.mark.parametrize("fs", [48000, 96000, 192000])
def test_batched_ring_bit_exact(fs, rng):
    Cn = 5; hw = fs // 1000 * 2
    d = slb.DspIf(Cn, fs=fs, chain=slb.CHAIN_PASS)
    d.DSP_Init()
    rings = [oracle_ring(fs)[1] for _ in range(Cn)]
    sched = ["w", "r"] * 20 + ["w", "w", "r"] * 30 + ["w", "r", "r"] * 30 + list(rng.choice(["w", "r"], 100))
    for step, op in enumerate(sched):
        if op == "w":
            blk = rng.integers(-32768, 32768, (Cn, hw)).astype(np.int16)
            d.DSP_In_Buff_Write(blk, hw); d.DSP_Out_Buff_Write(blk, hw * 2)
            for c in range(Cn):
                rings[c][0](blk[c]); rings[c][2](blk[c])
        else:
            got_rx = d.DSP_In_Buff_Read(hw * 2); got_tx = d.DSP_Out_Buff_Read(hw)
            for c in range(Cn):
                assert np.array_equal(got_rx[c], rings[c][1](hw * 2)), ("rx", step, c)
                assert np.array_equal(got_tx[c], rings[c][3](hw)), ("tx", step, c)
        assert d.ring_ptrs(0) == tuple(rings[0][4](0)) and d.ring_ptrs(1) == tuple(rings[0][4](1)), step
    d.DSP_Out_Buff_Mute()
    for c in range(Cn):
        rings[c][5]()
    got = d.DSP_Out_Buff_Read(hw)
    assert not got.any() and np.array_equal(got[0], rings[0][3](hw))


def test_firmware_named_dropin_single_channel(rng):
    """DSP_In_Buff_Write(uint16_t*, uint16_t) etc. exactly as Core/Inc/dsp_if.h:42-51 declares them."""
    lib = _lib.load()
    _, o = oracle_ring(48000)
    lib.DSP_Init()
    assert lib.slb_dropin_status() == 0
    for b in range(60):
        blk = rng.integers(-32768, 32768, 96).astype(np.int16)
        lib.DSP_In_Buff_Write(blk.ctypes.data, 96); o[0](blk)
        out = np.zeros(96, np.int16)
        lib.DSP_In_Buff_Read(out.ctypes.data, 192)
        assert np.array_equal(out, o[1](192)), b
        lib.DSP_Out_Buff_Write(blk.ctypes.data, 192); o[2](blk)
        lib.DSP_Out_Buff_Read(out.ctypes.data, 96)
        assert np.array_equal(out, o[3](96)), b
    assert lib.slb_dropin_status() == 0


def test_chain_behind_the_1ms_cadence(best_oracle):
    """RX-SSB-f32 inserted in DSP_In_Buff_Write: 48-frame blocks in, the ring out, against
    [oracle chain over the stream, delayed by one 384-frame super-block] pushed through the oracle ring."""
    Cn, nblk = 3, 8 * 12
    x = slb.synth_iq(Cn, nblk * 48)
    d = slb.DspIf(Cn, chain=slb.CHAIN_RX_SSB_F32)
    d.DSP_Init()
    exp_chain, _ = best_oracle.rx_ssb_f32_batch(d.oracle_params(), x)
    delayed = np.concatenate([np.zeros((Cn, 384, 2), np.int16), exp_chain], 1)
    rings = [oracle_ring(48000)[1] for _ in range(Cn)]
    worst = 0
    for b in range(nblk):
        d.DSP_In_Buff_Write(np.ascontiguousarray(x[:, 48 * b:48 * b + 48]).reshape(Cn, 96), 96)
        got = d.DSP_In_Buff_Read(192)
        for c in range(Cn):
            rings[c][0](delayed[c, 48 * b:48 * b + 48].reshape(-1))
            e = rings[c][1](192)
            worst = max(worst, int(np.max(np.abs(got[c].astype(np.int32) - e.astype(np.int32)))))
    assert worst <= 1


def _oracle_rings(n, fs=48000):
    try:
        return [oracle_lib.RefRing(fs) for _ in range(n)]
    except FileNotFoundError:
        return [oracle_lib.PortRing(fs) for _ in range(n)]


def test_per_channel_drift_compensation():
    """producers and consumers of different channels tick at different rates. Every channel carries its
    own pointers on the device and must follow its own instance of the firmware ring (one unmodified dsp_if.c image per
    channel) through both slip branches: frames delivered, pointers and ring contents, bit for bit."""
    C, ticks, hw = 6, 700, 96
    d = slb.DspIf(C, chain=slb.CHAIN_PASS)
    rings = _oracle_rings(C)
    rng = np.random.Generator(np.random.PCG64(11))
    # channel 0: nominal; 1: producer drops every 37th tick; 2: consumer drops every 41st; 3: producer drops often;
    # 4: consumer drops often; 5: both irregular
    def fires(c, t):
        w = not ((c == 1 and t % 37 == 36) or (c == 3 and t % 9 == 8) or (c == 5 and rng.random() < 0.04))
        r = not ((c == 2 and t % 41 == 40) or (c == 4 and t % 11 == 10) or (c == 5 and rng.random() < 0.04))
        return w, r
    def ptrs_of(r, which):
        return tuple(r.ptrs(which)) if isinstance(r, oracle_lib.RefRing) else tuple(r.ptrs())
    slips = np.zeros(C, np.int64)
    for t in range(ticks):
        blk = rng.integers(-30000, 30000, (C, hw)).astype(np.int16)
        act = np.array([fires(c, t) for c in range(C)])
        d.DSP_In_Buff_Write_Ch(blk, act[:, 0].astype(np.uint8))
        got = d.DSP_In_Buff_Read_Ch(2 * hw, act[:, 1].astype(np.uint8))
        for c in range(C):
            wr_before = ptrs_of(rings[c], 0)[2]
            if act[c, 0]:
                rings[c].in_write(blk[c])
                slips[c] += int((ptrs_of(rings[c], 0)[2] - wr_before) % 384 != 48)
            exp = rings[c].in_read(2 * hw) if act[c, 1] else np.zeros(hw, np.int16)
            assert np.array_equal(got[c], exp), (t, c)
            assert d.ring_ptrs_channel(c, 0) == ptrs_of(rings[c], 0), (t, c)
    assert slips[0] <= 3 and slips[3] > 5 and slips[4] > 5            # the nominal channel only slips while arming; drifting ones keep slipping
    # TX ring with per-channel consumers
    e = slb.DspIf(C, chain=slb.CHAIN_PASS)
    rings = _oracle_rings(C)
    for t in range(200):
        blk = rng.integers(-30000, 30000, (C, hw)).astype(np.int16)
        act = np.array([(t % (7 + c)) != 0 for c in range(C)])
        e.DSP_Out_Buff_Write_Ch(blk)
        got = e.DSP_Out_Buff_Read_Ch(hw, act.astype(np.uint8))
        for c in range(C):
            rings[c].out_write(blk[c])
            exp = rings[c].out_read(hw) if act[c] else np.zeros(hw, np.int16)
            assert np.array_equal(got[c], exp), (t, c)
    # checkpoint carries the per-channel pointers
    snap = e.state_save()
    f = slb.DspIf(C, chain=slb.CHAIN_PASS); f.state_load(snap)
    blk = rng.integers(-30000, 30000, (C, hw)).astype(np.int16)
    e.DSP_Out_Buff_Write_Ch(blk); f.DSP_Out_Buff_Write_Ch(blk)
    assert np.array_equal(e.DSP_Out_Buff_Read_Ch(hw), f.DSP_Out_Buff_Read_Ch(hw))
    assert all(e.ring_ptrs_channel(c, 1) == f.ring_ptrs_channel(c, 1) for c in range(C))


@pytest.mark.parametrize("chain,ticks", [(slb.CHAIN_PASS, 61), (slb.CHAIN_RX_SSB_Q15, 50), (slb.CHAIN_RX_SSB_F32, 48), (slb.CHAIN_TX_SSB_F32, 40)])
def test_stream_feeder_equals_the_per_tick_calls(chain, ticks, rng):
    """slb_feeder_run == `ticks` rounds of {Out_Buff_Read, In_Buff_Write, Out_Buff_Write, In_Buff_Read}
    (the I2S callback, then AUDIO_CMD_PLAY / AUDIO_CMD_RECORD), including the start-up slips, chain latency and every
    piece of carried state: two feeder runs back to back, then per-tick calls on both contexts must keep agreeing."""
    C, B = 7, 48
    adc = slb.synth_iq(C, 2 * ticks * B + 16 * B)
    pc = rng.integers(-20000, 20000, (C, 2 * ticks * B + 16 * B, 2)).astype(np.int16)
    a = slb.DspIf(C, chain=chain); b = slb.DspIf(C, chain=chain)
    # reference run on `a`: one call per event
    exp_in, exp_dac = [], []
    for t in range(2 * ticks):
        blk = slice(t * B, (t + 1) * B)
        exp_dac.append(a.DSP_Out_Buff_Read(2 * B))
        a.DSP_In_Buff_Write(adc[:, blk].reshape(C, -1))
        a.AUDIO_AudioCmd(pc[:, blk].reshape(C, -1), 4 * B, 2)                                   # AUDIO_CMD_PLAY
        exp_in.append(a.AUDIO_AudioCmd(np.zeros((C, 2 * B), np.int16), 4 * B, 4))                # AUDIO_CMD_RECORD
    exp_in = np.stack(exp_in, 1).reshape(C, -1, 2); exp_dac = np.stack(exp_dac, 1).reshape(C, -1, 2)
    n = ticks * B
    in1, dac1 = b.feeder_run(adc[:, :n], pc[:, :n])
    in2, dac2 = b.feeder_run(adc[:, n:2 * n], pc[:, n:2 * n])
    assert np.array_equal(np.concatenate([dac1, dac2], 1), exp_dac)
    assert np.array_equal(np.concatenate([in1, in2], 1), exp_in)
    assert a.ring_ptrs(0) == b.ring_ptrs(0) and a.ring_ptrs(1) == b.ring_ptrs(1)
    for t in range(2 * ticks, 2 * ticks + 16):                                                   # and the two contexts stay in step
        blk = slice(t * B, (t + 1) * B)
        for d in (a, b):
            d.DSP_In_Buff_Write(adc[:, blk].reshape(C, -1))
        assert np.array_equal(a.DSP_In_Buff_Read(4 * B), b.DSP_In_Buff_Read(4 * B))


def test_firmware_names_i2s_callbacks_and_audiocmd():
    """The single-channel drop-in surface beyond dsp_if.h: i2s_buff + HAL_I2SEx_TxRx{Half,}CpltCallback (dsp_if.c:32, :50-67)
    and AUDIO_AudioCmd_FS (usbd_audio_if.c:179-202) drive the same global context as DSP_*."""
    lib = _lib.load()
    lib.DSP_Init()
    # I2S_Buff_TypeDef at 48 kHz (dsp_if.h:69-79): rx[192]; tx[192] — the layout a dsp_if.h consumer compiled for this rate sees
    buf_t = (C.c_uint16 * 192) * 2
    i2s = buf_t.in_dll(lib, "i2s_buff")
    ring, fns = oracle_ring(48000)
    in_write, in_read, out_write, out_read, ptrs, mute = fns
    rng = np.random.Generator(np.random.PCG64(3))
    for t in range(30):
        half = t & 1
        rx = rng.integers(-30000, 30000, 96).astype(np.int16); pkt = rng.integers(-30000, 30000, 96).astype(np.int16)
        C.memmove(C.addressof(i2s[0]) + half * 192, rx.ctypes.data, 192)
        (lib.HAL_I2SEx_TxRxCpltCallback if half else lib.HAL_I2SEx_TxRxHalfCpltCallback)(None)
        tx = np.frombuffer(bytes(i2s[1]), np.int16)[half * 96:(half + 1) * 96]
        exp_tx = out_read(96); in_write(rx)
        assert np.array_equal(tx, exp_tx), t
        lib.AUDIO_AudioCmd_FS(pkt.ctypes.data, 192, 2); out_write(pkt)
        got = np.zeros(96, np.int16); lib.AUDIO_AudioCmd_FS(got.ctypes.data, 192, 4)
        assert np.array_equal(got, in_read(192)), t
    assert lib.slb_dropin_status() == 0


@pytest.mark.parametrize("fs", [48000, 96000])
def test_cw_sidetone_at_the_firmware_hook(best_oracle, rng, fs):
    """'mix CW tone to speaker signal here' (dsp_if.c:218): DSP_Out_Buff_Read adds the side-tone to L and R of every keyed channel.
    Against the oracle composition (arm_sin_f32 -> arm_scale_f32 -> arm_float_to_q15 -> arm_add_q15) applied to the oracle ring's
    output, bit for bit, through key-down / key-up sequences, with samples near the rails (the add saturates)."""
    Cn = 6; B = fs // 1000; hw = 2 * B
    d = slb.DspIf(Cn, fs=fs, chain=slb.CHAIN_PASS); d.DSP_Init()
    d.DSP_Set_Sidetone(700, 0.3)
    rings = [oracle_ring(fs)[1] for _ in range(Cn)]
    cnt = [0] * Cn
    keys = np.zeros(Cn, bool)
    for step in range(60):
        if step % 7 == 0:
            keys = rng.random(Cn) < 0.5
            if step == 28: keys[:] = False
            d.DSP_Key(keys if keys.any() else None)
        blk = rng.integers(-32768, 32768, (Cn, hw)).astype(np.int16)
        d.DSP_Out_Buff_Write(blk, hw * 2)
        got = d.DSP_Out_Buff_Read(hw)
        for c in range(Cn):
            rings[c][2](blk[c])
            e = rings[c][3](hw).reshape(B, 2)
            e, cnt[c] = best_oracle.sidetone_mix(e, cnt[c], bool(keys[c]), 700, fs, 0.3)
            assert np.array_equal(got[c].reshape(B, 2), e), (step, c)
    # the feeder mixes the same tone over a whole run
    a = slb.DspIf(Cn, fs=fs, chain=slb.CHAIN_PASS); b = slb.DspIf(Cn, fs=fs, chain=slb.CHAIN_PASS)
    ticks = 16
    pc = rng.integers(-20000, 20000, (Cn, ticks * B, 2)).astype(np.int16)
    keys = np.array([1, 0, 1, 1, 0, 0], bool)
    for x in (a, b):
        x.DSP_Set_Sidetone(650, 0.2); x.DSP_Key(keys)
    exp = []
    for t in range(ticks):
        exp.append(a.DSP_Out_Buff_Read(2 * B)); a.DSP_Out_Buff_Write(pc[:, t * B:(t + 1) * B].reshape(Cn, -1))
    _, dac = b.feeder_run(None, pc)
    assert np.array_equal(dac, np.stack(exp, 1).reshape(Cn, -1, 2))
    assert np.array_equal(a.DSP_Out_Buff_Read(2 * B), b.DSP_Out_Buff_Read(2 * B))


@pytest.mark.parametrize("chain", [slb.CHAIN_RX_SSB_F32, slb.CHAIN_RX_SSB_Q15])
def test_rx_tx_switch_flushes_rings_and_chain_state(chain):
    """DSP_Set_TX / DSP_Set_RX as ptt_set_tx / ptt_set_rx use them (rxtx_if.c:255-317): a change of direction flushes the samples of
    both rings (pointers keep running, as DSP_Out_Buff_Mute, dsp_if.c:188-195) and the chain's carried state — after the switch the
    context behaves like a fresh one that has seen the same number of ticks; repeating the current direction changes nothing."""
    Cn, B = 4, 48
    x = slb.synth_iq(Cn, 64 * B)
    a = slb.DspIf(Cn, chain=chain); a.DSP_Init()
    fresh = slb.DspIf(Cn, chain=chain); fresh.DSP_Init()
    zeros = np.zeros((Cn, 2 * B), np.int16)
    for t in range(24):                                               # 24 ticks of signal on `a`, 24 ticks of silence on `fresh`: same pointers
        a.DSP_In_Buff_Write(x[:, t * B:(t + 1) * B].reshape(Cn, -1)); a.DSP_In_Buff_Read(4 * B)
        a.DSP_Out_Buff_Write(x[:, t * B:(t + 1) * B].reshape(Cn, -1)); a.DSP_Out_Buff_Read(2 * B)
        fresh.DSP_In_Buff_Write(zeros); fresh.DSP_In_Buff_Read(4 * B)
        fresh.DSP_Out_Buff_Write(zeros); fresh.DSP_Out_Buff_Read(2 * B)
    assert a.direction() == 0
    a.DSP_Set_RX()                                                    # no change of direction: nothing is flushed
    assert np.any(a.ring_iq(0)[0] != 0)
    a.DSP_Set_TX(); fresh.DSP_Set_TX()
    assert a.direction() == 1
    assert a.ring_ptrs(0) == fresh.ring_ptrs(0) and a.ring_ptrs(1) == fresh.ring_ptrs(1)
    for w in (0, 1):
        assert not np.any(a.ring_iq(w)[0]) and not np.any(a.ring_iq(w)[1])
    for t in range(24, 64):                                           # from here on the two contexts are indistinguishable
        for d in (a, fresh):
            d.DSP_In_Buff_Write(x[:, t * B:(t + 1) * B].reshape(Cn, -1)); d.DSP_Out_Buff_Write(x[:, t * B:(t + 1) * B].reshape(Cn, -1))
        assert np.array_equal(a.DSP_In_Buff_Read(4 * B), fresh.DSP_In_Buff_Read(4 * B)), t
        assert np.array_equal(a.DSP_Out_Buff_Read(2 * B), fresh.DSP_Out_Buff_Read(2 * B)), t


@pytest.mark.parametrize("chain", [slb.CHAIN_RX_SSB_Q15, slb.CHAIN_RX_SSB_F32, slb.CHAIN_TX_SSB_F32])
def test_chain_behind_the_ring_under_per_channel_cadence(chain):
    """dsp_if.c:252-300 per channel with a demodulator in front (SURVEY §8f.2 + §8f.3): channels whose producers and consumers tick at
    different rates each follow their own ring AND their own chain state / super-block fill. Every channel of the batched context
    must equal, bit for bit, a one-channel context that sees only that channel's events (that path is held to the oracle by
    test_chain_behind_the_1ms_cadence); a checkpoint taken while the channels are at different fill levels carries on the same way."""
    C, ticks, B = 7, 150, 48
    modes = [slb.MODE_USB, slb.MODE_LSB, slb.MODE_USB, slb.MODE_CW, slb.MODE_USB, slb.MODE_LSB, slb.MODE_USB]
    if chain == slb.CHAIN_RX_SSB_F32:
        modes[4] = slb.MODE_AM; modes[2] = slb.MODE_FM                # the FFT kernel and the complex-detector kernel serve their channels next to the SSB ones
    x = slb.synth_iq(C, ticks * B)
    d = slb.DspIf(C, chain=chain); single = [slb.DspIf(1, chain=chain) for _ in range(C)]
    for c in range(C):
        d.DSP_Set_Mode(modes[c], channel=c); single[c].DSP_Set_Mode(modes[c])
    rng = np.random.Generator(np.random.PCG64(5))
    def fires(c, t):
        w = not ((c == 1 and t % 13 == 12) or (c == 3 and t % 5 == 4) or (c == 5 and rng.random() < 0.1))
        r = not ((c == 2 and t % 17 == 16) or (c == 4 and t % 7 == 6) or (c == 5 and rng.random() < 0.1))
        return w, r
    pos = [0] * C                                                       # every producer walks its own stream
    for t in range(ticks):
        if t == 70:                                                     # checkpoint in mid-stream into a fresh context
            snap = d.state_save(); d = slb.DspIf(C, chain=chain); d.state_load(snap)
        act = np.array([fires(c, t) for c in range(C)])
        blk = np.zeros((C, 2 * B), np.int16)
        for c in range(C):
            if act[c, 0]:
                blk[c] = x[c, pos[c]:pos[c] + B].reshape(-1); pos[c] += B
        d.DSP_In_Buff_Write_Ch(blk, act[:, 0].astype(np.uint8))
        got = d.DSP_In_Buff_Read_Ch(4 * B, act[:, 1].astype(np.uint8))
        for c in range(C):
            if act[c, 0]: single[c].DSP_In_Buff_Write(blk[c:c + 1])
            exp = single[c].DSP_In_Buff_Read(4 * B)[0] if act[c, 1] else np.zeros(2 * B, np.int16)
            assert np.array_equal(got[c], exp), (t, c)
            assert d.ring_ptrs_channel(c, 0) == single[c].ring_ptrs(0), (t, c)


@pytest.mark.parametrize("chain,ticks", [(slb.CHAIN_PASS, 5), (slb.CHAIN_RX_SSB_Q15, 10), (slb.CHAIN_RX_SSB_F32, 16), (slb.CHAIN_TX_SSB_F32, 8)])
def test_live_feeder_pipeline_equals_the_per_tick_calls(chain, ticks, rng):
    """SURVEY.md §8f.1, the live form: chunks pushed into a three-stream pipeline (pinned staging; chunk n is copied in while n-1
    computes and n-2 is copied out) and popped in order give, bit for bit, what the per-tick calls give — with several chunks in
    flight at once, a side-tone keyed in between, and the per-tick API carrying on from the same state afterwards."""
    C, B, chunks, depth = 5, 48, 9, 3
    n = ticks * B
    adc = slb.synth_iq(C, (chunks * ticks + 8) * B)
    pc = rng.integers(-20000, 20000, (C, (chunks * ticks + 8) * B, 2)).astype(np.int16)
    a = slb.DspIf(C, chain=chain); b = slb.DspIf(C, chain=chain)
    for x in (a, b):
        x.DSP_Set_Sidetone(800, 0.1)
    exp_in, exp_dac = [], []
    for t in range(chunks * ticks):
        if t == 4 * ticks:
            a.DSP_Key(np.array([1, 0, 0, 1, 0], bool))
        blk = slice(t * B, (t + 1) * B)
        exp_dac.append(a.DSP_Out_Buff_Read(2 * B))
        a.DSP_In_Buff_Write(adc[:, blk].reshape(C, -1))
        a.DSP_Out_Buff_Write(pc[:, blk].reshape(C, -1))
        exp_in.append(a.DSP_In_Buff_Read(4 * B))
    exp_in = np.stack(exp_in, 1).reshape(C, -1, 2); exp_dac = np.stack(exp_dac, 1).reshape(C, -1, 2)
    lv = b.live_open(ticks, depth)
    got_in, got_dac, pushed = [], [], 0
    with pytest.raises(slb.SeleniteError):
        lv.pop()                                                                                # nothing in flight
    for k in range(chunks):
        if k == 4:
            while lv.in_flight():                                                               # the key changes between chunks: drain, as the per-tick run did at tick 4 * ticks
                u, v, _ = lv.pop(); got_in.append(u); got_dac.append(v)
            b.DSP_Key(np.array([1, 0, 0, 1, 0], bool))
        if lv.in_flight() == depth:
            u, v, lat = lv.pop(); got_in.append(u); got_dac.append(v); assert lat > 0
        lv.push(adc[:, k * n:(k + 1) * n], pc[:, k * n:(k + 1) * n]); pushed += 1
    assert lv.in_flight() > 1                                                                   # the pipeline really held several chunks
    while lv.in_flight():
        u, v, _ = lv.pop(); got_in.append(u); got_dac.append(v)
    lv.close()
    assert np.array_equal(np.concatenate(got_dac, 1), exp_dac)
    assert np.array_equal(np.concatenate(got_in, 1), exp_in)
    assert a.ring_ptrs(0) == b.ring_ptrs(0) and a.ring_ptrs(1) == b.ring_ptrs(1)
    for t in range(chunks * ticks, chunks * ticks + 8):                                          # and the per-tick API carries on from the same state
        blk = slice(t * B, (t + 1) * B)
        for d in (a, b):
            d.DSP_In_Buff_Write(adc[:, blk].reshape(C, -1))
        assert np.array_equal(a.DSP_In_Buff_Read(4 * B), b.DSP_In_Buff_Read(4 * B))
