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
