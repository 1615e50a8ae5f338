"""GPU parity of the fused CHAN-64-f32 chain (BASELINE config 4: 192 kHz wideband streams -> 64-bin polyphase FFT
channelizer -> per-channel demod + AGC), called through the C ABI, against the oracle and the golden vectors.
Tolerances: demodulated audio within 1e-5 * max(|ref|, rms of the whole spectrum frame window) (test_golden.chan_tolerance),
gains to 2e-5 relative, int16 output within 1 LSB on < 2 % of samples (truncating pack of a float chain)."""
import os

import numpy as np
import pytest

import selenite_lite_b200 as slb
from test_golden import GOLD, chan_tolerance

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def run_gpu(d, x, want_dbg=True):
    S, T = x.shape[0], x.shape[1]
    xd = torch.from_numpy(x).cuda()
    audio = torch.zeros((S, 64, T // 64), dtype=torch.float32, device="cuda") if want_dbg else None
    gain = torch.zeros((S, 64, T // 192), dtype=torch.float32, device="cuda") if want_dbg else None
    d.set_debug_taps(audio, gain)
    y = d.chan_process(xd)
    torch.cuda.synchronize()
    d.set_debug_taps(None, None)
    return y.cpu().numpy(), (audio.cpu().numpy() if want_dbg else None), (gain.cpu().numpy() if want_dbg else None)


def check_int16(y, exp):
    d = np.abs(y.astype(np.int32) - exp.astype(np.int32))
    assert d.max() <= 1, "int16 output differs by %d LSB" % d.max()
    assert np.mean(d > 0) < 0.02, "%.2f %% of samples differ" % (100 * np.mean(d > 0))
    assert np.array_equal(y[..., 0], y[..., 1])


@pytest.mark.parametrize("name,mode", [("prod", slb.MODE_USB), ("env", slb.MODE_AM)])
def test_golden_vectors(name, mode):
    g = np.load(os.path.join(GOLD, "chan64_f32.npz"))
    d = slb.DspIf(1, fs=192000, chain=slb.CHAIN_CHAN64_F32)
    d.DSP_Set_Mode(mode)
    assert np.array_equal(np.array(d.chan_params().proto[:512], np.float32), g["chan_proto"])
    y, audio, gain = run_gpu(d, g["chan_in"][None])
    ref_audio = g["chan_%s_audio" % name]
    err = np.abs(audio[0] - ref_audio); tol = chan_tolerance(ref_audio)
    assert np.all(err <= tol + 1e-9), "worst audio error %.2f x tolerance" % np.max(err / (tol + 1e-9))
    assert np.allclose(gain[0], g["chan_%s_gain" % name], rtol=2e-5)
    check_int16(y[0], g["chan_%s_out" % name])


@pytest.mark.parametrize("streams,frames", [(1, 768), (2, 6144), (3, 6144 + 768), (5, 6144 * 3), (9, 6144 * 2 + 768 * 5)])
def test_vs_oracle_ragged_shapes(best_oracle, streams, frames):
    """Single partial tile, exactly one tile, ragged last tiles, several tiles chained by the look-back."""
    x = slb.synth_wideband(streams, frames)
    d = slb.DspIf(streams, fs=192000, chain=slb.CHAIN_CHAN64_F32)
    y, audio, gain = run_gpu(d, x)
    for s in range(streams):
        exp, a, g_, _ = best_oracle.chan_f32(d.oracle_params(), x[s])
        err = np.abs(audio[s] - a); tol = chan_tolerance(a)
        assert np.all(err <= tol + 1e-9), (s, float(np.max(err / (tol + 1e-9))))
        assert np.allclose(gain[s], g_, rtol=2e-5)
        check_int16(y[s], exp)


def test_worst_case_inputs(best_oracle):
    """Wideband streams at the rails: both rails held at -32768, at +32767 (DC lands in bin 0), and a full-scale complex tone at the
    centre of bin 5 — product and envelope detector, against the oracle."""
    T = 6144 * 2
    n = np.arange(T)
    tone = np.exp(2j * np.pi * (5 * 3000.0 + 400.0) * n / 192000.0)
    x = np.stack([np.full((T, 2), -32768, np.int16), np.full((T, 2), 32767, np.int16),
                  np.stack([np.clip(np.rint(tone.real * 32767), -32768, 32767), np.clip(np.rint(tone.imag * 32767), -32768, 32767)], axis=1).astype(np.int16)])
    for mode in (slb.MODE_USB, slb.MODE_AM):
        d = slb.DspIf(3, fs=192000, chain=slb.CHAIN_CHAN64_F32); d.DSP_Set_Mode(mode)
        y, audio, gain = run_gpu(d, x)
        for s_ in range(3):
            exp, a, g_, _ = best_oracle.chan_f32(d.oracle_params(mode), x[s_])
            err = np.abs(audio[s_] - a); tol = chan_tolerance(a)
            assert np.all(err <= tol + 1e-9), (mode, s_, float(np.max(err / (tol + 1e-9))))
            dd = np.abs(y[s_].astype(np.int32) - exp.astype(np.int32))
            assert dd.max() <= 1 and np.mean(dd > 0) < 0.10, (mode, s_, int(dd.max()), float(np.mean(dd > 0)))


def test_release_walk_across_many_tiles_is_exact(best_oracle):
    """The AGC release is a long sequential recurrence: a burst followed by near silence makes every later tile's
    envelope depend on a carry-in that is many tiles old. The look-back must reproduce the oracle's gains exactly."""
    S, T = 2, 6144 * 12
    x = slb.synth_wideband(S, T, sigma=0.0005)
    x[:, 6144:] //= 64                                        # loud first tile, then 36 dB down
    d = slb.DspIf(S, fs=192000, chain=slb.CHAIN_CHAN64_F32)
    y, audio, gain = run_gpu(d, x)
    for s in range(S):
        exp, a, g_, _ = best_oracle.chan_f32(d.oracle_params(), x[s])
        quiet = np.abs(audio[s] - a) <= chan_tolerance(a) + 1e-9
        assert np.all(quiet)
        assert np.allclose(gain[s], g_, rtol=2e-5)
        check_int16(y[s], exp)


def test_state_carries_across_calls_and_checkpoint(best_oracle):
    S, T = 3, 6144 * 2
    x = slb.synth_wideband(S, 2 * T)
    whole = run_gpu(slb.DspIf(S, fs=192000, chain=slb.CHAIN_CHAN64_F32), x, want_dbg=False)[0]
    a = slb.DspIf(S, fs=192000, chain=slb.CHAIN_CHAN64_F32)
    y0 = run_gpu(a, np.ascontiguousarray(x[:, :T]), want_dbg=False)[0]
    snap = a.state_save()
    y1 = run_gpu(a, np.ascontiguousarray(x[:, T:]), want_dbg=False)[0]
    assert np.array_equal(np.concatenate([y0, y1], 2), whole)      # tile-aligned cut: the same arithmetic
    b = slb.DspIf(S, fs=192000, chain=slb.CHAIN_CHAN64_F32)
    b.state_load(snap)
    assert np.array_equal(run_gpu(b, np.ascontiguousarray(x[:, T:]), want_dbg=False)[0], y1)
    # 4 ms calls (768 frames) cut inside tiles: same result (FIR sums and the envelope walk do not depend on the cut)
    c = slb.DspIf(S, fs=192000, chain=slb.CHAIN_CHAN64_F32)
    parts = [run_gpu(c, np.ascontiguousarray(x[:, o:o + 768]), want_dbg=False)[0] for o in range(0, T, 768)]
    assert np.array_equal(np.concatenate(parts, 2), whole[:, :, :T // 64])


def test_config4_width_shards_and_host_path(best_oracle):
    """BASELINE config 4: 64 wideband streams x 64 bins = 4096 narrowband channels. Shards by stream reproduce the whole
    byte for byte; the host-buffer path equals the device path; a sample of streams is checked against the oracle."""
    S, T = 64, 6144 * 4
    base = slb.synth_wideband(8, T)
    rng = np.random.Generator(np.random.PCG64(11))
    x = np.ascontiguousarray(base[rng.integers(0, 8, S)] + rng.integers(-40, 40, (S, T, 2)).astype(np.int16))
    whole = run_gpu(slb.DspIf(S, fs=192000, chain=slb.CHAIN_CHAN64_F32), x, want_dbg=False)[0]
    assert whole.shape == (S, 64, T // 64, 2)
    for world, rank in ((2, 1), (8, 0), (8, 7)):
        lo, hi = slb.shard.shard_range(S, rank, world)
        part = run_gpu(slb.DspIf(hi - lo, fs=192000, chain=slb.CHAIN_CHAN64_F32), np.ascontiguousarray(x[lo:hi]), want_dbg=False)[0]
        assert np.array_equal(part, whole[lo:hi]), (world, rank)
    assert np.array_equal(slb.DspIf(S, fs=192000, chain=slb.CHAIN_CHAN64_F32).chan_process(x), whole)
    d = slb.DspIf(1, fs=192000, chain=slb.CHAIN_CHAN64_F32)
    for s in rng.integers(0, S, 3):
        exp, _, _, _ = best_oracle.chan_f32(d.oracle_params(), x[s])
        check_int16(whole[s], exp)


def test_bad_calls_are_rejected():
    d = slb.DspIf(2, fs=192000, chain=slb.CHAIN_CHAN64_F32)
    with pytest.raises(slb.SeleniteError):
        d.chan_process(torch.zeros((2, 1000, 2), dtype=torch.int16, device="cuda"))     # not a multiple of 768
    with pytest.raises(slb.SeleniteError):
        d.rx_process(torch.zeros((2, 768, 2), dtype=torch.int16, device="cuda"))
    with pytest.raises(slb.SeleniteError):
        d.DSP_In_Buff_Write(np.zeros((2, 384), np.int16))                               # no firmware ring for this chain
    with pytest.raises(slb.SeleniteError):
        d.DSP_Set_Mode(slb.MODE_FM)
    assert d.kernel_launches() == 0
