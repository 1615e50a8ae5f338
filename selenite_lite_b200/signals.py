"""Synthetic inputs frozen by SURVEY.md §8(d): complex tone + complex white Gaussian noise, int16 I/Q."""
import numpy as np

SEED = 0x5E1E217E


def channel_tone_hz(c):
    """Config 2/5 tone placement: f0 = 300 + 2400 * ((c * 2654435761 mod 2^32) / 2^32) Hz."""
    return 300.0 + 2400.0 * (((int(c) * 2654435761) % (1 << 32)) / float(1 << 32))


def synth_iq(channels, frames, fs=48000, f0=None, amp=0.25, sigma=0.01, seed=SEED, sideband=+1, first_channel=0):
    """int16[channels][frames][2]. Channel c uses PCG64(seed + c); config 1 is channels=1, f0=1000."""
    out = np.empty((channels, frames, 2), np.int16)
    n = np.arange(frames, dtype=np.float64)
    for k in range(channels):
        c = first_channel + k
        rng = np.random.Generator(np.random.PCG64(seed + c))
        f = channel_tone_hz(c) if f0 is None else float(f0)
        ph = 2.0 * np.pi * sideband * f * n / fs
        i = amp * np.cos(ph) + sigma * rng.standard_normal(frames)
        q = amp * np.sin(ph) + sigma * rng.standard_normal(frames)
        out[k, :, 0] = np.clip(np.rint(i * 32768.0), -32768, 32767).astype(np.int16)
        out[k, :, 1] = np.clip(np.rint(q * 32768.0), -32768, 32767).astype(np.int16)
    return out


def synth_mic(channels, frames, fs=48000, tones=(700.0, 1900.0), amp=0.2, sigma=0.005, seed=SEED, first_channel=0):
    """Config 3 input: int16[channels][frames][2] with L = R (what the codec delivers in TX,
    Core/Src/codec_if.c:304-306): two-tone 700 Hz + 1900 Hz at 0.2 FS each + noise; channel c detunes the pair by c Hz."""
    out = np.empty((channels, frames, 2), np.int16)
    n = np.arange(frames, dtype=np.float64)
    for k in range(channels):
        c = first_channel + k
        rng = np.random.Generator(np.random.PCG64(seed + 0x7A + c))
        m = sigma * rng.standard_normal(frames)
        for f in tones:
            m = m + amp * np.cos(2.0 * np.pi * (f + (c % 97)) * n / fs + 0.1 * c)
        v = np.clip(np.rint(m * 32768.0), -32768, 32767).astype(np.int16)
        out[k, :, 0] = v; out[k, :, 1] = v
    return out


def wideband_occupancy(stream, bins=64, seed=SEED):
    """Config 4: which bins of wideband stream `stream` carry a tone (random half of the bins)."""
    rng = np.random.Generator(np.random.PCG64(seed + 0xC4A7 + int(stream)))
    return rng.random(bins) < 0.5


def synth_wideband(streams, frames, fs=192000, bins=64, amp=0.02, sigma=0.01, offset_hz=500.0, seed=SEED, first_stream=0):
    """Config 4 input: int16[streams][frames][2]; one complex tone at bin-centre + 500 Hz in every occupied
    bin (bin k is centred on k * fs/bins, k >= bins/2 meaning negative frequencies), plus complex white noise."""
    out = np.empty((streams, frames, 2), np.int16)
    n = np.arange(frames, dtype=np.float64)
    for k in range(streams):
        s = first_stream + k
        rng = np.random.Generator(np.random.PCG64(seed + 0x1D00 + s))
        z = sigma * (rng.standard_normal(frames) + 1j * rng.standard_normal(frames))
        occ = wideband_occupancy(s, bins, seed)
        for b in np.nonzero(occ)[0]:
            fb = (b if b < bins // 2 else b - bins) * fs / bins + offset_hz
            z = z + amp * np.exp(1j * (2.0 * np.pi * fb * n / fs + 0.37 * b + 0.11 * s))
        out[k, :, 0] = np.clip(np.rint(z.real * 32768.0), -32768, 32767).astype(np.int16)
        out[k, :, 1] = np.clip(np.rint(z.imag * 32768.0), -32768, 32767).astype(np.int16)
    return out


def synth_fm(channels, frames, fs=48000, audio_hz=1000.0, deviation_hz=2500.0, carrier_hz=0.0, amp=0.25, sigma=0.002, seed=SEED, first_channel=0):
    """FM test input (mode byte 0x08): int16[channels][frames][2], a carrier at `carrier_hz` frequency-modulated by a tone of
    `audio_hz` (channel c detunes it by c Hz) with peak deviation `deviation_hz`, plus complex white noise."""
    out = np.empty((channels, frames, 2), np.int16)
    n = np.arange(frames, dtype=np.float64)
    for k in range(channels):
        c = first_channel + k
        rng = np.random.Generator(np.random.PCG64(seed + 0xF3 + c))
        fa = audio_hz + (c % 89)
        ph = 2.0 * np.pi * carrier_hz * n / fs + (deviation_hz / fa) * np.sin(2.0 * np.pi * fa * n / fs + 0.3 * c)
        i = amp * np.cos(ph) + sigma * rng.standard_normal(frames)
        q = amp * np.sin(ph) + sigma * rng.standard_normal(frames)
        out[k, :, 0] = np.clip(np.rint(i * 32768.0), -32768, 32767).astype(np.int16)
        out[k, :, 1] = np.clip(np.rint(q * 32768.0), -32768, 32767).astype(np.int16)
    return out
