"""Named workload shapes (BASELINE.json `configs`) and the synthetic I/Q generator.

Every shape is the reference's own configuration surface (devices[] of channels[], reference src/config.cpp)
already resolved to bins / phase steps / filter settings.  Synthetic input follows SURVEY.md §8d: per device a
sum of keyed carriers at the configured channel frequencies (AM 60 % / 1 kHz, NFM +-2.5 kHz / 1 kHz plus an
optional CTCSS sub-tone at 10 % of that deviation), white noise sigma = 0.002 full scale, rng seed 1234 + device.
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np

from . import config as cfgm
from .config import Channel, Config, Device, MOD_AM, MOD_NFM, SFMT_F32, SFMT_S16, SFMT_S8, SFMT_U8, make_channel


def _raster(n_channels: int, spacing: int) -> List[int]:
    """Channel offsets from centre: spacing * (i - C/2), skipping 0 (the DC bin)."""
    offs = []
    for i in range(n_channels):
        k = i - n_channels // 2
        if k >= 0:
            k += 1
        offs.append(k * spacing)
    return offs


def cfg1(two_channels: bool = False, squelch_dbfs: float = -30.0) -> Config:
    """config/basic_multichannel.conf shape: 1 device, 2.56 Msps U8, fft_size 512, AM 119.5 MHz @ cf 120.0 MHz (bin 411)."""
    sr, n, w, cf = 2560000, 512, 8000, 120000000
    freqs = [119500000] + ([120225000] if two_channels else [])
    chans = [make_channel(f, cf, sr, n, w, squelch_dbfs=squelch_dbfs) for f in freqs]
    return Config(fft_size=n, wave_rate=w, devices=[Device(sample_rate=sr, sfmt=SFMT_U8, centerfreq=cf, channels=chans)])


def cfg2(n_devices: int = 64, n_channels: int = 8, squelch_dbfs: float = -30.0, fft_size: int = 2048) -> Config:
    """64 synthetic devices x 2.56 Msps U8, fft_size 2048, 8 AM channels each (the headline throughput shape)."""
    sr, w, cf = 2560000, 8000, 120000000
    devs = []
    for _ in range(n_devices):
        chans = [make_channel(cf + o, cf, sr, fft_size, w, squelch_dbfs=squelch_dbfs) for o in _raster(n_channels, 25000)]
        devs.append(Device(sample_rate=sr, sfmt=SFMT_U8, centerfreq=cf, channels=chans))
    return Config(fft_size=fft_size, wave_rate=w, devices=devs)


def cfg3(n_devices: int = 8, n_channels: int = 32, sfmt: int = SFMT_S16, parity: bool = False, ctcss_hz: float = 100.0,
         fft_size: int = 4096, sample_rate: int = 10000000) -> Config:
    """config/noaa.conf shape scaled up: 10 Msps, NFM build (WAVE_RATE 16000), fft_size 4096, 32 NFM channels with
    bandwidth 5000 (low-pass 2500 Hz), ampfactor 2, CTCSS + notch on the same tone.  Throughput runs keep
    squelch_snr_threshold = 0 (noaa.conf:24); the parity variant uses a manual -30 dBFS level."""
    w, cf = 16000, 162000000
    devs = []
    for _ in range(n_devices):
        chans = []
        for o in _raster(n_channels, 25000):
            chans.append(make_channel(cf + o, cf, sample_rate, fft_size, w, modulation=MOD_NFM, bandwidth=5000, ampfactor=2.0,
                                      squelch_dbfs=-30.0 if parity else 0.0, squelch_snr_db=-1.0 if parity else 0.0,
                                      notch_hz=ctcss_hz, ctcss_hz=ctcss_hz))
        devs.append(Device(sample_rate=sample_rate, sfmt=sfmt, centerfreq=cf, channels=chans))
    return Config(fft_size=fft_size, wave_rate=w, devices=devs)


def cfg4(n_devices: int = 4, n_channels: int = 4, squelch_dbfs: float = -30.0) -> Config:
    """config/big_mixer.conf shape as BASELINE.json reads it: 4 devices x 4 AM channels, fft_size 512; mixer m sums
    channel m of every device (see mixers_cfg4)."""
    sr, n, w, cf = 2560000, 512, 8000, 156737500
    devs = []
    for _ in range(n_devices):
        chans = [make_channel(cf + o, cf, sr, n, w, squelch_dbfs=squelch_dbfs) for o in _raster(n_channels, 25000)]
        devs.append(Device(sample_rate=sr, sfmt=SFMT_U8, centerfreq=cf, channels=chans))
    return Config(fft_size=n, wave_rate=w, devices=devs)


def mixers_cfg4(cfg: Config):
    """[(mixer_index, [(device, channel, ampfactor, balance), ...])] — mixer m takes channel m of every device."""
    n_mix = len(cfg.devices[0].channels)
    return [(m, [(d, m, 1.0, 0.0) for d in range(len(cfg.devices))]) for m in range(n_mix)]


def cfg5(n_devices: int = 512, n_channels: int = 8, squelch_dbfs: float = -30.0) -> Config:
    """4096 synthetic devices x 2.56 Msps U8, fft_size 512, 8 AM channels each, sharded 512 per GPU."""
    return cfg2(n_devices=n_devices, n_channels=n_channels, squelch_dbfs=squelch_dbfs, fft_size=512)


# ------------------------------------------------------------------------------------------------------------------
# synthetic I/Q
# ------------------------------------------------------------------------------------------------------------------
def synth_iq(cfg: Config, dev_index: int, n_samples: int, *, seed: Optional[int] = None, key_on_s: float = 2.0,
             key_off_s: float = 1.0, amplitude: Optional[float] = None, noise_sigma: float = 0.002, am_depth: float = 0.6,
             tone_hz: float = 1000.0, fm_dev_hz: float = 2500.0, chunk: int = 1 << 20) -> np.ndarray:
    """Raw ring-format samples (interleaved I,Q of the device's sample_format_t) for one device."""
    dev = cfg.devices[dev_index]
    sr = float(dev.sample_rate)
    rng = np.random.default_rng(1234 + dev_index if seed is None else seed)
    C = max(1, len(dev.channels))
    if amplitude is None:
        amplitude = min(0.15, 0.8 / (C * (1.0 + am_depth)))
    out = np.empty(2 * n_samples, dtype=cfgm.NP_DTYPE[dev.sfmt])
    period = key_on_s + key_off_s
    # FM phase accumulators continue across chunks
    fm_phase = [0.0] * C
    for start in range(0, n_samples, chunk):
        n = min(chunk, n_samples - start)
        t = (start + np.arange(n, dtype=np.float64)) / sr
        x = np.zeros(n, dtype=np.complex128)
        for ci, ch in enumerate(dev.channels):
            # staggered keying so that channels open and close at different times
            # (each channel starts in its key-off phase so automatic squelches see the noise floor first)
            tk = np.mod(t + ci * 0.37 * period / C, period) if key_off_s > 0 else None
            gate = (tk >= key_off_s).astype(np.float64) if tk is not None else 1.0
            car = 2.0 * np.pi * ch.offset_hz * t
            if ch.modulation == MOD_NFM:
                mod = fm_dev_hz * np.cos(2.0 * np.pi * tone_hz * t)
                sub = ch.ctcss_hz if ch.synth_ctcss_hz < 0 else ch.synth_ctcss_hz
                if sub > 0:
                    mod = mod + 0.1 * fm_dev_hz * np.cos(2.0 * np.pi * sub * t)
                ph = fm_phase[ci] + 2.0 * np.pi * np.cumsum(mod) / sr
                fm_phase[ci] = float(ph[-1])
                x += amplitude * gate * np.exp(1j * (car + ph))
            else:
                env = 1.0 + am_depth * np.cos(2.0 * np.pi * tone_hz * t)
                x += amplitude * gate * env * np.exp(1j * car)
        x += noise_sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
        iq = np.empty(2 * n, dtype=np.float64)
        iq[0::2] = x.real
        iq[1::2] = x.imag
        sl = slice(2 * start, 2 * (start + n))
        if dev.sfmt == SFMT_U8:
            out[sl] = np.clip(np.rint(127.5 * iq + 127.5), 0, 255).astype(np.uint8)
        elif dev.sfmt == SFMT_S8:
            out[sl] = np.clip(np.rint(127.5 * iq - 0.5), -127, 127).astype(np.int8)
        elif dev.sfmt == SFMT_S16:
            out[sl] = np.clip(np.rint(32767.0 * iq), -32767, 32767).astype(np.int16)
        else:
            out[sl] = iq.astype(np.float32)
    return out


def samples_for_batches(cfg: Config, dev_index: int, n_batches: int) -> int:
    """Complex samples a device must have buffered for n_batches to complete under the reference's ring rule
    `available >= bps + fft_size*bytes_per_sample*2` (reference src/rtl_airband.cpp:394-400): the first batch fires
    after WAVE_BATCH + AGC_EXTRA frames, later ones every WAVE_BATCH frames (rtl_airband.cpp:494)."""
    hop = cfg.hop(dev_index)
    frames = n_batches * cfg.wave_batch + cfgm.AGC_EXTRA
    return (frames - 1) * hop + hop + cfg.fft_size
