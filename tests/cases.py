"""Shared small parity cases (configuration + seeded synthetic input) used by the oracle, golden and GPU tests.
Sizes are chosen so the oracle finishes each in well under a second."""
from __future__ import annotations

import numpy as np

from airband_b200 import config as cm
from airband_b200 import workloads as wl
from airband_b200.config import Config, Device, make_channel


def case_am_u8(n_batches=3):
    """2 AM channels, U8, fft 256, manual squelch; one channel also writes raw I/Q (rawfile) without a low-pass."""
    sr, n, w, cf = 512000, 256, 8000, 120000000
    chans = [make_channel(cf - 100000, cf, sr, n, w, squelch_dbfs=-30.0),
             make_channel(cf + 75000, cf, sr, n, w, squelch_dbfs=-30.0, rawfile=True, ampfactor=1.5)]
    cfg = Config(fft_size=n, wave_rate=w, devices=[Device(sample_rate=sr, sfmt=cm.SFMT_U8, centerfreq=cf, channels=chans)])
    raws = [wl.synth_iq(cfg, 0, wl.samples_for_batches(cfg, 0, n_batches), key_on_s=0.11, key_off_s=0.07, amplitude=0.2)]
    return cfg, raws


def case_nfm_s16(n_batches=6, fm_demod=cm.FM_FAST_ATAN2):
    """3 NFM channels (NFM build: WAVE_RATE 16000), S16, fft 256, bandwidth 5000 + CTCSS + notch, ampfactor 2.
    Channel 0 receives its tone, channel 1 its tone without de-emphasis, channel 2 a wrong tone (never opens)."""
    sr, n, w, cf = 400000, 256, 16000, 162000000
    chans = [make_channel(cf - 125000, cf, sr, n, w, modulation=cm.MOD_NFM, bandwidth=5000, ampfactor=2.0, squelch_dbfs=-30.0,
                          notch_hz=100.0, ctcss_hz=100.0),
             make_channel(cf + 100000, cf, sr, n, w, modulation=cm.MOD_NFM, bandwidth=5000, ampfactor=2.0, squelch_dbfs=-30.0,
                          notch_hz=123.0, ctcss_hz=123.0, tau_us=0),
             make_channel(cf + 50000, cf, sr, n, w, modulation=cm.MOD_NFM, bandwidth=5000, ampfactor=2.0, squelch_dbfs=-30.0,
                          ctcss_hz=85.4)]
    chans[2].synth_ctcss_hz = 67.0
    cfg = Config(fft_size=n, wave_rate=w, fm_demod=fm_demod,
                 devices=[Device(sample_rate=sr, sfmt=cm.SFMT_S16, centerfreq=cf, channels=chans)])
    raws = [wl.synth_iq(cfg, 0, wl.samples_for_batches(cfg, 0, n_batches), key_on_s=0.55, key_off_s=0.04, amplitude=0.2)]
    return cfg, raws


def case_am_bw_f32(n_batches=4):
    """AM with `bandwidth` (raw I/Q derotation + low-pass + post-filter squelch), rawfile output, F32 input,
    automatic (SNR) squelch; plus a plain AM channel with the default squelch."""
    sr, n, w, cf = 128000, 256, 8000, 118000000
    chans = [make_channel(cf - 25000, cf, sr, n, w, bandwidth=6000, rawfile=True, squelch_snr_db=12.0),
             make_channel(cf + 25000, cf, sr, n, w)]
    cfg = Config(fft_size=n, wave_rate=w, devices=[Device(sample_rate=sr, sfmt=cm.SFMT_F32, centerfreq=cf, channels=chans)])
    raws = [wl.synth_iq(cfg, 0, wl.samples_for_batches(cfg, 0, n_batches), key_on_s=0.14, key_off_s=0.21, amplitude=0.2,
                        noise_sigma=0.004)]
    return cfg, raws


def case_s8_two_devices(n_batches=2):
    """Two devices with different formats/rates in one process (S8 @ 1.024 Msps and U8 @ 2.56 Msps), fft 512."""
    n, w = 512, 8000
    d0 = Device(sample_rate=1024000, sfmt=cm.SFMT_S8, centerfreq=130000000,
                channels=[make_channel(130000000 + 150000, 130000000, 1024000, n, w, squelch_dbfs=-35.0)])
    d1 = Device(sample_rate=2560000, sfmt=cm.SFMT_U8, centerfreq=120000000,
                channels=[make_channel(120000000 + o, 120000000, 2560000, n, w, squelch_dbfs=-30.0) for o in (-500000, 225000, 600000)])
    cfg = Config(fft_size=n, wave_rate=w, devices=[d0, d1])
    raws = [wl.synth_iq(cfg, i, wl.samples_for_batches(cfg, i, n_batches), key_on_s=0.1, key_off_s=0.05) for i in range(2)]
    return cfg, raws


CASES = {
    "am_u8": case_am_u8,
    "nfm_s16": case_nfm_s16,
    "nfm_s16_quadri": lambda: case_nfm_s16(fm_demod=cm.FM_QUADRI_DEMOD),
    "am_bw_f32": case_am_bw_f32,
    "s8_two_devices": case_s8_two_devices,
}
