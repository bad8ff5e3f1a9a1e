"""The reference's own behavioural unit tests — src/test_squelch.cpp:51-281 (8 tests) and src/test_ctcss.cpp:122-155 (5 tests) —
driven through the GPU demodulation state machine (K2) instead of the CPU classes.  The per-sample inputs those tests feed
(`process_raw_sample(level)`, `process_audio_sample(tone)`) enter through the stage tap abg_debug_inject_wavein as the
|X[bin]| series of an AM channel: a constant level is the raw sample, and a level modulated by 1.5 * a(t) demodulates
(rtl_airband.cpp:553-563) to the audio a(t) the CTCSS detectors see.  Every scenario is checked twice: against the CPU
oracle given the same injection (strict BASELINE.md gate, identical counters), and against the upstream assertions, read
off the GPU results at batch granularity (axcindicate, Squelch getters).  Noise is a seeded numpy normal (upstream seeds
from std::random_device, generate_signal.cpp:41-46, which is not reproducible)."""
import math

import numpy as np
import pytest

import oracle_py as op
import parity
from airband_b200 import config as cm
from airband_b200 import lib

pytestmark = pytest.mark.gpu

B = 1000
RAW_NO_SIGNAL, RAW_SIGNAL = 0.05, 0.75          # test_squelch.cpp:33-34
TONES = [67.0, 69.3, 71.9, 74.4, 77.0, 79.7, 82.5, 85.4, 88.5, 91.5, 94.8, 97.4, 100.0, 103.5, 107.2, 110.9, 114.8, 118.8, 123.0, 127.3,
         131.8, 136.5, 141.3, 146.2, 150.0, 151.4, 156.7, 159.8, 162.2, 165.5, 167.9, 171.3, 173.8, 177.3, 179.9, 183.5, 186.2, 189.9,
         192.8, 196.6, 199.5, 203.5, 206.5, 210.7, 218.1, 225.7, 229.1, 233.6, 241.8, 250.3, 254.1]


def make_cfg(ctcss_tones):
    """One device whose channels differ only in the CTCSS tone they wait for (0 = no CTCSS), default automatic squelch."""
    sr, n, w = 2560000, 512, 8000
    chans = [cm.make_channel(100000, 0, sr, n, w, ctcss_hz=float(t)) for t in ctcss_tones]
    return cm.Config(fft_size=n, wave_rate=w, devices=[cm.Device(sample_rate=sr, sfmt=cm.SFMT_U8, centerfreq=0, channels=chans)])


def tone_audio(freq, n, ampl=0.2, noise=0.0, seed=0, start=1):
    k = np.arange(start, start + n, dtype=np.float64)               # generate_signal.cpp:32-35: first sample is n = 1
    x = np.float32(ampl) * np.sin(2 * math.pi * k * float(np.float32(freq)) / 8000.0) if freq else np.zeros(n)
    if noise:
        x = x + noise * np.random.default_rng(seed).normal(0.0, 0.1, n)
    return x


def carrier(audio):
    """|X| series whose AM demodulation is `audio`: level * (1 + 1.5 * a), cf. waveout = (wavein - agc) / (1.5 * agc)."""
    return (RAW_SIGNAL * (1.0 + 1.5 * np.asarray(audio))).astype(np.float32)


def pad(x):
    x = np.asarray(x, np.float32)
    r = (-len(x)) % B
    return np.concatenate([x, np.full(r, x[-1], np.float32)]) if r else x


class Both:
    """The GPU engine and the oracle fed the same injected series, batch by batch, with the parity checks built in."""

    def __init__(self, ctcss_tones):
        self.cfg = make_cfg(ctcss_tones)
        self.C = len(ctcss_tones)
        self.e = lib.Engine(self.cfg, max_batches_per_run=4)
        self.o = op.Oracle(self.cfg)
        self.axc = []           # per batch: axcindicate[C] of the GPU
        self.audio = []

    def feed(self, series):
        """series: 1-D (same for every channel) whole number of batches."""
        series = pad(series)
        for b0 in range(0, len(series), 4 * B):
            chunk = series[b0:b0 + 4 * B]
            w = np.tile(chunk, (self.C, 1))
            nb = len(chunk) // B
            assert self.e.inject_wavein(0, w) == nb and self.o.inject_wavein(0, w) == nb
            for _ in range(nb):
                g, o = self.e.fetch(0, want_iq=False), self.o.fetch(0)
                assert g is not None and o is not None
                r = parity.strict((g[0], None, g[2][None, :]), (o[0], None, o[2][None, :]))
                assert r["ok"], r
                self.axc.append(g[2].copy())
                self.audio.append(g[0].copy())
        for c in range(self.C):
            gs, os_ = self.e.stats(0, c), self.o.stats(0, c)
            for f in ("open_count", "flappy_count", "ctcss_count", "no_ctcss_count", "active_counter"):
                assert getattr(gs, f) == getattr(os_, f), (c, f, getattr(gs, f), getattr(os_, f))
            for f in ("noise_level", "signal_level", "squelch_level", "agcavgfast"):
                a, b = getattr(gs, f), getattr(os_, f)
                assert abs(a - b) <= 1e-4 * max(1.0, abs(a), abs(b)), (c, f, a, b)

    def settle_noise_floor(self):
        """send_samples_for_noise_floor(), test_squelch.cpp:39-45"""
        for _ in range(40):
            if self.e.stats(0, 0).noise_level <= 1.01 * RAW_NO_SIGNAL:
                break
            self.feed(np.full(B, RAW_NO_SIGNAL, np.float32))
        s = self.e.stats(0, 0)
        assert s.noise_level <= 1.01 * RAW_NO_SIGNAL and RAW_SIGNAL > s.squelch_level

    def open_batches(self, c=0):
        return [bool(a[c] != ord(' ')) for a in self.axc]

    def close(self):
        self.e.close()
        self.o.close()


# ---------------------------------------------------------------------------------------------- test_squelch.cpp:51-281
def test_squelch_default_object():
    t = Both([0])
    assert t.e.stats(0, 0).open_count == 0
    t.close()


def test_squelch_noise_floor():
    t = Both([0])
    assert t.e.stats(0, 0).noise_level > 10.0 * RAW_NO_SIGNAL
    last = t.e.stats(0, 0).noise_level
    for _ in range(40):
        t.feed(np.full(B, RAW_NO_SIGNAL, np.float32))
        this = t.e.stats(0, 0).noise_level
        assert this <= last
        if this == last:
            break
        last = this
    assert t.e.stats(0, 0).noise_level < 1.01 * RAW_NO_SIGNAL
    assert not any(t.open_batches())
    t.close()


def test_squelch_normal_operation():
    t = Both([0])
    t.settle_noise_floor()
    n0 = len(t.axc)
    t.feed(np.full(2 * B, RAW_SIGNAL, np.float32))       # opens within 500 samples and stays open
    assert t.open_batches()[n0:] == [True, True] and t.e.stats(0, 0).open_count == 1
    first_open = np.flatnonzero(t.audio[n0][0] != 0.0)
    assert first_open.size and first_open[0] < 500
    t.feed(np.full(2 * B, RAW_NO_SIGNAL, np.float32))    # the squelch closes within 100 raw samples; audio lags wavein by
    assert t.open_batches()[n0 + 2:] == [True, False]    # AGC_EXTRA = 100 samples (rtl_airband.cpp:558), so it ends within 200
    assert not np.any(t.audio[n0 + 2][0][200:] != 0.0)
    t.close()


def test_squelch_dead_spot():
    t = Both([0])
    t.settle_noise_floor()
    n0 = len(t.axc)
    x = np.concatenate([np.full(1500, RAW_SIGNAL), np.full(50, RAW_NO_SIGNAL), np.full(1450, RAW_SIGNAL)]).astype(np.float32)
    t.feed(x)
    assert t.open_batches()[n0:] == [True, True, True]
    assert t.e.stats(0, 0).open_count == 1               # the 50-sample dead spot did not close it
    t.close()


def test_squelch_should_process_audio():
    t = Both([0])
    t.settle_noise_floor()
    n0 = len(t.axc)
    t.feed(np.concatenate([np.full(B, RAW_SIGNAL), np.full(B, RAW_NO_SIGNAL)]).astype(np.float32))
    a = np.concatenate([t.audio[n0][0], t.audio[n0 + 1][0]])
    nz = np.flatnonzero(a != 0.0)
    # audio appears once (no flapping): processed from the opening until shortly after the signal ends
    assert nz.size > 0 and nz[0] < 500 and B <= nz[-1] < B + 200
    t.close()


def _ctcss_scenario(expected, actual, n_signal):
    t = Both([expected])
    t.settle_noise_floor()
    n0 = len(t.axc)
    t.feed(carrier(tone_audio(actual, n_signal)))
    return t, n0


def test_squelch_good_ctcss():
    t, n0 = _ctcss_scenario(TONES[5], TONES[5], 20 * B)
    ob = t.open_batches()[n0:]
    assert ob[0] and all(ob)                              # opens within the first batch (fast detector) and stays open
    s = t.e.stats(0, 0)
    assert s.ctcss_count > 0 and s.no_ctcss_count == 0
    t.close()


def test_squelch_wrong_ctcss():
    t, n0 = _ctcss_scenario(TONES[7], TONES[0], 20 * B)
    assert not any(t.open_batches()[n0:])
    s = t.e.stats(0, 0)
    assert s.ctcss_count == 0 and s.no_ctcss_count > 0
    t.close()


def test_squelch_close_ctcss():
    t, n0 = _ctcss_scenario(TONES[7], TONES[5], 20 * B)
    ob = t.open_batches()[n0:]
    assert ob[0]                                          # the fast (0.05 s) detector cannot separate 79.7 from 85.4 Hz
    assert not any(ob[5:])                                # the slow (0.4 s) detector can: closed within 3000 samples
    s = t.e.stats(0, 0)
    assert s.ctcss_count == 0 and s.no_ctcss_count > 0
    t.close()


# ------------------------------------------------------------------------------------------------ test_ctcss.cpp:122-155
def _all_detectors(present, seed, noise=0.2):
    """One channel per standard tone, all hearing the same audio; returns which channels opened (= detector has_tone)."""
    t = Both(TONES)
    t.settle_noise_floor()
    n0 = len(t.axc)
    audio = tone_audio(present, 8 * B, noise=noise, seed=seed)
    t.feed(carrier(audio))
    opened = np.array([any(t.open_batches(c)[n0 + 4:]) for c in range(len(TONES))])   # after the slow detector's first verdicts
    t.close()
    return opened


def _check_all(opened, present):
    for c, tone in enumerate(TONES):
        if abs(np.float32(tone) - np.float32(present)) < 5:
            continue
        assert not opened[c], f"tone {tone} found, expected {present}"


def test_ctcss_creation_without_a_tone_does_not_gate_the_squelch():
    t = Both([0])
    t.settle_noise_floor()
    n0 = len(t.axc)
    t.feed(np.full(2 * B, RAW_SIGNAL, np.float32))
    assert all(t.open_batches()[n0:])
    s = t.e.stats(0, 0)
    assert s.ctcss_count == 0 and s.no_ctcss_count == 0
    t.close()


def test_ctcss_no_signal():
    """Upstream feeds exact zeros to the detector.  Through the channel loop the only way to give the detectors nothing is no
    carrier at all (an unmodulated carrier leaves a ~5e-8 DC residue of the AGC division in the audio, which the reference's
    "largest and above the mean" rule (ctcss.cpp:139-156) attributes to the lowest tones: same behaviour on CPU and GPU, checked by
    the parity half of every feed)."""
    t = Both(TONES)
    t.settle_noise_floor()
    n0 = len(t.axc)
    t.feed(np.full(8 * B, RAW_NO_SIGNAL, np.float32))
    assert not any(any(t.open_batches(c)[n0:]) for c in range(len(TONES)))
    for c in (0, 12, 50):
        s = t.e.stats(0, c)
        assert s.ctcss_count == 0 and s.no_ctcss_count == 0 and s.open_count == 0
    t.close()


def test_ctcss_has_tone():
    opened = _all_detectors(TONES[0], 2)
    _check_all(opened, TONES[0])
    assert opened[0]


def test_ctcss_has_non_standard_tone():
    tone = (TONES[3] + TONES[4]) / 2
    t = Both([tone] + TONES)
    t.settle_noise_floor()
    n0 = len(t.axc)
    t.feed(carrier(tone_audio(tone, 8 * B, noise=0.2, seed=3)))
    opened = np.array([any(t.open_batches(c)[n0 + 4:]) for c in range(1 + len(TONES))])
    t.close()
    assert opened[0]
    _check_all(opened[1:], tone)


@pytest.mark.parametrize("k", range(0, len(TONES), 6))
def test_ctcss_has_each_standard_tone(k):
    opened = _all_detectors(TONES[k], 100 + k)
    _check_all(opened, TONES[k])
    assert opened[k]
