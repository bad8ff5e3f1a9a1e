"""Ports of the reference's own Squelch unit tests (reference src/test_squelch.cpp:51-281) driven through the
oracle's per-sample harness — against the restated classes and, where oracle/_ref is built, against the
reference's own squelch.cpp.  These are the only behavioural pins upstream holds for the squelch.
Tone generator restates reference src/generate_signal.cpp:32-35 (first sample is n = 1)."""
import math

import numpy as np
import pytest

import oracle_py as op

VARIANTS = ["restated"] + (["ref"] if op.available("ref") else [])
RAW_NO_SIGNAL = 0.05
RAW_SIGNAL = 0.75
STANDARD_TONES = [67.0, 69.3, 71.9, 74.4, 77.0, 79.7, 82.5, 85.4]


class Tone:
    def __init__(self, sample_rate, freq, ampl):
        self.sr, self.f, self.a, self.n = sample_rate, np.float32(freq), np.float32(ampl), 0

    def get(self):
        self.n += 1
        return float(np.float32(float(self.a) * math.sin(2 * math.pi * self.n * float(self.f) / self.sr)))


def settle_noise_floor(s):
    # send_samples_for_noise_floor(), test_squelch.cpp:39-45
    n = 0
    while s.noise_level() > 1.01 * RAW_NO_SIGNAL:
        s.raw(RAW_NO_SIGNAL)
        n += 1
        assert n < 100000
    assert s.noise_level() <= 1.01 * RAW_NO_SIGNAL
    assert RAW_SIGNAL > s.squelch_level()


@pytest.mark.parametrize("variant", VARIANTS)
def test_default_object(variant):
    assert op.SquelchHarness(variant).open_count() == 0


@pytest.mark.parametrize("variant", VARIANTS)
def test_noise_floor(variant):
    s = op.SquelchHarness(variant)
    assert s.noise_level() > 10.0 * RAW_NO_SIGNAL
    this = s.noise_level()
    while True:
        last = this
        for _ in range(25):
            s.raw(RAW_NO_SIGNAL)
        this = s.noise_level()
        assert this <= last
        if this == last:
            break
    assert s.noise_level() < 1.01 * RAW_NO_SIGNAL


@pytest.mark.parametrize("variant", VARIANTS)
def test_normal_operation(variant):
    s = op.SquelchHarness(variant)
    settle_noise_floor(s)
    for _ in range(500):
        if s.is_open():
            break
        s.raw(RAW_SIGNAL)
    assert s.is_open() and s.should_process_audio()
    for _ in range(1000):
        s.raw(RAW_SIGNAL)
    assert s.is_open() and s.should_process_audio()
    for _ in range(100):
        if not s.is_open():
            break
        s.raw(RAW_NO_SIGNAL)
    assert not s.is_open() and not s.should_process_audio()


@pytest.mark.parametrize("variant", VARIANTS)
def test_dead_spot(variant):
    s = op.SquelchHarness(variant)
    settle_noise_floor(s)
    for _ in range(500):
        if s.is_open():
            break
        s.raw(RAW_SIGNAL)
    assert s.is_open()
    for _ in range(1000):
        s.raw(RAW_SIGNAL)
    assert s.is_open() and s.should_process_audio()
    for _ in range(50):
        s.raw(RAW_NO_SIGNAL)
        assert s.is_open() and s.should_process_audio()
    for _ in range(1000):
        s.raw(RAW_SIGNAL)
        assert s.is_open() and s.should_process_audio()


@pytest.mark.parametrize("variant", VARIANTS)
def test_should_process_audio(variant):
    s = op.SquelchHarness(variant)
    settle_noise_floor(s)
    for _ in range(500):
        if s.is_open():
            break
        assert not s.should_process_audio()
        s.raw(RAW_SIGNAL)
    assert s.is_open() and s.should_process_audio()
    for _ in range(100):
        if not s.is_open():
            break
        assert s.should_process_audio()
        s.raw(RAW_NO_SIGNAL)
    assert not s.is_open() and not s.should_process_audio()


def _until_audio(s):
    for _ in range(500):
        if s.should_process_audio():
            break
        s.raw(RAW_SIGNAL)


@pytest.mark.parametrize("variant", VARIANTS)
def test_good_ctcss(variant):
    tone, sr = STANDARD_TONES[5], 8000
    s = op.SquelchHarness(variant)
    s.set_ctcss(tone, sr)
    settle_noise_floor(s)
    sig = Tone(sr, tone, 0.2)
    _until_audio(s)
    assert not s.is_open() and s.should_process_audio()
    for _ in range(500):
        if s.is_open():
            break
        s.audio(sig.get())
        s.raw(RAW_SIGNAL)
    assert s.is_open() and s.should_process_audio()
    for _ in range(100000):
        s.audio(sig.get())
        s.raw(RAW_SIGNAL)
        assert s.is_open()
    assert s.should_process_audio()
    assert s.ctcss_count() > 0 and s.no_ctcss_count() == 0


@pytest.mark.parametrize("variant", VARIANTS)
def test_wrong_ctcss(variant):
    actual, expected, sr = STANDARD_TONES[0], STANDARD_TONES[7], 8000
    s = op.SquelchHarness(variant)
    s.set_ctcss(expected, sr)
    settle_noise_floor(s)
    sig = Tone(sr, actual, 0.2)
    _until_audio(s)
    assert s.should_process_audio() and not s.is_open()
    for _ in range(100000):
        s.audio(sig.get())
        s.raw(RAW_SIGNAL)
        assert not s.is_open()
    assert s.should_process_audio()
    assert s.ctcss_count() == 0 and s.no_ctcss_count() > 0


@pytest.mark.parametrize("variant", VARIANTS)
def test_close_ctcss(variant):
    actual, expected, sr = STANDARD_TONES[5], STANDARD_TONES[7], 8000
    s = op.SquelchHarness(variant)
    s.set_ctcss(expected, sr)
    settle_noise_floor(s)
    sig = Tone(sr, actual, 0.2)
    _until_audio(s)
    assert s.should_process_audio() and not s.is_open()
    for _ in range(500):
        if s.is_open():
            break
        s.audio(sig.get())
        s.raw(RAW_SIGNAL)
        assert s.should_process_audio()
    assert s.is_open()  # the fast (0.05 s) detector cannot separate 79.7 from 85.4 Hz
    for _ in range(3000):
        if not s.is_open():
            break
        s.audio(sig.get())
        s.raw(RAW_SIGNAL)
        assert s.should_process_audio()
    assert not s.is_open()  # the slow (0.4 s) detector can
    for _ in range(100000):
        s.audio(sig.get())
        s.raw(RAW_SIGNAL)
        assert not s.is_open()
    assert s.should_process_audio()
    assert s.ctcss_count() == 0 and s.no_ctcss_count() > 0
