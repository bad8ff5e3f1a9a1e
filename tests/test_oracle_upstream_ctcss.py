"""Ports of the reference's CTCSS unit tests (reference src/test_ctcss.cpp:122-155) against the oracle's CTCSS
(restated, and the reference's own ctcss.cpp where oracle/_ref is built).  The upstream Noise source is seeded
from std::random_device (generate_signal.cpp:41-46) and therefore not reproducible; here it is a seeded
numpy normal(0, 0.1) * amplitude, as SURVEY.md §4 prescribes."""
import math

import numpy as np
import pytest

import oracle_py as op

VARIANTS = ["restated"] + (["ref"] if op.available("ref") else [])
SR = 8000
SLOW = int(SR * 0.4)
TONES = [67.0, 69.3, 71.9, 74.4, 77.0, 79.7, 82.5, 85.4, 88.5, 91.5, 94.8, 97.4, 100.0, 103.5, 107.2, 110.9, 114.8, 118.8, 123.0, 127.3,
         131.8, 136.5, 141.3, 146.2, 150.0, 151.4, 156.7, 159.8, 162.2, 165.5, 167.9, 171.3, 173.8, 177.3, 179.9, 183.5, 186.2, 189.9,
         192.8, 196.6, 199.5, 203.5, 206.5, 210.7, 218.1, 225.7, 229.1, 233.6, 241.8, 250.3, 254.1]


def signal(tone, n, seed, noise=0.2, ampl=0.2):
    k = np.arange(1, n + 1, dtype=np.float64)
    x = np.zeros(n)
    if tone:
        x += np.float32(ampl) * np.sin(2 * math.pi * k * float(np.float32(tone)) / SR)
    if noise:
        x += noise * np.random.default_rng(seed).normal(0.0, 0.1, n)
    return x.astype(np.float32)


def run(variant, detect_tone, x):
    c = op.CtcssHarness(detect_tone, SR, SLOW, variant)
    assert c.enabled()
    for v in x:
        if c.enough():
            break
        c.sample(float(v))
    assert c.enough()
    return c.has_tone()


def check_all(variant, x, present):
    for t in TONES:
        if abs(np.float32(t) - np.float32(present)) < 5:
            continue
        assert not run(variant, t, x), f"tone {t} found, expected {present}"
    if present:
        assert run(variant, present, x), f"expected tone {present} not found"


@pytest.mark.parametrize("variant", VARIANTS)
def test_creation(variant):
    assert not op.CtcssHarness(0.0, SR, SLOW, variant).enabled()


@pytest.mark.parametrize("variant", VARIANTS)
def test_no_signal(variant):
    check_all(variant, signal(0, SLOW, 1, noise=0.0), 0)


@pytest.mark.parametrize("variant", VARIANTS)
def test_has_tone(variant):
    check_all(variant, signal(TONES[0], SLOW, 2), TONES[0])


@pytest.mark.parametrize("variant", VARIANTS)
def test_has_non_standard_tone(variant):
    t = (TONES[0] + TONES[0]) / 2
    check_all(variant, signal(t, SLOW, 3), t)


@pytest.mark.parametrize("variant", ["restated"])
def test_has_each_standard_tone(variant):
    for i, t in enumerate(TONES):
        check_all(variant, signal(t, SLOW, 100 + i), t)
