"""Pins the oracle's FFT stand-in (oracle/fft32.cpp) for fftwf_plan_dft_1d(FFTW_FORWARD) / fftwf_execute
(reference src/rtl_airband.cpp:262-264,460): unnormalised forward DFT, sign -1.  fftw3f itself is not installed
and the reference has no test at this boundary ("parity unpinned" there); the pin is numpy's complex128 FFT."""
import numpy as np
import pytest

import oracle_py as op


@pytest.mark.parametrize("n", [256, 512, 1024, 2048, 4096, 8192])
def test_fft_matches_complex128(n):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    y = op.fft(x)
    ref = np.fft.fft(x.astype(np.complex128))
    rel_rms = np.sqrt(np.mean(np.abs(y - ref) ** 2) / np.mean(np.abs(ref) ** 2))
    assert rel_rms < 3e-7, rel_rms  # FFTW float level (the VideoCore engine the reference also ships is 0.3-4.4 ppm)
    assert np.max(np.abs(y - ref)) / np.max(np.abs(ref)) < 2e-6


@pytest.mark.parametrize("n", [256, 2048])
def test_fft_impulse_and_tone(n):
    x = np.zeros(n, np.complex64)
    x[3] = 1.0
    y = op.fft(x)
    k = np.arange(n)
    assert np.allclose(y, np.exp(-2j * np.pi * 3 * k / n), atol=1e-6)  # forward sign is e^{-i...}
    x = np.exp(2j * np.pi * 17 * k / n).astype(np.complex64)
    y = op.fft(x)
    assert abs(y[17] - n) < 1e-3 * n and np.abs(np.delete(y, 17)).max() < 1e-3 * n


def test_fft_linearity():
    n = 1024
    rng = np.random.default_rng(0)
    a = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    b = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    assert np.allclose(op.fft(a + b), op.fft(a) + op.fft(b), atol=2e-4)
