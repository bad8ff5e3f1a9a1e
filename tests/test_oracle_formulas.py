"""Cross-checks of the small config-time / helper formulas: the oracle's C restatements vs independent Python
mirrors (airband_b200.config) vs worked examples from the reference's own configs."""
import math

import numpy as np
import pytest

import oracle_py as op
from airband_b200 import config as cm
from airband_b200 import workloads as wl

L = op.lib("restated")


def test_bins_worked_examples():
    # config/basic_multichannel.conf: cf 120.0, 119.5 / 120.225 MHz, default fft_size 512 -> 411, 44
    assert L.abo_calc_bin(119500000, 120000000, 2560000, 512) == 411
    assert L.abo_calc_bin(120225000, 120000000, 2560000, 512) == 44
    # config/noaa.conf: cf 162.482, 2.4 Msps, fft_size 1024 (integer bin width 2343, not 2343.75)
    got = [L.abo_calc_bin(f, 162482000, 2400000, 1024) for f in (162400000, 162425000, 162450000, 162475000, 162500000, 162525000, 162550000)]
    assert got == [989, 999, 1010, 1021, 8, 18, 29]


@pytest.mark.parametrize("sr,n", [(2560000, 512), (2400000, 1024), (10000000, 4096), (2560000, 2048), (3200000, 8192)])
def test_bin_and_dphi_python_mirror(sr, n):
    rng = np.random.default_rng(sr // n)
    for _ in range(200):
        cf = int(rng.integers(100_000_000, 400_000_000))
        f = cf + int(rng.integers(-sr // 2 + 1, sr // 2))
        assert cm.calc_bin(f, cf, sr, n) == L.abo_calc_bin(f, cf, sr, n)
        for w in (8000, 16000):
            assert cm.calc_dm_dphi(f, cf, sr, w) == L.abo_calc_dm_dphi(f, cf, sr, w)


def test_dbfs_level():
    for n in (256, 512, 2048, 4096):
        for db in (-10.0, -30.0, -47.0, -70.0):
            a, b = cm.dbfs_to_level(db, n), L.abo_dbfs_to_level(db, n)
            assert abs(a - b) <= 2e-6 * abs(b)  # numpy log10(float32) vs libm log10f may differ by an ulp
            assert abs(L.abo_level_to_dbfs(b, n) - db) < 1e-3
    assert abs(L.abo_dbfs_to_level(-30.0, 512) - 0.5587) < 1e-3  # SURVEY.md §7 hard part 3


def test_alpha_and_hop():
    assert abs(L.abo_default_alpha(16000) - math.exp(-1.0 / (16000 * 2e-4))) < 1e-7
    assert cm.default_alpha(8000) == L.abo_default_alpha(8000)
    assert cm.hop_samples(2560000, 8000) == 320 and cm.hop_samples(2400000, 16000) == 150 and cm.hop_samples(10000000, 16000) == 625


@pytest.mark.parametrize("n", [256, 512, 2048, 8192])
def test_window_is_blackman_harris7(n):
    # reference src/rtl_airband.cpp:335-351: float literals held in double, evaluated in double, stored as float
    o = op.Oracle(wl.cfg2(1, 1, fft_size=n))
    w = o.window()
    a = [np.float64(np.float32(v)) for v in (0.27105140069342, 0.43329793923448, 0.21812299954311, 0.06592544638803,
                                             0.01081174209837, 0.00077658482522, 0.00001388721735)]
    i = np.arange(n, dtype=np.float64)
    x = sum(((-1) ** k) * a[k] * np.cos(2.0 * k * np.pi * i / (n - 1)) for k in range(7))
    assert np.max(np.abs(w - x)) < 1.5e-7
    assert abs(w.sum() / n - a[0]) < 5e-3 * a[0] + 1.0 / n  # coherent gain a0 ~ 0.271
    assert np.allclose(w, w[::-1], atol=1e-7)


def test_sincos_lut_and_fm():
    s, c = op.C.c_float(), op.C.c_float()
    for phi in (0, 1, 0x8000, 0x123456, 0xFFFFFF):
        L.abo_sincosf_lut(phi, op.C.byref(s), op.C.byref(c))
        ang = 2 * math.pi * phi / (1 << 24)
        assert abs(s.value - math.sin(ang)) < 4e-4 and abs(c.value - math.cos(ang)) < 4e-4  # linear interp of 256 pts
    rng = np.random.default_rng(1)
    for _ in range(200):
        y, x = rng.standard_normal(2)
        assert abs(L.abo_fast_atan2(y, x) - math.atan2(y, x)) < 0.08  # piecewise-rational approximation
    assert L.abo_fast_atan2(0.0, 0.0) == 0.0
    # discriminator: constant rotation by theta per sample -> theta / pi
    th = 0.3
    a = (math.cos(th), math.sin(th))
    assert abs(L.abo_polar_disc_fast(a[0], a[1], 1.0, 0.0) - th / math.pi) < 0.03
    assert abs(L.abo_fm_quadri_demod(a[0], a[1], 1.0, 0.0) - (math.sin(th) / 2 / math.pi)) < 1e-6
