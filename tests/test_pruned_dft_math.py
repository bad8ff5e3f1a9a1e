"""The algebra behind the output-pruned K1 (rtlsdr-airband_b200/csrc/k1_pruned.cu), checked with numpy: with N = R1*M1,
X[b] = sum_{c<M1} W_N^(c*b) * Y_c[b mod R1], Y_c = R1-point DFT of the column x[c + M1*n1]; the coefficient of column
c = PAIR*l + 64*m + p factors into a per-lane part W^(PAIR*l*b) and a warp-uniform part W^((64*m+p)*b); and the window can be
folded into the first radix-2 stage of the column FFT.  (The CUDA kernel itself is tested against the oracle on the GPU.)"""
import numpy as np
import pytest


@pytest.mark.parametrize("n,r1", [(2048, 8), (2048, 16), (512, 8), (8192, 16), (256, 8)])
def test_column_split_reproduces_selected_bins(n, r1):
    rng = np.random.default_rng(n + r1)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    m1 = n // r1
    ref = np.fft.fft(x)
    cols = x.reshape(r1, m1)                 # cols[n1, c] = x[c + M1*n1]
    y = np.fft.fft(cols, axis=0)             # y[k1, c] = Y_c[k1]
    w = np.exp(-2j * np.pi * np.arange(n) / n)
    for b in rng.integers(0, n, 12):
        c = np.arange(m1)
        got = np.sum(w[(c * b) % n] * y[b % r1, c])
        assert abs(got - ref[b]) <= 1e-9 * max(1.0, abs(ref[b]))


def test_lane_and_warp_uniform_factors():
    n, r1 = 2048, 8
    m1, ncol, pair = n // r1, (n // 32) // r1, 2
    w = np.exp(-2j * np.pi * np.arange(n) / n)
    for b in (0, 1, 205, 1023, 2047):
        for lane in (0, 1, 17, 31):
            for j in range(ncol):
                col = pair * lane + (pair * 32) * (j // pair) + (j % pair)       # column owned by (lane, j)
                assert col < m1
                coff = (pair * 32) * (j // pair) + (j % pair)                     # part that does not depend on the lane
                lhs = w[(col * b) % n]
                rhs = w[((pair * lane) * b) % n] * w[(coff * b) % n]
                assert abs(lhs - rhs) < 1e-12
    # every column is owned exactly once
    owned = sorted(pair * l + (pair * 32) * (j // pair) + (j % pair) for l in range(32) for j in range(ncol))
    assert owned == list(range(m1))


def test_window_folds_into_the_first_radix2_stage():
    r1 = 8
    rng = np.random.default_rng(1)
    x = rng.standard_normal(r1) + 1j * rng.standard_normal(r1)
    win = rng.random(r1)
    ref = np.fft.fft(x * win)
    # first DIT stage pairs sample n1 with n1 + R1/2: a' = xa*wa + xb*wb, b' = xa*wa - xb*wb; then an R1/2-point DFT of
    # the sums gives the even bins and one of the (twiddled) differences the odd bins
    h = r1 // 2
    s = x[:h] * win[:h] + x[h:] * win[h:]
    d = (x[:h] * win[:h] - x[h:] * win[h:]) * np.exp(-2j * np.pi * np.arange(h) / r1)
    assert np.allclose(np.fft.fft(s), ref[0::2]) and np.allclose(np.fft.fft(d), ref[1::2])
