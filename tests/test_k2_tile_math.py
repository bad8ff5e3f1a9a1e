"""CPU checks of the identities the lane-parallel K2 tiles rely on (rtlsdr-airband_b200/csrc/k2_demod.cu: k2_am_tile,
k2_nfm_tile).  A tile computes, once per 8 / 16 samples and across the lanes, what the reference loop computes sample by sample
(reference src/squelch.cpp:195-276,501-514, src/filters.cpp:49-64, src/rtl_airband.cpp:510-518); it may only do so where the
re-arranged form gives the SAME IEEE single-precision value.  Each test walks the sequential form and the tile's form in numpy
float32 / integer arithmetic (one rounding per operation, like the -fmad=false build) and demands bit equality."""
import numpy as np
import pytest

F = np.float32
RNG = np.random.default_rng(20260923)


def bits(a):
    return np.asarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("n", [8, 16])
def test_low_signal_counter_from_a_ballot(n):
    """low_signal_count_ after sample k = samples since the last one at or above the level: bit arithmetic on the ballot of
    (x >= level), as in both tiles, against the sequential counter (squelch.cpp:234-245)."""
    for _ in range(500):
        x = RNG.random(n).astype(F)
        lvl = F(RNG.choice([0.05, 0.5, 0.95]))
        low_in = int(RNG.integers(0, 90))
        seq, low = [], low_in
        for k in range(n):
            low = 0 if x[k] >= lvl else low + 1
            seq.append(low)
        ballot = sum(1 << k for k in range(n) if x[k] >= lvl)
        for k in range(n):
            z = ballot & ((2 << k) - 1)
            tile = (k - (z.bit_length() - 1)) if z else (low_in + k + 1)   # 31 - clz(z) == bit_length - 1
            assert tile == seq[k]


def test_capped_average_as_one_compare_against_plus_infinity():
    """capped = (capped >= cap && sample >= cap) ? cap : min(cap, 0.99 capped + t)  (squelch.cpp:506-513) with the two
    conditions folded into capped >= xc, xc = cap where the sample reaches the cap and +inf elsewhere."""
    cap = F(0.37)
    for _ in range(2000):
        c = F(RNG.choice([0.1, 0.37, 0.3699999, 0.5]))
        x = F(RNG.choice([0.2, 0.37, 0.9]))
        t = F(x * F(0.01))
        c2 = min(cap, F(F(c * F(0.99)) + t))
        ref = cap if (c >= cap and x >= cap) else c2
        xc = cap if x >= cap else F(np.inf)
        tile = cap if c >= xc else c2
        assert bits(ref) == bits(tile)


@pytest.mark.parametrize("n", [8, 16])
def test_shortcuts_for_fully_capped_and_never_capped_tiles(n):
    """k2_am_tile: with every sample at or above the cap and the average capped on entry the average stays at the cap; with no
    sample at or above the cap it is min(cap, 0.99 avg + t) - both bit-identical to the general recurrence."""
    cap = F(0.4)

    def general(c, xs):
        out = []
        for x in xs:
            t = F(x * F(0.01))
            c2 = min(cap, F(F(c * F(0.99)) + t))
            c = cap if (c >= cap and x >= cap) else c2
            out.append(c)
        return out
    for _ in range(300):
        hi = (cap + RNG.random(n).astype(F)).astype(F)            # all >= cap
        for c0 in (cap, F(0.47)):                                  # capped (or above a cap that just dropped) on entry
            assert all(bits(v) == bits(cap) for v in general(c0, hi))
        lo = (RNG.random(n).astype(F) * F(0.399)).astype(F)        # all < cap
        c, short = F(RNG.random() * 0.5), []
        c_in = c
        for x in lo:
            c = min(cap, F(F(c * F(0.99)) + F(x * F(0.01))))
            short.append(c)
        assert [int(bits(v)) for v in general(c_in, lo)] == [int(bits(v)) for v in short]


def test_derotation_phase_of_lane_k():
    """dm_phi advances by dm_dphi and is masked to 24 bits every sample (rtl_airband.cpp:516); lane k of a tile uses
    (phi0 + k * dphi) & 0xffffff computed in 32-bit unsigned arithmetic."""
    for _ in range(2000):
        phi0 = int(RNG.integers(0, 1 << 24))
        dphi = int(RNG.integers(0, 1 << 24))
        phi = phi0
        for k in range(17):
            assert ((phi0 + k * dphi) & 0xffffffff) & 0xffffff == phi
            phi = (phi + dphi) & 0xffffff


@pytest.mark.parametrize("n", [8, 16])
def test_notch_feed_forward_half_per_lane_recursive_half_serial(n):
    """NotchFilter::apply (filters.cpp:49-64): y = d0*x[k] - d1*x[k-1] + d0*x[k-2] + d1*y[k-1] - d2*y[k-2], evaluated left to
    right.  The tile computes A[k] = (d0*x[k] - d1*x[k-1]) + d0*x[k-2] per lane and walks y = (A[k] + d1*y1) - d2*y0."""
    d0, d1, d2 = F(0.9893), F(1.9771), F(0.9786)
    for _ in range(300):
        x = (RNG.standard_normal(n) * 0.3).astype(F)
        x1, x2, y1, y2 = (F(v) for v in RNG.standard_normal(4) * 0.3)   # nx1 (older), nx2, ny1 (older), ny2
        # sequential, as the reference writes it
        sx1, sx2, sy1, sy2, seq = x1, x2, y1, y2, []
        for k in range(n):
            x0 = sx1
            sx1, sx2 = sx2, x[k]
            y0 = sy1
            sy1 = sy2
            sy2 = F(F(F(F(F(d0 * sx2) - F(d1 * sx1)) + F(d0 * x0)) + F(d1 * sy1)) - F(d2 * y0))
            seq.append(sy2)
        # tile
        xm1 = np.concatenate(([x2], x[:-1])).astype(F)
        xm2 = np.concatenate(([x1, x2], x[:-2])).astype(F)
        A = [F(F(F(d0 * x[k]) - F(d1 * xm1[k])) + F(d0 * xm2[k])) for k in range(n)]
        ty1, ty2, tile = y1, y2, []
        for k in range(n):
            y0 = ty1
            ty1 = ty2
            ty2 = F(F(A[k] + F(d1 * ty1)) - F(d2 * y0))
            tile.append(ty2)
        assert [int(bits(v)) for v in seq] == [int(bits(v)) for v in tile]


@pytest.mark.parametrize("n", [8, 16])
def test_lowpass_real_and_imaginary_parts_are_independent_recurrences(n):
    """LowpassFilter::apply (filters.cpp:146-163) filters the complex sample with real coefficients: the real and the imaginary
    part never mix, so the tile walks them in the two half-warps at once; each is the same scalar recurrence."""
    c0, c1, gain = F(-0.6413), F(1.5610), F(49.8)
    z = (RNG.standard_normal(n) + 1j * RNG.standard_normal(n)).astype(np.complex64)

    def scalar(xs, st):
        x1, x2, y1, y2 = st
        out = []
        for v in xs:
            x0 = x1
            x1, x2 = x2, F(v / gain)
            y0 = y1
            y1 = y2
            y2 = F(F(F(F(x0 + x2) + F(F(2.0) * x1)) + F(c0 * y0)) + F(c1 * y1))
            out.append(y2)
        return out
    st_r = tuple(F(v) for v in RNG.standard_normal(4))
    st_i = tuple(F(v) for v in RNG.standard_normal(4))
    # complex form as the reference writes it (std::complex<float> * float and + act per component)
    x1, x2 = complex(st_r[0], st_i[0]), complex(st_r[1], st_i[1])
    y1, y2 = complex(st_r[2], st_i[2]), complex(st_r[3], st_i[3])
    ref = []
    for v in z:
        x0 = x1
        x1 = x2
        x2 = complex(F(v.real / gain), F(v.imag / gain))
        y0 = y1
        y1 = y2

        def comp(part):
            g = (lambda c: F(c.real)) if part == 0 else (lambda c: F(c.imag))
            return F(F(F(F(g(x0) + g(x2)) + F(F(2.0) * g(x1))) + F(c0 * g(y0))) + F(c1 * g(y1)))
        y2 = complex(comp(0), comp(1))
        ref.append(y2)
    re, im = scalar(z.real, st_r), scalar(z.imag, st_i)
    assert [int(bits(v.real)) for v in ref] == [int(bits(v)) for v in re]
    assert [int(bits(v.imag)) for v in ref] == [int(bits(v)) for v in im]


def test_am_agc_recurrence_without_the_clip_inside_a_committed_tile():
    """rtl_airband.cpp:553-562: agc = 0.995 agc + 0.005 x where x > level, then waveout = (wavein[j-100] - agc) / (1.5 agc) and
    the |waveout| > 0.8 clip (waveout *= 0.85, agc *= 1.15).  k2_am_tile walks the AGC without the clip and commits only when
    every sample's |n| < 1.199985 * agc (n = numerator), i.e. |n / (1.5 agc)| < 0.79999: then no sample clipped."""
    for _ in range(300):
        n = 16
        lvl = F(0.1)
        x = (0.5 + 0.2 * RNG.standard_normal(n)).astype(F)
        lag = (0.5 + 0.2 * RNG.standard_normal(n)).astype(F)
        a_seq, a_tile, clipped, calm = F(0.5), F(0.5), False, True
        for k in range(n):
            if x[k] > lvl:
                a_seq = F(F(a_seq * F(0.995)) + F(x[k] * F(0.005)))
                a_tile = F(F(a_tile * F(0.995)) + F(x[k] * F(0.005)))
            w = F(F(lag[k] - a_seq) / F(a_seq * F(1.5)))
            if abs(w) > F(0.8):
                clipped = True
                a_seq = F(a_seq * F(1.15))
            calm = calm and abs(F(lag[k] - a_tile)) < F(a_tile * F(1.199985))
        if calm:
            assert not clipped and bits(a_seq) == bits(a_tile)
