// ORACLE — TEST INFRASTRUCTURE ONLY (see fft32.h).  Stockham autosort DIF FFT, radix 4 with one radix-2
// stage when log2(N) is odd.  Twiddles are evaluated in double and rounded once to float.
#include "fft32.h"

#include <cmath>
#include <cstring>
#include <stdexcept>

namespace abo {

Fft32::Fft32(size_t n) : n_(n) {
    if (n < 2 || n > 65536 || (n & (n - 1))) throw std::invalid_argument("Fft32: size must be a power of two");
    work_.resize(2 * n);
    // stage list: radix-4 while length >= 4, then radix-2 if 2 remains
    size_t len = n;
    while (len >= 4) {
        tw_off_.push_back(tw_.size());
        size_t n1 = len / 4;
        for (size_t p = 0; p < n1; p++) {
            for (int m = 1; m <= 3; m++) {
                double ang = -2.0 * M_PI * (double)(m * p) / (double)len;
                tw_.push_back((float)cos(ang));
                tw_.push_back((float)sin(ang));
            }
        }
        len /= 4;
    }
    if (len == 2) tw_off_.push_back(tw_.size());  // radix-2 tail stage has s = n/2, len = 2 -> only p = 0, w = 1
}

namespace {

inline void r4_stage(size_t len, size_t s, const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ tw) {
    const size_t n1 = len / 4;
    for (size_t p = 0; p < n1; p++) {
        const float w1r = tw[6 * p + 0], w1i = tw[6 * p + 1];
        const float w2r = tw[6 * p + 2], w2i = tw[6 * p + 3];
        const float w3r = tw[6 * p + 4], w3i = tw[6 * p + 5];
        const float* xa = x + 2 * s * p;
        const float* xb = x + 2 * s * (p + n1);
        const float* xc = x + 2 * s * (p + 2 * n1);
        const float* xd = x + 2 * s * (p + 3 * n1);
        float* y0 = y + 2 * s * (4 * p + 0);
        float* y1 = y + 2 * s * (4 * p + 1);
        float* y2 = y + 2 * s * (4 * p + 2);
        float* y3 = y + 2 * s * (4 * p + 3);
        for (size_t q = 0; q < s; q++) {
            const float ar = xa[2 * q], ai = xa[2 * q + 1];
            const float br = xb[2 * q], bi = xb[2 * q + 1];
            const float cr = xc[2 * q], ci = xc[2 * q + 1];
            const float dr = xd[2 * q], di = xd[2 * q + 1];
            const float apcr = ar + cr, apci = ai + ci;
            const float amcr = ar - cr, amci = ai - ci;
            const float bpdr = br + dr, bpdi = bi + di;
            // j*(b-d) = (-(bi-di), (br-dr))
            const float jr = -(bi - di), ji = (br - dr);
            y0[2 * q] = apcr + bpdr;
            y0[2 * q + 1] = apci + bpdi;
            const float t1r = amcr - jr, t1i = amci - ji;
            const float t2r = apcr - bpdr, t2i = apci - bpdi;
            const float t3r = amcr + jr, t3i = amci + ji;
            y1[2 * q] = t1r * w1r - t1i * w1i;
            y1[2 * q + 1] = t1r * w1i + t1i * w1r;
            y2[2 * q] = t2r * w2r - t2i * w2i;
            y2[2 * q + 1] = t2r * w2i + t2i * w2r;
            y3[2 * q] = t3r * w3r - t3i * w3i;
            y3[2 * q + 1] = t3r * w3i + t3i * w3r;
        }
    }
}

inline void r2_tail(size_t s, const float* __restrict__ x, float* __restrict__ y) {
    // len == 2: single p = 0, twiddle 1
    const float* xa = x;
    const float* xb = x + 2 * s;
    float* y0 = y;
    float* y1 = y + 2 * s;
    for (size_t q = 0; q < s; q++) {
        const float ar = xa[2 * q], ai = xa[2 * q + 1];
        const float br = xb[2 * q], bi = xb[2 * q + 1];
        y0[2 * q] = ar + br;
        y0[2 * q + 1] = ai + bi;
        y1[2 * q] = ar - br;
        y1[2 * q + 1] = ai - bi;
    }
}

}  // namespace

void Fft32::forward(const float* in, float* out) {
    // count stages to decide which buffer the first stage must write so the last lands in `out`
    size_t nstages = tw_off_.size();
    float* bufs[2] = {out, work_.data()};
    // stage i writes bufs[(nstages - 1 - i) & 1]; so the last stage (i = nstages-1) writes bufs[0] = out
    const float* src = in;
    size_t len = n_, s = 1, i = 0;
    while (len >= 4) {
        float* dst = bufs[(nstages - 1 - i) & 1];
        r4_stage(len, s, src, dst, tw_.data() + tw_off_[i]);
        src = dst;
        len /= 4;
        s *= 4;
        i++;
    }
    if (len == 2) {
        float* dst = bufs[(nstages - 1 - i) & 1];
        r2_tail(s, src, dst);
    }
}

}  // namespace abo
