// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// Single-precision forward complex DFT, X[k] = sum_n x[n] * exp(-2*pi*i*k*n/N), unnormalised, out of place.
// Stands in for the reference's fftwf_plan_dft_1d(N, in, out, FFTW_FORWARD, FFTW_MEASURE) + fftwf_execute()
// (reference src/rtl_airband.cpp:262-264,460).  FFTW3 itself (libfftw3f, unpinned distro package, 3.3.10 on
// the reference's Debian bookworm image) is not installed here and cannot be fetched, so PARITY IS UNPINNED at
// this one call: no reference test exercises it.  We pin it ourselves against numpy/scipy complex128 in
// tests/test_oracle_fft.py (relative rms error must stay at FFTW's float level, ~1e-7).
#pragma once
#include <cstddef>
#include <vector>

namespace abo {

class Fft32 {
   public:
    explicit Fft32(size_t n);  // n must be a power of two, 2..65536
    size_t size() const { return n_; }
    // in/out: interleaved (re,im) float pairs, n each.  in is not modified.  in != out.
    void forward(const float* in, float* out);

   private:
    size_t n_;
    std::vector<float> tw_;    // per-stage twiddle tables
    std::vector<size_t> tw_off_;
    std::vector<float> work_;  // ping-pong buffer
};

}  // namespace abo
