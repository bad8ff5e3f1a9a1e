// Shared device-side building blocks of the K1 kernels: compile-time twiddles, the fully unrolled radix-2 DIT
// register FFT, sample conversion (reference src/rtl_airband.cpp:316-324,402-455) and the TMA bulk-copy helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <utility>

#include "../../include/airband_b200.h"

namespace k1 {

// ---------------------------------------------------------------------------------------------------------------
// compile-time twiddles
// ---------------------------------------------------------------------------------------------------------------
constexpr double kPi = 3.14159265358979323846264338327950288;

constexpr double taylor_sin(double x) {  // |x| <= pi/4
    double x2 = x * x, term = x, sum = x;
    for (int i = 1; i < 14; ++i) {
        term *= -x2 / double((2 * i) * (2 * i + 1));
        sum += term;
    }
    return sum;
}
constexpr double taylor_cos(double x) {  // |x| <= pi/4
    double x2 = x * x, term = 1.0, sum = 1.0;
    for (int i = 1; i < 14; ++i) {
        term *= -x2 / double((2 * i - 1) * (2 * i));
        sum += term;
    }
    return sum;
}
// cos / sin of 2*pi*k/L for 0 <= k < L, L a power of two >= 8, by exact octant reduction
constexpr double cos2pi(int k, int L) {
    int q = L / 4, quad = k / q, r = k % q;
    double c = (r <= L / 8) ? taylor_cos(2.0 * kPi * r / L) : taylor_sin(2.0 * kPi * (q - r) / L);
    double s = (r <= L / 8) ? taylor_sin(2.0 * kPi * r / L) : taylor_cos(2.0 * kPi * (q - r) / L);
    return quad == 0 ? c : quad == 1 ? -s : quad == 2 ? -c : s;
}
constexpr double sin2pi(int k, int L) {
    int q = L / 4, quad = k / q, r = k % q;
    double c = (r <= L / 8) ? taylor_cos(2.0 * kPi * r / L) : taylor_sin(2.0 * kPi * (q - r) / L);
    double s = (r <= L / 8) ? taylor_sin(2.0 * kPi * r / L) : taylor_cos(2.0 * kPi * (q - r) / L);
    return quad == 0 ? s : quad == 1 ? c : quad == 2 ? -s : -c;
}

template <int K, int L>
struct Tw {  // forward transform: W_L^K = exp(-2*pi*i*K/L)
    static constexpr float re = (float)cos2pi(K, L);
    static constexpr float im = (float)(-sin2pi(K, L));
};

template <int R>
__host__ __device__ constexpr int brev(int i) {
    int r = 0;
    for (int b = 1; b < R; b <<= 1) {
        r = (r << 1) | (i & 1);
        i >>= 1;
    }
    return r;
}

// ---------------------------------------------------------------------------------------------------------------
// register FFT: radix-2 decimation in time over R complex registers.
// Logical array A[i] lives in v[brev(i)]; after run(), X[k] is in v[brev(k)] ... i.e. the SAME storage rule on
// input and output: input sample x[i] must be placed in v[i] (A[brev(i)] = x[i] is what DIT needs), and
// output bin X[k] is read from v[brev<R>(k)]... see note below.
// ---------------------------------------------------------------------------------------------------------------
template <int K, int L>
__device__ __forceinline__ void bfly(float2& a, float2& b) {
    if constexpr (K == 0) {
        const float ax = a.x, ay = a.y;
        a.x = ax + b.x;
        a.y = ay + b.y;
        b.x = ax - b.x;
        b.y = ay - b.y;
    } else if constexpr (4 * K == L) {  // W = -j : W*b = (b.y, -b.x)
        const float ax = a.x, ay = a.y, bx = b.x, by = b.y;
        a.x = ax + by;
        a.y = ay - bx;
        b.x = ax - by;
        b.y = ay + bx;
    } else {
        constexpr float wr = Tw<K, L>::re, wi = Tw<K, L>::im;
        const float tr = fmaf(-wi, b.y, fmaf(wr, b.x, a.x));
        const float ti = fmaf(wi, b.x, fmaf(wr, b.y, a.y));
        b.x = fmaf(2.0f, a.x, -tr);
        b.y = fmaf(2.0f, a.y, -ti);
        a.x = tr;
        a.y = ti;
    }
}

// DIT stage of span L on the logical array A[i] = v[brev<R>(i)] (so that feeding natural-order samples into v[]
// IS the bit-reversed load DIT needs, and the natural-order result X[k] = A[k] sits in v[brev<R>(k)]).
template <int R, int L, int... Is>
__device__ __forceinline__ void dit_stage(float2 (&v)[R], std::integer_sequence<int, Is...>) {
    (bfly<Is % (L / 2), L>(v[brev<R>((Is / (L / 2)) * L + Is % (L / 2))], v[brev<R>((Is / (L / 2)) * L + Is % (L / 2) + L / 2)]), ...);
}
template <int R, int L>
__device__ __forceinline__ void dit_from(float2 (&v)[R]) {
    dit_stage<R, L>(v, std::make_integer_sequence<int, R / 2>{});
    if constexpr (L < R) dit_from<R, L * 2>(v);
}
// in: v[i] = x[i] (natural order).  out: X[k] = v[brev<R>(k)].
template <int R>
__device__ __forceinline__ void reg_fft(float2 (&v)[R]) {
    dit_from<R, 2>(v);
}

__host__ __device__ constexpr int ilog2(int x) { return x <= 1 ? 0 : 1 + ilog2(x / 2); }

// ---------------------------------------------------------------------------------------------------------------
// sample conversion (reference src/rtl_airband.cpp:316-324,402-455).  Returns the raw integer-valued (or float)
// sample; the 1/full-scale factor lives in the window table.
//   U8 : (b - 127.5)   built exactly with the 2^22 magic number (0x4A800000 | b<<1 == 2^22 + b, ulp 0.5)
//   S8 : b             (levels_s8[(uint8)i] = i/128; the /128 is folded into the table, exact)
//   S16: x             ((float)buf2[k]; scale = 1/fullscale folded into the table)
//   F32: x
// ---------------------------------------------------------------------------------------------------------------
template <int SFMT>
__device__ __forceinline__ float2 load_sample(const unsigned char* tile, int byte_off) {
    if constexpr (SFMT == ABG_SFMT_U8) {
        const unsigned int u = *reinterpret_cast<const unsigned short*>(tile + byte_off);
        const float i = __uint_as_float(((u & 0xFFu) << 1) | 0x4A800000u) - 4194431.5f;
        const float q = __uint_as_float(((u >> 7) & 0x1FEu) | 0x4A800000u) - 4194431.5f;
        return make_float2(i, q);
    } else if constexpr (SFMT == ABG_SFMT_S8) {
        const unsigned int u = *reinterpret_cast<const unsigned short*>(tile + byte_off);
        const int i = (int)(signed char)(u & 0xFFu), q = (int)(signed char)(u >> 8);
        return make_float2(__int_as_float(0x4B400000 + i) - 12582912.0f, __int_as_float(0x4B400000 + q) - 12582912.0f);
    } else if constexpr (SFMT == ABG_SFMT_S16) {
        const unsigned int u = *reinterpret_cast<const unsigned int*>(tile + byte_off);
        const int i = (int)(short)(u & 0xFFFFu), q = (int)(short)(u >> 16);
        return make_float2(__int_as_float(0x4B400000 + i) - 12582912.0f, __int_as_float(0x4B400000 + q) - 12582912.0f);
    } else {
        return *reinterpret_cast<const float2*>(tile + byte_off);
    }
}
template <int SFMT>
__host__ __device__ constexpr int bytes_per_cplx() {
    return SFMT == ABG_SFMT_U8 || SFMT == ABG_SFMT_S8 ? 2 : SFMT == ABG_SFMT_S16 ? 4 : 8;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }


// ---- TMA bulk copy global -> shared with an mbarrier (one-shot, parity 0) ----
__device__ __forceinline__ void tma_tile_load(unsigned long long* mbar, unsigned char* dst, const unsigned char* src, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(mbar))
                 : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned long long* mbar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait0(unsigned long long* mbar) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(mbar))
            : "memory");
    }
}

}  // namespace k1
