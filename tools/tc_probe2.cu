// tc_probe2 — second hardware probe for k1_tc.cu: (1) how fast tcgen05.mma kind::i8 (M=128, N=64, K=32, operands in shared
// memory) runs back to back with the no-swizzle 16-byte-row layout versus the 128-byte-swizzle layout, and (2) whether the
// row-shift trick (frame f+q = the same tile, start address advanced by q rows) also works with 128-byte swizzled rows, and
// with which `base_offset` convention.   tc_probe2 <mode>     mode 0: throughput table;  modes 1..: swizzled correctness
// variants (one per process).  Not part of the product; built by `make`.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../rtlsdr-airband_b200/csrc/tc_ptx.cuh"

using namespace tc;

__host__ __device__ constexpr uint64_t smem_desc_sw128(uint32_t addr, uint32_t sbo_bytes, uint32_t base_offset) {
    // K-major, SWIZZLE_128B: LBO unused (1), SBO = byte distance between 8-row groups, layout type 2 in bits [61,64)
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) |
           ((uint64_t)(base_offset & 7u) << 49) | (2ull << 61);
}

// ------------------------------------------------------------------------------------------------------------------ throughput
struct TpArgs {
    long long* cycles;  // [n_variants]
    int reps;
};
__global__ void __launch_bounds__(160) tp_kernel(const TpArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 160 * 1024 / 4; i += 160) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u * (i & 3);
    if (tid == 0) {
        mbar_init(smem_u32(&bar), 1);
        fence_mbar_init();
    }
    if (warp == 4) tmem_alloc<512>(smem_u32(&tmem_base_s));
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (warp == 4) {
        const uint32_t leader = elect_one();
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 96 * 1024);
        uint32_t phase = 0;
        for (int v = 0; v < 8; v++) {
            // v0: no-swizzle, rows 16 B apart (k1_tc layout), column pitch 2160   v1: same, start shifted by 3 rows
            // v2: 128B swizzle, aligned start                                     v3: 128B swizzle, start shifted by 3 rows
            // v4: no-swizzle N=48   v5: no-swizzle N=128   v6: no-swizzle N=256   v7: 128B swizzle N=256
            const int n = v == 4 ? 48 : (v == 5 ? 128 : (v >= 6 ? 256 : 64));
            const uint32_t idesc = idesc_i8(128, n, 0, 1);
            const bool sw = (v == 2 || v == 3 || v == 7);
            const uint32_t shift = (v == 1) ? 48u : (v == 3 ? 384u : 0u);
            const uint64_t ad = sw ? smem_desc_sw128(a0 + shift, 1024, (shift >> 7) & 7) : smem_desc_noswizzle(a0 + shift, 2160, 128);
            const uint64_t bd = sw ? smem_desc_sw128(b0, 1024, 0) : smem_desc_noswizzle(b0, (uint32_t)n * 16, 128);
            __syncwarp();
            const long long t0 = clock64();
            if (leader) {
                for (int i = 0; i < a.reps; i++) {
                    // walk A like the kernel does (different 32-byte K slices), 4 partial accumulators
                    const uint32_t koff = sw ? (uint32_t)((i & 3) * 32 + ((i >> 2) % 5) * 17408) >> 4 : (uint32_t)(((i % 20) * 2 * 2160) >> 4);
                    mma_i8(tmem + (uint32_t)((i & 3) * 64) % (n > 64 ? 1u : 256u), ad + koff, bd, idesc, i >= 4 ? 1u : 0u);
                }
                mma_commit(smem_u32(&bar));
            }
            __syncwarp();
            mbar_wait_spin(smem_u32(&bar), phase);
            phase ^= 1;
            const long long t1 = clock64();
            if (leader) a.cycles[v] = t1 - t0;
            tc_fence_after();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------------------ swizzled correctness
constexpr int ROWS = 144;  // 128 + 16 halo rows, 128-byte rows, one 128-byte column block
struct CkArgs {
    const unsigned char* raw;  // [ROWS][128] logical rows
    const signed char* b;      // [64][128] logical rows (K-major)
    int32_t* out;              // [128][64]
    int q, k32, base_mode;     // row shift, 32-byte K slice inside the 128-byte row, 0: base_offset 0, 1: (addr >> 7) & 7
};
__global__ void __launch_bounds__(160) ck_kernel(const CkArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* at = smem;              // [ROWS][128] swizzled
    unsigned char* bs = smem + 32 * 1024;  // [64][128] swizzled
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // 16-byte chunk c of row r lives at r*128 + ((c ^ (r & 7)) * 16)   (Swizzle<3,4,3> on byte addresses)
    for (int i = tid; i < ROWS * 8; i += 160) {
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(at + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(a.raw + r * 128 + c * 16);
    }
    for (int i = tid; i < 64 * 8; i += 160) {
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(bs + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(a.b + r * 128 + c * 16);
    }
    if (tid == 0) {
        mbar_init(smem_u32(&bar), 1);
        fence_mbar_init();
    }
    if (warp == 4) tmem_alloc<64>(smem_u32(&tmem_base_s));
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (warp == 4 && lane == 0) {
        const uint32_t a_addr = smem_u32(at) + (uint32_t)a.q * 128u + (uint32_t)a.k32 * 32u;
        const uint32_t b_addr = smem_u32(bs) + (uint32_t)a.k32 * 32u;
        const uint32_t bo = a.base_mode ? ((a_addr >> 7) & 7u) : 0u;
        mma_i8(tmem, smem_desc_sw128(a_addr, 1024, bo), smem_desc_sw128(b_addr, 1024, 0), idesc_i8(128, 64, 0, 1), 0u);
        mma_commit(smem_u32(&bar));
    }
    if (warp < 4) {
        mbar_wait_spin(smem_u32(&bar), 0);
        tc_fence_after();
        const int row = warp * 32 + lane;
        for (int c0 = 0; c0 < 64; c0 += 8) {
            uint32_t r[8];
            tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
            tmem_ld_wait();
            for (int k = 0; k < 8; k++) a.out[(size_t)row * 64 + c0 + k] = (int32_t)r[k];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc<64>(tmem);
}

#define CK(x)                                                                               \
    do {                                                                                    \
        cudaError_t e_ = (x);                                                               \
        if (e_ != cudaSuccess) {                                                            \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return 2;                                                                       \
        }                                                                                   \
    } while (0)

int main(int argc, char** argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    if (mode == 0) {
        long long* d_c;
        CK(cudaMalloc(&d_c, 8 * sizeof(long long)));
        TpArgs a{d_c, 2000};
        CK(cudaFuncSetAttribute(tp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        tp_kernel<<<1, 160, 200 * 1024>>>(a);
        CK(cudaGetLastError());
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("throughput kernel fault: %s\n", cudaGetErrorString(e));
            return 3;
        }
        long long c[8];
        CK(cudaMemcpy(c, d_c, sizeof(c), cudaMemcpyDeviceToHost));
        const char* names[8] = {"no-swizzle N=64 aligned", "no-swizzle N=64 start +3 rows", "swizzle128 N=64 aligned", "swizzle128 N=64 start +3 rows",
                                "no-swizzle N=48", "no-swizzle N=128", "no-swizzle N=256", "swizzle128 N=256"};
        for (int v = 0; v < 8; v++) printf("throughput %-32s %8.1f cycles per MMA (2000 back to back, one CTA)\n", names[v], (double)c[v] / 2000.0);
        return 0;
    }
    // correctness: mode = 1 + q*8 + k32*2 + base_mode  with q in {0, 3, 8, 11}
    const int m = mode - 1;
    const int qs[4] = {0, 3, 8, 11};
    CkArgs a{};
    a.base_mode = m & 1;
    a.k32 = (m >> 1) & 3;
    a.q = qs[(m >> 3) & 3];
    std::vector<unsigned char> raw(ROWS * 128);
    std::vector<signed char> b(64 * 128);
    uint32_t s = 999u + mode;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return s >> 24; };
    for (auto& v : raw) v = (unsigned char)rnd();
    for (auto& v : b) v = (signed char)rnd();
    std::vector<int32_t> ref(128 * 64);
    for (int r = 0; r < 128; r++)
        for (int n = 0; n < 64; n++) {
            int acc = 0;
            for (int t = 0; t < 32; t++) acc += (int)raw[(r + a.q) * 128 + a.k32 * 32 + t] * (int)b[n * 128 + a.k32 * 32 + t];
            ref[r * 64 + n] = acc;
        }
    unsigned char* d_raw; signed char* d_b; int32_t* d_out;
    CK(cudaMalloc(&d_raw, raw.size())); CK(cudaMalloc(&d_b, b.size())); CK(cudaMalloc(&d_out, ref.size() * 4));
    CK(cudaMemcpy(d_raw, raw.data(), raw.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_b, b.data(), b.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(d_out, 0xFF, ref.size() * 4));
    a.raw = d_raw; a.b = d_b; a.out = d_out;
    CK(cudaFuncSetAttribute(ck_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
    ck_kernel<<<1, 160, 48 * 1024>>>(a);
    CK(cudaGetLastError());
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("swizzle128 q=%d k32=%d base_mode=%d: KERNEL FAULT %s\n", a.q, a.k32, a.base_mode, cudaGetErrorString(e));
        return 3;
    }
    std::vector<int32_t> out(ref.size());
    CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t i = 0; i < ref.size(); i++) bad += out[i] != ref[i];
    printf("swizzle128 q=%d k32=%d base_offset=%s: %s (%zu / %zu mismatches)\n", a.q, a.k32, a.base_mode ? "(addr>>7)&7" : "0", bad ? "MISMATCH" : "OK", bad, ref.size());
    return bad ? 4 : 0;
}
