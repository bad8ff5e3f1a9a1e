// tc_probe — hardware probe for the encodings the tensor-core K1 (rtlsdr-airband_b200/csrc/k1_tc.cu) relies on:
// tcgen05.mma kind::i8 with a U8/S8 A operand and an S8 B operand, both K-major WITHOUT swizzle; an A operand whose rows
// are 16 bytes apart in shared memory (so that "frame f+q" is the same buffer shifted by q*16 bytes: the sliding-window
// overlap of consecutive FFT frames costs no data movement); cp.async-filled A tiles made visible to the async proxy;
// bulk-copied B tiles; tcgen05.ld of the S32 accumulator.  One variant per process (a faulting variant must not poison
// the others):   tc_probe <variant>     prints OK / MISMATCH against a CPU integer reference.
// Not part of the product; built by `make probe`.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../rtlsdr-airband_b200/csrc/tc_ptx.cuh"

using namespace tc;

constexpr int HC = 4;            // 16-byte chunks per hop row
constexpr int HOPB = HC * 16;    // hop bytes
constexpr int ROWS = 136;        // 128 + 8 halo rows
constexpr int SPITCH = (ROWS + 1) * 16;  // column pitch of the transposed tile (odd number of 16-byte units)
constexpr int MAXKS = 16;

struct ProbeArgs {
    const unsigned char* raw;   // [ROWS][HOPB]
    const signed char* bimg;    // [nks][2][NC][16]
    int32_t* out;               // [128][NC]
    int32_t* status;            // [8]
    int nks, nc, a_signed, swap_lbo_sbo;
    int q[MAXKS], j[MAXKS];
};

__global__ void __launch_bounds__(160) probe_kernel(const ProbeArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* at = smem;                          // [HC][SPITCH]
    unsigned char* bs = smem + ((HC * SPITCH + 127) & ~127);  // [nks][2][NC][16]
    __shared__ __align__(8) unsigned long long bars[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_b = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);
    if (tid == 0) {
        mbar_init(bar_b, 1);
        mbar_init(bar_mma, 1);
        fence_mbar_init();
    }
    if (warp == 4) tmem_alloc<256>(smem_u32(&tmem_base_s));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (tid < 128) {
        for (int i = tid; i < ROWS * HC; i += 128) {
            const int r = i / HC, jj = i % HC;
            cp_async16(smem_u32(at + jj * SPITCH + r * 16), a.raw + (size_t)i * 16);
        }
        cp_async_wait_all();
        fence_proxy_async();
    }
    const uint32_t bbytes = (uint32_t)(a.nks * 2 * a.nc * 16);
    if (tid == 0) {
        mbar_arrive_expect_tx(bar_b, bbytes);
        bulk_g2s(smem_u32(bs), a.bimg, bbytes, bar_b);
    }
    __syncthreads();
    if (warp == 4 && lane == 0) {
        if (!mbar_wait(bar_b, 0, 1ll << 28)) a.status[0] = 1;
        fence_proxy_async();
        tc_fence_after();
        const uint32_t idesc = idesc_i8(128, a.nc, a.a_signed, 1);
        for (int ks = 0; ks < a.nks; ks++) {
            const uint32_t a_addr = smem_u32(at + a.j[ks] * SPITCH + a.q[ks] * 16);
            const uint32_t b_addr = smem_u32(bs + (size_t)ks * 2 * a.nc * 16);
            uint32_t a_lbo = SPITCH, a_sbo = 128, b_lbo = (uint32_t)a.nc * 16, b_sbo = 128;
            if (a.swap_lbo_sbo) {
                uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t;
                t = b_lbo; b_lbo = b_sbo; b_sbo = t;
            }
            mma_i8(tmem, smem_desc_noswizzle(a_addr, a_lbo, a_sbo), smem_desc_noswizzle(b_addr, b_lbo, b_sbo), idesc, ks > 0 ? 1u : 0u);
        }
        mma_commit(bar_mma);
    }
    if (warp < 4) {
        if (!mbar_wait(bar_mma, 0, 1ll << 28)) {
            if (lane == 0) a.status[1] = 1;
        } else {
            tc_fence_after();
            const int row = warp * 32 + lane;
            for (int c0 = 0; c0 < a.nc; c0 += 8) {
                uint32_t r[8];
                tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
                tmem_ld_wait();
                for (int k = 0; k < 8; k++) a.out[(size_t)row * a.nc + c0 + k] = (int32_t)r[k];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc<256>(tmem);
    if (tid == 0) a.status[2] = 1;  // reached the end
}

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);     \
            return 2;                                                                           \
        }                                                                                       \
    } while (0)

int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    ProbeArgs a{};
    a.nc = 64; a.nks = 1; a.a_signed = 0; a.swap_lbo_sbo = 0;
    const char* what = "";
    switch (variant) {
        case 0: what = "1 k-step, q=0, U8 x S8, N=64"; break;
        case 1: what = "1 k-step, q=0, LBO/SBO swapped"; a.swap_lbo_sbo = 1; break;
        case 2: what = "1 k-step, q=3 (row shift 48 bytes)"; a.q[0] = 3; break;
        case 3: what = "1 k-step, q=5, j=2"; a.q[0] = 5; a.j[0] = 2; break;
        case 4: what = "8 k-steps accumulate, mixed (q, j)"; a.nks = 8; for (int k = 0; k < 8; k++) { a.q[k] = k; a.j[k] = (k & 1) * 2; } break;
        case 5: what = "S8 x S8, 4 k-steps"; a.a_signed = 1; a.nks = 4; for (int k = 0; k < 4; k++) { a.q[k] = 7 - k; a.j[k] = (k & 1) * 2; } break;
        case 6: what = "N=48, 4 k-steps"; a.nc = 48; a.nks = 4; for (int k = 0; k < 4; k++) { a.q[k] = k; a.j[k] = 2 - (k & 1) * 2; } break;
        case 7: what = "N=16, 2 k-steps"; a.nc = 16; a.nks = 2; a.q[1] = 1; a.j[1] = 2; break;
        case 8: what = "N=256, 16 k-steps"; a.nc = 256; a.nks = 16; for (int k = 0; k < 16; k++) { a.q[k] = k & 7; a.j[k] = (k & 1) * 2; } break;
        default: printf("unknown variant\n"); return 1;
    }
    std::vector<unsigned char> raw(ROWS * HOPB);
    std::vector<signed char> bimg((size_t)a.nks * 2 * a.nc * 16);
    uint32_t s = 12345u + variant;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return s >> 24; };
    for (auto& v : raw) v = (unsigned char)rnd();
    for (auto& v : bimg) v = (signed char)rnd();
    std::vector<int32_t> ref((size_t)128 * a.nc, 0);
    for (int m = 0; m < 128; m++)
        for (int n = 0; n < a.nc; n++) {
            long long acc = 0;
            for (int ks = 0; ks < a.nks; ks++)
                for (int t = 0; t < 32; t++) {
                    const int byte = (m + a.q[ks]) * HOPB + a.j[ks] * 16 + t;
                    const int av = a.a_signed ? (int)(signed char)raw[byte] : (int)raw[byte];
                    const int bv = bimg[(((size_t)ks * 2 + t / 16) * a.nc + n) * 16 + t % 16];
                    acc += av * bv;
                }
            ref[(size_t)m * a.nc + n] = (int32_t)acc;
        }
    unsigned char* d_raw; signed char* d_b; int32_t *d_out, *d_st;
    CK(cudaMalloc(&d_raw, raw.size()));
    CK(cudaMalloc(&d_b, bimg.size()));
    CK(cudaMalloc(&d_out, ref.size() * 4));
    CK(cudaMalloc(&d_st, 32));
    CK(cudaMemcpy(d_raw, raw.data(), raw.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_b, bimg.data(), bimg.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(d_out, 0xFF, ref.size() * 4));
    CK(cudaMemset(d_st, 0, 32));
    a.raw = d_raw; a.bimg = d_b; a.out = d_out; a.status = d_st;
    const size_t smem = ((HC * SPITCH + 127) & ~127) + bimg.size() + 128;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_kernel<<<1, 160, smem>>>(a);
    CK(cudaGetLastError());
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("variant %d (%s): KERNEL FAULT %s\n", variant, what, cudaGetErrorString(e));
        return 3;
    }
    std::vector<int32_t> out(ref.size());
    int32_t st[8];
    CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(st, d_st, 32, cudaMemcpyDeviceToHost));
    size_t bad = 0, first = 0;
    for (size_t i = 0; i < ref.size(); i++)
        if (out[i] != ref[i]) {
            if (!bad) first = i;
            bad++;
        }
    printf("variant %d (%s): %s  mismatches=%zu/%zu status=[%d %d %d]", variant, what, bad ? "MISMATCH" : "OK", bad, ref.size(), st[0], st[1], st[2]);
    if (bad) printf("  first at row %zu col %zu: got %d want %d; out[0]=%d ref[0]=%d", first / a.nc, first % a.nc, out[first], ref[first], out[0], ref[0]);
    printf("\n");
    return bad ? 4 : 0;
}
