// K1 (tensor-core) — sample conversion + window + DFT of the configured bins as ONE integer GEMM on the 5th-generation
// tensor cores (tcgen05.mma kind::i8, accumulators in TMEM), sm_100a.  Same inputs, same outputs as k1_pruned.cu /
// k1_fft.cu; replaces the same reference code (reference src/rtl_airband.cpp:402-455 convert+window, :460
// fftwf_execute, :483-489 bin extraction) for 8-bit sample formats (U8 / S8: RTL-SDR, file input).
//
// The reference transforms N points per frame and reads C bins.  Written out for one bin b of frame f,
//     X[f,b] = sum_n  level(raw[f*hop + n]) * window[n] * W_N^(n*b)            (level(u) = (u - 127.5)/127.5 for U8)
// is a dense contraction  [frames x 2N raw BYTES] x [2N x 2C coefficients]: the A operand is the device's raw byte
// stream itself (8-bit integers are exact tensor-core operands: no conversion instruction is ever executed), the B operand
// is window*twiddle quantised to ND signed 8-bit digits (ND = 4: 30-bit fixed point, |error| <= 2^-31 of the largest
// coefficient, below the FP32 FFT's own rounding), accumulated EXACTLY in S32 and recombined in the epilogue:
//     X = (Gmax / 2Q) * (2 * sum_d 256^(ND-1-d) * acc_d  -  255 * sum_k q_k)        (the -127.5 offset, folded out)
//
// Sliding windows without data movement: consecutive frames start hop_bytes apart (84 % overlap at N = 2048, hop 320).
// The raw tile is stored in shared memory "transposed by 16-byte chunk": column j (of hop_bytes/16) holds bytes
// [16j, 16j+16) of every hop-row r at AT[j][r], rows 16 bytes apart.  That IS the canonical K-major no-swizzle operand
// layout (core matrix = 8 rows x 16 bytes), and the rows of frame f+q are the rows of frame f shifted by q*16 bytes, so
// the MMA for K bytes [q*hop_bytes + 16j, +32) simply takes descriptor start address AT + j*S + q*16: every raw byte is
// fetched from HBM once and written to shared memory once, and the tensor core reads it for each of the ~N/hop frames
// containing it.
//
// One CTA per SM, persistent, dynamic tile queue (tile = 128 consecutive frames of one device), warp-specialised:
//   warps 0-3  A producers: cp.async (LDGSTS) 16-byte chunks global -> transposed tile, double-buffered
//   warps 4-7  epilogue: tcgen05.ld the S32 accumulators (double-buffered in TMEM), recombine digits in int64,
//              one double multiply, |X| with the reference's rounding (rtl_airband.cpp:484), coalesced stores
//   warp 8     tile scheduler (atomic counter) + B loader: the device's coefficient table streams through a ring of
//              bulk-copy (TMA) stages, one pass per tile, out of L2
//   warp 9     TMEM allocation + the single MMA-issuing thread
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include <algorithm>

#include "../../include/airband_b200.h"
#include "abg_internal.h"
#include "tc_ptx.cuh"

namespace {
using namespace tc;
extern long long* g_tc_trace;

constexpr int TC_THREADS = 320;
constexpr int TC_CTRL_BYTES = 1024;
constexpr int TC_MAX_BSTAGES = 16;
constexpr int TC_INFO_SLOTS = 4;

struct TileInfo {
    const unsigned char* src;  // first byte of the tile's first frame
    int nf;                    // valid frames (1..128); <= 0: end of work
    int pos;                   // row of win/iqin of the first frame
    int gbase;                 // first channel index
    int nch;                   // channels
    int tab;                   // coefficient table
    int pad;
};

struct TcArgs {
    const K1Dev* devs;
    const int32_t* tab_of_dev;
    const signed char* btab;
    const long long* sq;
    float* win;
    float2* iqin;
    int* counter;
    int32_t* status;
    double cscale;
    int Gp, n_devices, tiles_per_dev, total_tiles;
    int K, hop_bytes, HC, S, NC, ND, C2p, KBS, NSTB, a_signed;
    int mul, off;
    uint16_t aoff[512];        // per k-step start-address offset of the A operand, 16-byte units (K <= 16384)
    int stage_in_row;          // every B stage's k-steps lie inside one hop-row (HC/2 is a multiple of KBS)
    long long* trace;          // measurement aid (ABG_K1_TC_TRACE): clock64 stamps [cta][role 0..3][tile 0..15][event 0..3], or null
    int dbg_skip;              // measurement aid (ABG_K1_TC_SKIP): 1 = do not copy coefficients, 2 = do not copy samples, 4 = issue no MMA (results invalid)
    int rotate;                // start every CTA's K loop at a different coefficient block (exact integer sums commute)
};

struct Ctrl {
    unsigned long long full_a[2], empty_a[2], tmem_full[2], tmem_empty[2];
    unsigned long long full_b[TC_MAX_BSTAGES], empty_b[TC_MAX_BSTAGES];
    unsigned long long info_full[TC_INFO_SLOTS], info_empty[TC_INFO_SLOTS];
    TileInfo info[TC_INFO_SLOTS];
    uint32_t tmem_base;
};
static_assert(sizeof(Ctrl) <= TC_CTRL_BYTES, "control block too large");

// Every wait below is mbar_wait_spin: bounded inside one asm statement, traps instead of hanging (tc_ptx.cuh).
// NACC independent partial accumulators per tile: consecutive MMAs into ONE accumulator serialise on the read-modify-write
// of TMEM (measured: ~130 cycles per 128x64x32 MMA, 4x its issue time), so k-step k accumulates into partial k % NACC and the
// epilogue adds the partials (integer sums are exact and commute).
template <int NACC, int KBS>
__global__ void __launch_bounds__(TC_THREADS, 1) k1_tc_kernel(const TcArgs a) {
    constexpr int TMEM_COLS = 512;  // the whole tensor memory of the SM: 2 tiles x NACC partials x NCS columns
    extern __shared__ __align__(1024) unsigned char smem[];
    Ctrl* c = reinterpret_cast<Ctrl*>(smem);
    unsigned char* abuf = smem + TC_CTRL_BYTES;
    const int abuf_bytes = a.HC * a.S;
    unsigned char* bring = smem + TC_CTRL_BYTES + ((2 * abuf_bytes + 127) & ~127);
    const int stage_bytes = KBS * a.NC * 32;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int ACC_STRIDE = TMEM_COLS / 2;
    constexpr int NCS = ACC_STRIDE / NACC;  // column stride between the partial accumulators of a tile (>= NC)

    if (tid == 0) {
        for (int i = 0; i < 2; i++) {
            mbar_init(smem_u32(&c->full_a[i]), 128);
            mbar_init(smem_u32(&c->empty_a[i]), 1);
            mbar_init(smem_u32(&c->tmem_full[i]), 1);
            mbar_init(smem_u32(&c->tmem_empty[i]), 128);
        }
        for (int i = 0; i < TC_MAX_BSTAGES; i++) {
            mbar_init(smem_u32(&c->full_b[i]), 1);
            mbar_init(smem_u32(&c->empty_b[i]), 1);
        }
        for (int i = 0; i < TC_INFO_SLOTS; i++) {
            mbar_init(smem_u32(&c->info_full[i]), 1);
            mbar_init(smem_u32(&c->info_empty[i]), 128);
        }
        fence_mbar_init();
    }
    if (warp == 9) tmem_alloc<TMEM_COLS>(smem_u32(&c->tmem_base));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = c->tmem_base;
    const int NKB = (a.K / 32) / KBS;  // B stages per tile
    // All CTAs stream the same (or a few) coefficient tables out of L2 at the same pace: starting each CTA at a different
    // block spreads the requests over the L2 slices (exact integer sums commute, so the K order is free).
    const int kb_rot = a.rotate ? (int)((blockIdx.x * 2654435761u >> 8) % (unsigned)NKB) : 0;

// stage-level stamps of tile 3: event e of stage st (< 16) goes to slot 32 + 2*st + e of the role's 64-slot block (tiles 8..15 unused then)
#define TC_TRACE_STAGE(role, it, st, e)                                                                          \
    do {                                                                                                         \
        if (a.trace && (it) == 3 && (st) < 16 && (threadIdx.x & 31) == 0) a.trace[((size_t)blockIdx.x * 4 + (role)) * 64 + 32 + 2 * (st) + (e)] = clock64(); \
    } while (0)
#define TC_TRACE(role, it, ev)                                                                                   \
    do {                                                                                                         \
        if (a.trace && (it) < 8 && (threadIdx.x & 31) == 0) a.trace[(((size_t)blockIdx.x * 4 + (role)) * 16 + (it)) * 4 + (ev)] = clock64(); \
    } while (0)
    if (warp < 4) {
        // ================= A producers =================
        for (int it = 0;; ++it) {
            const int slot = it & (TC_INFO_SLOTS - 1);
            mbar_wait_spin(smem_u32(&c->info_full[slot]), (it / TC_INFO_SLOTS) & 1);
            const TileInfo ti = c->info[slot];
            if (ti.nf <= 0) break;
            const int buf = it & 1;
            if (warp == 0) TC_TRACE(0, it, 0);
            mbar_wait_spin(smem_u32(&c->empty_a[buf]), ((it >> 1) & 1) ^ 1);
            if (warp == 0) TC_TRACE(0, it, 1);
            const uint32_t dst0 = smem_u32(abuf + buf * abuf_bytes);
            const int total_chunks = ((ti.nf - 1) * a.hop_bytes + a.K) >> 4;
            int r = tid / a.HC, j = tid - r * a.HC;
            const int dr = 128 / a.HC, dj = 128 - dr * a.HC;
            for (int i = tid; i < total_chunks; i += 128) {
                if (!(a.dbg_skip & 2)) cp_async16(dst0 + j * a.S + r * 16, ti.src + (size_t)i * 16);
                r += dr;
                j += dj;
                if (j >= a.HC) {
                    j -= a.HC;
                    r++;
                }
            }
            if (warp == 0) TC_TRACE(0, it, 2);
            cp_async_wait_all();
            fence_proxy_async();
            mbar_arrive(smem_u32(&c->full_a[buf]));
            if (warp == 0) TC_TRACE(0, it, 3);
        }
    } else if (warp < 8) {
        // ================= epilogue =================
        const int lane_base = (warp & 3) * 32;
        const int row = lane_base + lane;
        for (int it = 0;; ++it) {
            const int slot = it & (TC_INFO_SLOTS - 1);
            mbar_wait_spin(smem_u32(&c->info_full[slot]), (it / TC_INFO_SLOTS) & 1);
            const TileInfo ti = c->info[slot];
            if (ti.nf <= 0) break;
            const int acc = it & 1;
            if (warp == 4) TC_TRACE(1, it, 0);
            mbar_wait_spin(smem_u32(&c->tmem_full[acc]), (it >> 1) & 1);
            if (warp == 4) TC_TRACE(1, it, 1);
            tc_fence_after();
            const uint32_t t0 = tmem + (uint32_t)(acc * ACC_STRIDE) + ((uint32_t)lane_base << 16);
            const long long* sq = a.sq + (size_t)ti.tab * a.C2p;
            const size_t orow = (size_t)(ti.pos + row) * a.Gp + ti.gbase;
            const bool valid = row < ti.nf;
            const bool vec = ((ti.gbase & 3) == 0) && ((a.Gp & 3) == 0);
            for (int og = 0; og < a.C2p; og += 8) {
                long long v[8];
                for (int d = 0; d < a.ND; d++) {
                    uint32_t rr[NACC][8];
#pragma unroll
                    for (int pa = 0; pa < NACC; pa++) tmem_ld8(t0 + (uint32_t)(pa * NCS + d * a.C2p + og), rr[pa]);
                    tmem_ld_wait();
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        int32_t sum = (int32_t)rr[0][k];
#pragma unroll
                        for (int pa = 1; pa < NACC; pa++) sum += (int32_t)rr[pa][k];  // the full sum fits S32 (|sum| <= 2N * 255 * 128)
                        v[k] = (d == 0) ? (long long)sum : (v[k] << 8) + (long long)sum;
                    }
                }
                float x[8];
#pragma unroll
                for (int k = 0; k < 8; k++) x[k] = (float)((double)(a.mul * v[k] - a.off * __ldg(sq + og + k)) * a.cscale);
                if (valid) {
                    const int c0 = og >> 1;  // first channel of this group of four
                    float m[4];
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        m[k] = sqrtf(__fadd_rn(__fmul_rn(x[2 * k], x[2 * k]), __fmul_rn(x[2 * k + 1], x[2 * k + 1])));  // rtl_airband.cpp:484
                    if (vec && c0 + 4 <= ti.nch) {
                        *reinterpret_cast<float4*>(a.win + orow + c0) = make_float4(m[0], m[1], m[2], m[3]);
                        float4* q = reinterpret_cast<float4*>(a.iqin + orow + c0);
                        q[0] = make_float4(x[0], x[1], x[2], x[3]);
                        q[1] = make_float4(x[4], x[5], x[6], x[7]);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            if (c0 + k < ti.nch) {
                                a.win[orow + c0 + k] = m[k];
                                a.iqin[orow + c0 + k] = make_float2(x[2 * k], x[2 * k + 1]);
                            }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(smem_u32(&c->tmem_empty[acc]));
            mbar_arrive(smem_u32(&c->info_empty[slot]));
            if (warp == 4) TC_TRACE(1, it, 2);
        }
    } else if (warp == 8) {
        // ================= tile scheduler + B loader (whole warp in lock-step, one elected lane issues the copies) =================
        const uint32_t leader = elect_one();
        const size_t tab_bytes = (size_t)a.K * a.NC;
        auto publish = [&](int it) -> int {  // lane 0: claim the next tile and publish it in slot it % 4; returns its table, -1 = no more
            const int slot = it & (TC_INFO_SLOTS - 1);
            mbar_wait_spin(smem_u32(&c->info_empty[slot]), ((it / TC_INFO_SLOTS) & 1) ^ 1);
            TileInfo ti{};
            ti.nf = 0;
            ti.tab = -1;
            for (;;) {
                const int t = atomicAdd(a.counter, 1);
                if (t >= a.total_tiles) break;
                const int dev = t / a.tiles_per_dev, st = t - dev * a.tiles_per_dev;
                const K1Dev dv = a.devs[dev];
                const int f0 = st * 128;
                if (f0 >= dv.n_frames || dv.n_channels <= 0) continue;
                ti.src = dv.raw + dv.start_byte + (unsigned long long)f0 * dv.hop_bytes;
                ti.nf = min(128, dv.n_frames - f0);
                ti.pos = dv.pos0 + f0;
                ti.gbase = dv.g0;
                ti.nch = dv.n_channels;
                ti.tab = a.tab_of_dev[dev];
                break;
            }
            c->info[slot] = ti;
            mbar_arrive(smem_u32(&c->info_full[slot]));
            return ti.tab;
        };
        uint32_t stg = 0, ph = 0;
        const uint32_t fb_bar0 = smem_u32(&c->full_b[0]), eb_bar0 = smem_u32(&c->empty_b[0]);
        uint32_t fb_bar = fb_bar0, eb_bar = eb_bar0, dst = smem_u32(bring);
        int tab = lane == 0 ? publish(0) : 0;
        tab = __shfl_sync(0xffffffffu, tab, 0);
        for (int it = 0; tab >= 0; ++it) {
            int next_tab = lane == 0 ? publish(it + 1) : 0;  // the A producers start on tile it+1 while tile it's coefficients stream
            next_tab = __shfl_sync(0xffffffffu, next_tab, 0);
            const signed char* src = a.btab + (size_t)tab * tab_bytes;
            TC_TRACE(2, it, 0);
            int kb = kb_rot;
            for (int kbi = 0; kbi < NKB; kbi++) {
                mbar_wait_spin(eb_bar, ph ^ 1);
                TC_TRACE_STAGE(2, it, kbi, 0);
                if (leader) {
                    if (a.dbg_skip & 1) {
                        mbar_arrive(fb_bar);
                    } else {
                        mbar_arrive_expect_tx(fb_bar, (uint32_t)stage_bytes);
                        bulk_g2s(dst, src + (size_t)kb * stage_bytes, (uint32_t)stage_bytes, fb_bar);
                    }
                }
                TC_TRACE_STAGE(2, it, kbi, 1);
                if (++kb == NKB) kb = 0;
                dst += (uint32_t)stage_bytes;
                fb_bar += 8;
                eb_bar += 8;
                if (++stg == (uint32_t)a.NSTB) stg = 0, ph ^= 1, dst = smem_u32(bring), fb_bar = fb_bar0, eb_bar = eb_bar0;
            }
            TC_TRACE(2, it, 1);
            tab = next_tab;
        }
    } else {
        // ================= MMA issuer (whole warp in lock-step, one elected lane issues) =================
        const uint32_t leader = (a.dbg_skip & 4) ? 0u : elect_one();
        const uint32_t committer = elect_one();
        const uint32_t idesc = idesc_i8(128, a.NC, a.a_signed, 1);
        const int KSTEPS = a.K / 32;
        uint32_t stg = 0, ph = 0;
        // descriptor words that never change: LBO in bits [16,30) of the low word, SBO + version in the high word
        const uint64_t adesc0 = smem_desc_noswizzle(0, (uint32_t)a.S, 128u), bdesc0 = smem_desc_noswizzle(0, (uint32_t)a.NC * 16u, 128u);
        const uint32_t a_hi = (uint32_t)(adesc0 >> 32), b_hi = (uint32_t)(bdesc0 >> 32);
        const uint32_t bstep16 = (uint32_t)(a.NC * 32) >> 4, stage16 = (uint32_t)stage_bytes >> 4;
        const uint32_t b_lo_ring = (uint32_t)bdesc0 + (smem_u32(bring) >> 4);
        uint32_t b_lo = b_lo_ring;
        const uint32_t fb_bar0 = smem_u32(&c->full_b[0]), eb_bar0 = smem_u32(&c->empty_b[0]);
        uint32_t fb_bar = fb_bar0, eb_bar = eb_bar0;
        uint32_t probe = 0;  // result of the early try_wait on the stage about to be consumed
        const uint32_t a_step = 2u * ((uint32_t)a.S >> 4);
        const int spr_ks = a.HC / 2;       // k-steps per hop-row
        const int spr = spr_ks / KBS;      // stages per hop-row (exact when a.stage_in_row)
        // k-step number kbi * KBS + ks of the tile goes to partial accumulator (kbi * KBS + ks) % NACC; the first NACC k-steps
        // overwrite (accumulate = 0)
        auto acc_col = [](int kbi, int ks) -> uint32_t { return (uint32_t)(((kbi * KBS + ks) & (NACC - 1)) * NCS); };
        auto acc_first = [](int kbi, int ks) -> bool { return kbi * KBS + ks < NACC; };
        for (int it = 0;; ++it) {
            const int slot = it & (TC_INFO_SLOTS - 1);
            mbar_wait_spin(smem_u32(&c->info_full[slot]), (it / TC_INFO_SLOTS) & 1);
            if (c->info[slot].nf <= 0) break;
            const int buf = it & 1;
            TC_TRACE(3, it, 0);
            mbar_wait_spin(smem_u32(&c->full_a[buf]), (it >> 1) & 1);
            TC_TRACE(3, it, 1);
            mbar_wait_spin(smem_u32(&c->tmem_empty[buf]), ((it >> 1) & 1) ^ 1);
            TC_TRACE(3, it, 2);
            fence_proxy_async();
            tc_fence_after();
            const uint32_t d_tmem = tmem + (uint32_t)(buf * ACC_STRIDE);
            const uint32_t a_lo = (uint32_t)adesc0 + (smem_u32(abuf + buf * abuf_bytes) >> 4);
            // Stage `kb` of the frame covers k-steps [kb*KBS, kb*KBS + KBS).  When every stage lies inside one hop-row
            // (a.stage_in_row) its A start offsets are base + ks * a_step with base following a two-level counter (columns
            // inside the row, then the next row): pure uniform arithmetic, nothing loaded per stage.
            int kb = kb_rot;
            uint32_t base = a_lo + a.aoff[kb_rot * KBS];
            int in_row = (kb_rot * KBS) % spr_ks / KBS;  // stage index inside its hop-row
            for (int kbi = 0; kbi < NKB; kbi++) {
                // (the barrier of this stage was probed one iteration ago: its round trip overlaps the previous stage's issue)
                if (!probe) mbar_wait_spin(fb_bar, ph);
                {
                    uint32_t nb = fb_bar + 8, nph = ph;
                    if (stg + 1 == (uint32_t)a.NSTB) nb = fb_bar0, nph = ph ^ 1;
                    probe = mbar_test(nb, nph) ? 1u : 0u;  // may be a stage of the NEXT tile: same ring, same order
                }
                TC_TRACE_STAGE(3, it, kbi, 0);
                tc_fence_after();
                if (a.stage_in_row) {
#pragma unroll
                    for (int ks = 0; ks < KBS; ks++)
                        mma_i8_split_if(leader, d_tmem + acc_col(kbi, ks), base + ks * a_step, a_hi, b_lo + ks * bstep16, b_hi, idesc, acc_first(kbi, ks) ? 0u : 1u);
                } else {
#pragma unroll
                    for (int ks = 0; ks < KBS; ks++)
                        mma_i8_split_if(leader, d_tmem + acc_col(kbi, ks), a_lo + a.aoff[kb * KBS + ks], a_hi, b_lo + ks * bstep16, b_hi, idesc,
                                        acc_first(kbi, ks) ? 0u : 1u);
                }
                mma_commit_if(committer, eb_bar);
                TC_TRACE_STAGE(3, it, kbi, 1);
                // next stage of the frame (wrapping around to its start)
                if (++kb == NKB) {
                    kb = 0;
                    base = a_lo + a.aoff[0];
                    in_row = 0;
                } else if (++in_row == spr) {
                    in_row = 0;
                    base = base - (uint32_t)(spr - 1) * KBS * a_step + 1u;  // first column of the next hop-row (rows are 16 bytes apart)
                } else {
                    base += KBS * a_step;
                }
                b_lo += stage16;
                fb_bar += 8;
                eb_bar += 8;
                if (++stg == (uint32_t)a.NSTB) stg = 0, ph ^= 1, b_lo = b_lo_ring, fb_bar = fb_bar0, eb_bar = eb_bar0;
            }
            mma_commit_if(committer, smem_u32(&c->empty_a[buf]));
            mma_commit_if(committer, smem_u32(&c->tmem_full[buf]));
            TC_TRACE(3, it, 3);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc<TMEM_COLS>(tmem);
}

template <int NACC, int KBS>
cudaError_t tc_launch_one(const TcArgs& args, int grid, size_t smem, cudaStream_t s) {
    auto kern = k1_tc_kernel<NACC, KBS>;
    static AbgPerDeviceSize configured;  // per instantiation, per CUDA device
    cudaError_t e = configured.ensure(smem, [&]() {
        cudaError_t e2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e2 != cudaSuccess) return e2;
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        return cudaSuccess;
    });
    if (e != cudaSuccess) return e;
    kern<<<grid, TC_THREADS, smem, s>>>(args);
    return cudaGetLastError();
}
template <int NACC>
cudaError_t tc_launch_acc(const TcArgs& args, int grid, size_t smem, cudaStream_t s) {
    switch (args.KBS) {
        case 1: return tc_launch_one<NACC, 1>(args, grid, smem, s);
        case 2: return tc_launch_one<NACC, 2>(args, grid, smem, s);
        case 4: return tc_launch_one<NACC, 4>(args, grid, smem, s);
        case 8: return tc_launch_one<NACC, 8>(args, grid, smem, s);
    }
    return cudaErrorInvalidValue;
}

long long* g_tc_trace = nullptr;
}  // namespace

// measurement aid: copy the clock64 stamps of the last traced launches (ABG_K1_TC_TRACE set) to the host; 256*4*16*4 values
int abg_k1tc_trace_dump(long long* out) {
    if (!g_tc_trace) return -1;
    cudaDeviceSynchronize();
    return cudaMemcpy(out, g_tc_trace, sizeof(long long) * 256 * 4 * 16 * 4, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -1;
}

// Geometry of the tensor-core K1 for one launch group, or eligible == 0 when the group has to use the CUDA-core kernels.
int abg_k1tc_plan(int fft_size, int sfmt, int hop_bytes, int max_channels, int digits, K1TcPlan* p) {
    *p = K1TcPlan{};
    if (sfmt != ABG_SFMT_U8 && sfmt != ABG_SFMT_S8) return 0;  // 16/32-bit formats stay on the FP32 kernels
    if (hop_bytes % 32 != 0 || hop_bytes < 32) return 0;        // both 16-byte chunks of an MMA's K slice must lie in one hop-row
    if (digits != 3 && digits != 4) return 0;
    const int K = fft_size * 2;
    const int HC = hop_bytes / 16;
    const int halo = (K - 32) / hop_bytes;                      // extra hop-rows the last frame of a tile reaches into
    int rows = 128 + halo;
    if ((rows & 1) == 0) rows++;                                // odd 16-byte pitch: conflict-free transposed writes
    const int S = rows * 16;
    const int C2p = (2 * max_channels + 7) & ~7;
    const int NC = (digits * C2p + 15) & ~15;
    if (NC > 256) return 0;
    int stage_target = 8192;
    if (const char* ev = getenv("ABG_K1_TC_STAGE_BYTES")) stage_target = std::max(1024, atoi(ev));
    if (K > 16384) return 0;
    int KBS = 1;
    while (KBS < 8 && KBS * 2 * NC * 32 <= stage_target) KBS <<= 1;  // power of two: divides K/32
    const int stage = KBS * NC * 32;
    const size_t base = TC_CTRL_BYTES + (((size_t)2 * HC * S + 127) & ~(size_t)127);
    size_t cap = 200 * 1024;                                     // leaves room for K2's one-warp CTAs on the same SM (measured: best pipelined step)
    if (const char* ev = getenv("ABG_K1_TC_CAP_KB")) cap = std::min<size_t>(227, std::max(64, atoi(ev))) * 1024;
    int max_stages = TC_MAX_BSTAGES;
    if (const char* ev = getenv("ABG_K1_TC_STAGES")) max_stages = std::min(TC_MAX_BSTAGES, std::max(2, atoi(ev)));
    const size_t hard_cap = 227 * 1024;
    int nstb = (int)std::min<size_t>(max_stages, base < cap ? (cap - base) / stage : 0);
    if (nstb < 2) nstb = (int)std::min<size_t>(max_stages, base < hard_cap ? (hard_cap - base) / stage : 0);
    if (nstb < 2) return 0;
    int cols = 32;
    while (cols < NC) cols <<= 1;
    p->eligible = 1; p->K = K; p->HC = HC; p->S = S; p->NC = NC; p->ND = digits; p->C2p = C2p; p->KBS = KBS; p->NSTB = nstb;
    int nacc = 1;  // partial accumulators per tile (2 tiles x nacc x cols <= 512 TMEM columns): measured no gain over 1 on B200 (MMAs into
                   // one accumulator already run back to back, tools/tc_probe2.cu), kept selectable with ABG_K1_TC_NACC
    if (getenv("ABG_K1_TC_NACC")) nacc = 4;
    while (nacc > 1 && 2 * nacc * cols > 512) nacc >>= 1;
    if (const char* ev = getenv("ABG_K1_TC_NACC")) nacc = std::max(1, std::min(nacc, atoi(ev) >= 4 ? 4 : (atoi(ev) >= 2 ? 2 : 1)));
    p->nacc = nacc;
    p->tmem_cols = 512; p->smem_bytes = (int)(base + (size_t)nstb * stage); p->halo = halo;
    p->table_bytes = (size_t)K * NC;
    return 1;
}

// Coefficient table of one device for the plan: tab[K * NC] signed digits in the shared-memory image the MMA reads
// ([k-step][16-byte chunk][column][16 bytes]), sq[C2p] = sum over K of the quantised coefficient per output, and the scale
// that turns the recombined integer into the reference's float (returned through *cscale; identical for every table of
// the group).  wsc[n] = window[n] * 1/full-scale (reference src/rtl_airband.cpp:319-324,335-351).
void abg_k1tc_build_table(const K1TcPlan& p, int fft_size, int sfmt, const float* wsc, const int32_t* bins, int n_channels, signed char* tab,
                          long long* sq, double* cscale) {
    const int N = fft_size, K = p.K, NC = p.NC, ND = p.ND;
    double gmax = 0.0;
    for (int n = 0; n < N; n++) gmax = std::max(gmax, fabs((double)wsc[n]));
    if (gmax <= 0.0) gmax = 1.0;
    const double Q = ldexp(1.0, 8 * ND - 2);
    std::fill(tab, tab + (size_t)K * NC, (signed char)0);
    for (int o = 0; o < p.C2p; o++) sq[o] = 0;
    for (int c = 0; c < n_channels; c++) {
        const int b = bins[c] & (N - 1);
        for (int n = 0; n < N; n++) {
            const double th = 2.0 * M_PI * (double)(((long long)n * b) % N) / (double)N;
            const double w = (double)wsc[n], cs = cos(th) * w, sn = sin(th) * w;
            // (I + iQ) * w * (cos - i sin):  Re = I*cs + Q*sn,  Im = -I*sn + Q*cs
            const double coef[2][2] = {{cs, sn}, {-sn, cs}};  // [re/im output][I/Q input]
            for (int reim = 0; reim < 2; reim++)
                for (int comp = 0; comp < 2; comp++) {
                    const int o = 2 * c + reim;
                    const int kk = 2 * n + comp;
                    long long q = llround(coef[reim][comp] / gmax * Q);
                    sq[o] += q;
                    const int ks = kk >> 5, h = (kk >> 4) & 1, t = kk & 15;
                    for (int d = ND - 1; d >= 0; d--) {  // balanced base-256 digits, least significant first
                        long long dig = ((q + 128) & 255) - 128;
                        q = (q - dig) >> 8;
                        tab[((size_t)(ks * 2 + h) * NC + (size_t)(d * p.C2p + o)) * 16 + t] = (signed char)dig;
                    }
                }
        }
    }
    // U8: X = (gmax / 2Q) * (2v - 255 * sq);  S8: X = (gmax / Q) * v
    *cscale = (sfmt == ABG_SFMT_U8) ? gmax / (2.0 * Q) : gmax / Q;
}

cudaError_t abg_launch_k1_tc(const K1Launch& L, const K1TcPlan& p, const K1TcTables& T, int sm_count, cudaStream_t s) {
    TcArgs a{};
    a.devs = L.devs; a.tab_of_dev = T.tab_of_dev; a.btab = T.btab; a.sq = T.sq; a.win = L.win; a.iqin = L.iqin; a.counter = T.counter;
    a.status = T.status; a.cscale = T.cscale; a.Gp = L.Gp; a.n_devices = L.n_devices;
    a.tiles_per_dev = (L.max_frames + 127) / 128;
    a.total_tiles = a.tiles_per_dev * L.n_devices;
    a.K = p.K; a.hop_bytes = p.HC * 16; a.HC = p.HC; a.S = p.S; a.NC = p.NC; a.ND = p.ND; a.C2p = p.C2p; a.KBS = p.KBS; a.NSTB = p.NSTB;
    a.a_signed = (L.sfmt == ABG_SFMT_S8) ? 1 : 0;
    a.mul = a.a_signed ? 1 : 2;
    a.off = a.a_signed ? 0 : 255;
    a.rotate = 1;
    a.dbg_skip = 0;
    if (const char* ev = getenv("ABG_K1_TC_SKIP")) a.dbg_skip = atoi(ev);
    a.trace = nullptr;
    static long long* trace_buf = nullptr;  // measurement aid only: one process-wide buffer, dumped by abg_debug_k1tc_trace_dump()
    if (getenv("ABG_K1_TC_TRACE")) {
        if (!trace_buf) {
            cudaMalloc((void**)&trace_buf, sizeof(long long) * 256 * 4 * 16 * 4);
            cudaMemset(trace_buf, 0, sizeof(long long) * 256 * 4 * 16 * 4);
        }
        a.trace = trace_buf;
        g_tc_trace = trace_buf;
    }
    a.stage_in_row = ((p.HC / 2) % p.KBS == 0) ? 1 : 0;
    for (int ks = 0; ks < p.K / 32; ks++) {
        // K bytes [32ks, 32ks+32) of a frame are the two 16-byte columns j, j+1 of hop-row q
        const int k0 = ks * 32, q = k0 / a.hop_bytes, j = (k0 - q * a.hop_bytes) >> 4;
        a.aoff[ks] = (uint16_t)(j * (p.S >> 4) + q);
    }
    if (const char* ev = getenv("ABG_K1_TC_ROTATE")) a.rotate = atoi(ev) != 0;
    if (a.total_tiles <= 0) return cudaSuccess;
    const int grid = std::min(a.total_tiles, std::max(sm_count, 1));
    switch (p.nacc) {
        case 1: return tc_launch_acc<1>(a, grid, (size_t)p.smem_bytes, s);
        case 2: return tc_launch_acc<2>(a, grid, (size_t)p.smem_bytes, s);
        case 4: return tc_launch_acc<4>(a, grid, (size_t)p.smem_bytes, s);
    }
    return cudaErrorInvalidValue;
}
