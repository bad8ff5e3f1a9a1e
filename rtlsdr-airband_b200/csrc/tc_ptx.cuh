// Inline-PTX building blocks of the tensor-core K1 (k1_tc.cu) for sm_100a: mbarrier, cp.async / bulk copies,
// tcgen05 (TMEM allocation, UMMA shared-memory / instruction descriptors, kind::i8 MMA, commit, TMEM loads).
//
// Descriptor encodings follow the PTX ISA "tcgen05 matrix descriptor" / "instruction descriptor" tables:
//   shared-memory descriptor (64 bit): [0,14) start address >> 4 | [16,30) leading byte offset >> 4 |
//       [32,46) stride byte offset >> 4 | [46,48) version = 1 | [49,52) base offset | [61,64) swizzle (0 = none)
//   K-major operand without swizzle ("interleaved" canonical layout): 8 rows x 16 bytes form one 128-byte core
//   matrix; LBO = byte distance between the two 16-byte K chunks of one MMA (K = 32 bytes for 8-bit types),
//   SBO = byte distance between consecutive 8-row groups along M (or N).
//   instruction descriptor (32 bit): [4,6) D format (2 = S32) | [7,10) A format (0 = U8, 1 = S8) | [10,13) B format |
//       [15] A major (0 = K) | [16] B major (0 = K) | [17,23) N >> 3 | [24,29) M >> 4
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one lane of a fully converged warp (the others get 0)
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred)
        :
        : "memory");
    return pred;
}

// ---- mbarrier -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
// non-blocking probe of a phase (test_wait never suspends the warp)
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
// Bounded wait: a barrier that never completes (a protocol bug) must not hang the GPU.  Returns false on time-out.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, long long budget_cycles = (1ll << 31)) {
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > budget_cycles) return false;
    }
    return true;
}

// Spin inside ONE asm statement (so the compiler sees straight-line, warp-uniform code around it and keeps the role loops
// on the uniform datapath), bounded: a barrier that never completes is a protocol bug and must not hang the GPU, so after
// ~10^6 failed try_waits (each suspends up to the hardware's time limit) the kernel traps and the launch fails cleanly.
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .u32 n;\n\t"
        "mov.u32 n, 0;\n"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "add.u32 n, n, 1;\n\t"
        "setp.lt.u32 p, n, 0x100000;\n\t"
        "@p bra WAIT_LOOP;\n\t"
        "trap;\n"
        "WAIT_DONE:\n\t}"
        :
        : "r"(bar), "r"(parity)
        : "memory");
}

// Same contract, polling with the non-blocking test_wait: lower wake-up latency than the suspending try_wait, at the price
// of issue slots (use on the single-warp critical path of a pipeline only).
__device__ __forceinline__ void mbar_wait_poll(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .u32 n;\n\t"
        "mov.u32 n, 0;\n"
        "POLL_LOOP:\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra POLL_DONE;\n\t"
        "add.u32 n, n, 1;\n\t"
        "setp.lt.u32 p, n, 0x4000000;\n\t"
        "@p bra POLL_LOOP;\n\t"
        "trap;\n"
        "POLL_DONE:\n\t}"
        :
        : "r"(bar), "r"(parity)
        : "memory");
}

// ---- copies ---------------------------------------------------------------------------------------------------------
// 16-byte asynchronous copy global -> shared (LDGSTS), L2 only
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// generic-proxy writes (st.shared, cp.async) -> visible to the async proxy (tcgen05.mma operand reads, bulk copies)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
}

// ---- tcgen05 --------------------------------------------------------------------------------------------------------
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem) {  // whole warp; NCOLS a power of two in [32, 512]
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // whole warp (the one that allocated)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__host__ __device__ constexpr uint64_t smem_desc_noswizzle(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) |
           (1ull << 46);
}
__host__ __device__ constexpr uint32_t idesc_i8(int m, int n, int a_signed, int b_signed) {
    return (2u << 4) | ((uint32_t)(a_signed ? 1 : 0) << 7) | ((uint32_t)(b_signed ? 1 : 0) << 10) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T, 8-bit integer operands, S32 accumulator; issued by ONE thread
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// same, descriptors passed as 32-bit halves (the low word carries the start address and is the only part that changes
// from one K step to the next: no 64-bit arithmetic in the issue loop)
__device__ __forceinline__ void mma_i8_split(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %5, p;\n\t}"
        :
        : "r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// predicated forms (issue != 0 on exactly one lane of the warp): no divergent branch around the instruction, so the loop that
// computes the operands stays warp-uniform for the compiler
__device__ __forceinline__ void mma_i8_split_if(uint32_t issue, uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                                uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.ne.b32 q, %7, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %5, p;\n\t}"
        :
        : "r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(issue)
        : "memory");
}
__device__ __forceinline__ void mma_commit_if(uint32_t issue, uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "setp.ne.b32 q, %1, 0;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        :
        : "r"(bar), "r"(issue)
        : "memory");
}
// mbarrier arrive once every MMA issued so far by this thread has completed (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 bit, 8 consecutive columns: thread t of warp w reads TMEM lane 32*(w%4)+t
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace tc
