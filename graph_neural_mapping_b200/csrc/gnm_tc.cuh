// tcgen05 / TMEM / mbarrier helpers shared by the sm_100a tensor-core kernels (inline PTX).
// Conventions verified on hardware by tests/probes/tc_probe.cu:
//   * shared-memory matrix descriptor, no swizzle: start>>4 | (LBO>>4)<<16 | (SBO>>4)<<32 | 1<<46;
//     K-major operand: core matrix = 8 rows x 16 B, LBO = byte stride between k-cores, SBO = between 8-row groups;
//     MN-major operand: core matrix = 8 k-rows x 16 B, LBO = between k-cores, SBO = between 8-column groups;
//   * A operand may come from tensor memory (lane = row, 32-bit column j = elements k=2j (low half), 2j+1);
//   * accumulator tile M=128: row r = TMEM lane r; a warp reads lanes 32*(warp%4).. with tcgen05.ld.32x32b.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr long long TC_TIMEOUT_CYCLES = 4000000000LL;
__device__ int g_tc_abort = 0;     // raised when a bounded barrier wait times out (per translation unit)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait: false on timeout (the caller raises the abort flag and drains). try_wait returns after a short,
// implementation-defined time, so a waiting warp polls; the polls of a whole warp cost issue slots that the
// conversion warps need (ncu: 10-17% of all executed instructions in the row-GEMM kernels), hence BACKOFF_NS > 0
// parks the warp with nanosleep between polls. The single MMA-issuing thread polls without back-off (its wake-up
// latency is on the critical path). The clock / abort flag are only consulted every 64 polls.
// GNM_MBAR_HINT_NS > 0: every poll is a try_wait WITH a suspend-time hint - the thread sleeps in hardware until the phase
// completes or the hint expires and costs no issue slots meanwhile (without the hint try_wait came back after ~20 cycles:
// 40 % of all warp instructions of the fused aggregation were polls, and the MMA thread's polls - no back-off - competed
// with the epilogue / producer warps of its scheduler, profiles/r2_aggregate_tc_fused_before_lines.txt).
#ifndef GNM_MBAR_HINT_NS
#define GNM_MBAR_HINT_NS 2000
#endif
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t addr, uint32_t parity) {
    uint32_t done = 0;
#if GNM_MBAR_HINT_NS > 0
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, P;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity), "r"((uint32_t)GNM_MBAR_HINT_NS) : "memory");
#else
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
#endif
    return done;
}
template <int BACKOFF_NS = 0>
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, volatile int* abort_flag,
                                          long long* waited = nullptr) {
    const uint32_t addr = smem_u32(bar);
    if (mbar_try_wait(addr, parity)) return true;
    const long long t0 = clock64();
    int spins = 0;
    while (true) {
        if (GNM_MBAR_HINT_NS == 0 && BACKOFF_NS > 0) __nanosleep(BACKOFF_NS);
        if (mbar_try_wait(addr, parity)) {
            if (waited) *waited += clock64() - t0;
            return true;
        }
        if ((++spins & 63) == 0) {
            if (*abort_flag) return false;
            if (clock64() - t0 > TC_TIMEOUT_CYCLES) {
                *abort_flag = 1;
                g_tc_abort = 1;
                return false;
            }
        }
    }
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor, no swizzle (layout_type 0), version 1
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}

// D[tmem] (+)= A[tmem] . B[smem descriptor]   (A operand from tensor memory)
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}

__device__ __forceinline__ uint32_t bits2_bf16x2(uint32_t x) { return (x & 1u) * 0x3F80u + (x & 2u) * 0x1FC00000u; }

__device__ __forceinline__ void split3_tc(float x, float& hi, float& mid, float& lo) {
    hi = __bfloat162float(__float2bfloat16_rn(x));
    const float r1 = x - hi;
    mid = __bfloat162float(__float2bfloat16_rn(r1));
    lo = r1 - mid;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&v);
}
// exact fp32 -> (hi, mid, lo) bf16 split of two values at once (packed cvt.rn.bf16x2; bf16 -> fp32 is a shift / mask)
__device__ __forceinline__ void split3x2(float x0, float x1, uint32_t& hi2, uint32_t& mid2, uint32_t& lo2) {
    hi2 = pack2(x0, x1);
    float r0 = x0 - __uint_as_float(hi2 << 16), r1 = x1 - __uint_as_float(hi2 & 0xffff0000u);
    mid2 = pack2(r0, r1);
    r0 -= __uint_as_float(mid2 << 16);
    r1 -= __uint_as_float(mid2 & 0xffff0000u);
    lo2 = pack2(r0, r1);
}


}  // namespace
