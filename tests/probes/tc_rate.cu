// Probe: tcgen05.mma issue rate (cycles per M128 x N x K16 bf16 MMA, cta_group::1) for shared-memory operand
// layouts: no-swizzle ("interleave") vs 128-byte swizzle, A K-major, B MN-major. Timing only (operands are zeros).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_rate tc_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}

__global__ void __launch_bounds__(128) rate_kernel(int n, int a_swz, int b_swz, int reps, long long* cycles, int a_tmem_mode, int same_d, int a_mn = 0) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    unsigned char* sa = smem;               // 16 KB: 128 rows x 64 k
    unsigned char* sb = smem + 16384;       // up to 64 k x 256 n x 2 B = 32 KB (+ pad)
    for (int i = tid; i < (16384 + 40960) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((a_mn ? 1u : 0u) << 15) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const int ks = r & 3;
            uint64_t da, db;
            if (a_mn == 1) da = make_desc(smem_u32(sa) + ks * 2 * 128, 128, 4 * 128 + 16, 0);   // MN-major no swizzle (64 k)
            else if (a_mn == 2) da = make_desc(smem_u32(sa) + ks * 2 * 1024, 8192, 1024, 2);   // MN-major SW128: 64-wide m atoms
            else if (a_swz) da = make_desc(smem_u32(sa) + ks * 32, 16, 1024, 2);            // SW128 K-major: +32 B per k-step
            else da = make_desc(smem_u32(sa) + ks * 2 * 2048, 2048, 128, 0);          // no swizzle
            if (b_swz) db = make_desc(smem_u32(sb) + ks * 2 * 1024, 8192, 1024, 2);   // SW128 MN-major: 64-wide n atoms at LBO, 8 k-rows at SBO
            else db = make_desc(smem_u32(sb) + ks * 2 * 128, 128, 8 * 128 + 16, 0);   // no swizzle
            const uint32_t dcol = same_d ? 0u : (uint32_t)((r & 1) * 64);   // alternate accumulators (only valid for n <= 64)
            if (a_tmem_mode) {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                             ::"r"(tmem + dcol), "r"(tmem + 192 + ks * 8), "l"(db), "r"(idesc), "r"(1u) : "memory");
            } else {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem + dcol), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        for (long it = 0; it < 200000000 && !done; ++it)
            asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        const long long t1 = clock64();
        cycles[blockIdx.x] = done ? (t1 - t0) : -1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

int main() {
    long long* d; cudaMalloc(&d, 148 * 8);
    const int smem = 16384 + 40960 + 1024;
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int reps = 2048;
    const int ns[4] = {64, 128, 192, 256};
    for (int grid : {148})
        for (int cfg = 0; cfg < 4; ++cfg)
            for (int ni = 0; ni < 4; ++ni) {
                const int a_tm = cfg & 1, same_d = (cfg >> 1) ? 0 : 1;
                if (!same_d && ns[ni] > 64) continue;
                rate_kernel<<<grid, 128, smem>>>(ns[ni], 0, 0, reps, d, a_tm, same_d);
                cudaError_t e = cudaDeviceSynchronize();
                long long h[148]; cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
                long long mx = 0; for (int i = 0; i < grid; ++i) if (h[i] > mx) mx = h[i];
                printf("grid=%3d A_from_tmem=%d same_accumulator=%d N=%3d : %s  %.1f cycles/MMA (floor %d)\n", grid, a_tm, same_d, ns[ni],
                       cudaGetErrorString(e), (double)mx / reps, 128 * ns[ni] / 256);
                if (e != cudaSuccess) return 1;
            }
    // A operand MN-major from shared memory (weight-gradient kernel): no swizzle / 128-byte swizzle
    for (int a_mn = 1; a_mn <= 2; ++a_mn)
        for (int b_swz = 0; b_swz < 2; ++b_swz)
            for (int ni = 0; ni < 3; ++ni) {
                rate_kernel<<<148, 128, smem>>>(ns[ni], 0, b_swz, reps, d, 0, 1, a_mn);
                cudaError_t e = cudaDeviceSynchronize();
                long long h[148]; cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
                long long mx = 0; for (int i = 0; i < 148; ++i) if (h[i] > mx) mx = h[i];
                printf("A MN-major %s, B %s, N=%3d : %s  %.1f cycles/MMA (floor %d)\n", a_mn == 1 ? "no-swizzle" : "SW128", b_swz ? "SW128" : "no-swizzle",
                       ns[ni], cudaGetErrorString(e), (double)mx / reps, 128 * ns[ni] / 256);
                if (e != cudaSuccess) return 1;
            }
    return 0;
}
