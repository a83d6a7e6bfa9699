// Standalone probe (not part of the library): checks the tcgen05 shared-memory / instruction descriptor
// conventions used by gnm_aggregate_tc.cu on a single 128 x 192 x 64 bf16 GEMM.
//   A: K-major, no swizzle   (core matrix = 8 rows x 16 B; k-cores at LBO, 8-row groups at SBO)
//   B: MN-major, no swizzle  (core matrix = 8 k-rows x 16 B of n; n-cores at SBO, k-cores at LBO)
// Checks A from shared memory (SS) and A from tensor memory written with tcgen05.st (TS).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tc_probe.cu ; run: ./tc_probe
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

constexpr int M = 128, N = 192, K = 64;
constexpr int A_KCORE_STRIDE = 16 * 128;          // bytes between k-cores of A ([k-core][row-group][128 B])
constexpr int A_RGROUP_STRIDE = 128;              // bytes between 8-row groups of A
constexpr int B_KCORES = K / 8;
constexpr int B_NCORE_STRIDE = B_KCORES * 128 + 16;   // bytes between n-cores of B (padded)
constexpr int B_KCORE_STRIDE = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
    return d;                                      // layout_type = 0 (no swizzle), base_offset = 0
}

__global__ void __launch_bounds__(128) probe_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                                    float* __restrict__ d, int ts_mode, int unused, int* status) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sa = smem;                                   // 16 KB
    unsigned char* sb = smem + 16384;                           // 24 * B_NCORE_STRIDE
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // fill A: element (r, k) -> (k/8)*A_KCORE_STRIDE + (r/8)*128 + (r%8)*16 + (k%8)*2
    for (int i = tid; i < M * K; i += 128) {
        const int r = i / K, k = i % K;
        *reinterpret_cast<__nv_bfloat16*>(sa + (k / 8) * A_KCORE_STRIDE + (r / 8) * A_RGROUP_STRIDE + (r % 8) * 16 + (k % 8) * 2) = a[i];
    }
    // fill B: element (k, n) -> (n/8)*B_NCORE_STRIDE + (k/8)*128 + (k%8)*16 + (n%8)*2
    for (int i = tid; i < K * N; i += 128) {
        const int k = i / N, n = i % N;
        *reinterpret_cast<__nv_bfloat16*>(sb + (n / 8) * B_NCORE_STRIDE + (k / 8) * B_KCORE_STRIDE + (k % 8) * 16 + (n % 8) * 2) = b[i];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    if (ts_mode) {
        // A operand in TMEM: lane = row, 32-bit column j of the A region holds k = 2j (low half) and 2j+1 (high half)
        const int row = tid;                       // 128 threads = 128 rows; warp w owns lanes 32w..32w+31
        uint32_t v[32];
        for (int j = 0; j < 32; ++j) {
            const uint16_t lo = *reinterpret_cast<const uint16_t*>(&a[row * K + 2 * j]);
            const uint16_t hi = *reinterpret_cast<const uint16_t*>(&a[row * K + 2 * j + 1]);
            v[j] = (uint32_t)lo | ((uint32_t)hi << 16);
        }
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 192;
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
            "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
            ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
              "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
              "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
              "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (tid == 0) {
        // instruction descriptor: F32 accum, BF16 x BF16, A K-major, B MN-major, N = 192, M = 128
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint32_t a_addr = smem_u32(sa) + ks * 2 * A_KCORE_STRIDE;
            const uint32_t b_addr = smem_u32(sb) + ks * 2 * B_KCORE_STRIDE;
            const uint64_t da = make_desc(a_addr, A_KCORE_STRIDE, A_RGROUP_STRIDE);
            const uint64_t db = make_desc(b_addr, B_KCORE_STRIDE, B_NCORE_STRIDE);
            const uint32_t acc = ks > 0 ? 1u : 0u;
            if (ts_mode) {
                const uint32_t a_tmem = tmem + 192 + ks * 8;
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                    ::"r"(tmem), "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc) : "memory");
            } else {
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                    ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // bounded wait
    uint32_t done = 0;
    for (long it = 0; it < 20000000 && !done; ++it) {
        asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
    if (!done) { if (tid == 0) *status = 1; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (done) {
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                  "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                  "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                  "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int row = warp * 32 + (tid & 31);
            for (int j = 0; j < 32; ++j) d[row * N + c0 + j] = __uint_as_float(v[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

int main() {
    std::vector<__nv_bfloat16> ha(M * K), hb(K * N);
    std::vector<float> fa(M * K), fb(K * N), ref(M * N, 0.f), out(M * N);
    srand(1);
    for (int i = 0; i < M * K; ++i) { float v = (rand() % 17 - 8) / 8.f; ha[i] = __float2bfloat16(v); fa[i] = __bfloat162float(ha[i]); }
    for (int i = 0; i < K * N; ++i) { float v = (rand() % 33 - 16) / 16.f; hb[i] = __float2bfloat16(v); fb[i] = __bfloat162float(hb[i]); }
    for (int r = 0; r < M; ++r) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += fa[r * K + k] * fb[k * N + n]; ref[r * N + n] = s; }
    __nv_bfloat16 *da, *db; float* dd; int* ds;
    cudaMalloc(&da, M * K * 2); cudaMalloc(&db, K * N * 2); cudaMalloc(&dd, M * N * 4); cudaMalloc(&ds, 4);
    cudaMemcpy(da, ha.data(), M * K * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), K * N * 2, cudaMemcpyHostToDevice);
    const int smem = 16384 + 24 * B_NCORE_STRIDE + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int best = -1;
    for (int combo = 0; combo < 2; ++combo) {
        const int sa = combo, sb = 0;   // sa: 0 = A from shared memory, 1 = A from TMEM
        cudaMemset(dd, 0xff, M * N * 4); cudaMemset(ds, 0, 4);
        probe_kernel<<<1, 128, smem>>>(da, db, dd, sa, sb, ds);
        cudaError_t e = cudaDeviceSynchronize();
        int st = 0; cudaMemcpy(&st, ds, 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(out.data(), dd, M * N * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0; int nan = 0;
        for (int i = 0; i < M * N; ++i) { if (out[i] != out[i]) { nan++; continue; } double d = fabs((double)out[i] - ref[i]); if (d > maxerr) maxerr = d; }
        printf("a_from_tmem=%d (%d) : cuda=%s timeout=%d nan=%d max_abs_err=%.6f  (out[0]=%f ref[0]=%f out[last]=%f ref[last]=%f)\n", sa, sb,
               cudaGetErrorString(e), st, nan, maxerr, out[0], ref[0], out[M * N - 1], ref[M * N - 1]);
        if (e == cudaSuccess && st == 0 && nan == 0 && maxerr < 1e-3) best = combo;
        if (e != cudaSuccess) break;
    }
    printf("TC_PROBE_RESULT best_combo=%d\n", best);
    return 0;
}
