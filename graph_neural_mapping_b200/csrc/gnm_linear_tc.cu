// Linear layer on the 5th-generation tensor cores with fp32-level accuracy (reference: nn.Linear in
// models/mlp.py:25,32-35,48-49, fused with the BatchNorm-apply + ReLU of the previous op and the batch
// statistics of the next one, mlp.py:48 / graphcnn.py:163-166).
//
//   y[m, n] = sum_k f(x[m, k]) * W[n, k] + bias[n],   f(x) = relu(x*in_scale[k] + in_shift[k]) or x,  K, N <= 64
//
// fp32 operands are split exactly into three bf16 planes (hi + mid + lo, 8+8+8 significand bits). The products
// kept are hi*hi, hi*mid, mid*hi, hi*lo, lo*hi, mid*mid - everything down to 2^-24 relative - six bf16 MMAs per
// 16-wide k-step, all accumulating into the same 64 fp32 TMEM columns (the tensor pipe has ~20x headroom over the
// HBM time of this op, so the 6x MMA count is free).
//
// One persistent CTA per SM, work item = 128-row tile:
//   * producers (8 warps = two groups taking tiles alternately): a warp's 32 rows arrive by cp.async one own-tile ahead in
//     a padded staging tile (coalesced), are read back one row per lane, f() applied, split, and the three planes written
//     to TENSOR MEMORY (tcgen05.st; 96 columns per stage, 4-stage ring) - the MMA's A operand costs no shared-memory
//     bandwidth;
//   * B operand: the three W planes ([n][k], K-major core matrices) converted once per CTA into 24 KB of shared memory;
//   * MMA thread: 24 x tcgen05.mma (M=128, N=64, K=16, A from TMEM) per tile, tcgen05.commit to free the stage and
//     publish the accumulator (two 64-column slots);
//   * epilogue (8 warps: one group of four per accumulator slot): tcgen05.ld, + bias, staged through shared memory and
//     copied out with coalesced 128-bit stores; the per-column sum / sum of squares ride on that copy-out (each lane
//     owns four columns), are added over the eight warps in a fixed order and flushed with ONE fp64 atomic per column
//     and CTA (same-address atomics serialise in L2).
#include "gnm_common.cuh"
#include "gnm_tc.cuh"
#include "gnm_bn_tail.cuh"

namespace {

constexpr int LT_F = 64;                    // max K and N
constexpr int LT_STAGES = 4;
constexpr int LT_A_COLS = 96;               // TMEM columns per A stage: hi | mid | lo planes, 32 columns each
constexpr int LT_D_COLS = 64;
constexpr int LT_A_TMEM0 = 2 * LT_D_COLS;   // accumulator slots at columns [0,64) and [64,128)
constexpr int LT_EPI_WARPS = 8, LT_PROD_WARPS = 8, LT_GROUPS = LT_PROD_WARPS / 4;
constexpr int LT_THREADS = (LT_EPI_WARPS + LT_PROD_WARPS + 1) * 32;   // 544
constexpr int LT_MMA_WARP = LT_EPI_WARPS + LT_PROD_WARPS;
constexpr int LT_W_PLANE = LT_F * LT_F * 2;          // bytes of one bf16 W plane
constexpr int LT_W_KCORE = 8 * 128;                  // [k-core][n-core][8 n-rows x 16 B]: bytes between k-cores
constexpr int LT_PITCH = LT_F + 4;                  // floats per staged row: 272 B, row-per-lane 128-bit accesses are conflict-optimal
constexpr int LT_STG = 32 * LT_PITCH * 4;            // bytes of one warp's 32-row staging tile
constexpr int LT_SMEM = 3 * LT_W_PLANE + 3 * LT_F * 4 + 256 + (LT_EPI_WARPS + LT_PROD_WARPS) * LT_STG;

struct LinTcParams {
    const float* x; int64_t ldx;
    const float* w; int64_t ldw; int w_is_kn;
    const float* bias; const float* in_scale; const float* in_shift;
    float* y; int64_t ldy;
    double* col_stats;
    BnTailDev tail;              // BatchNorm finalisation of col_stats by the last CTA (kind 0: none)
    int n_rows, n_in, n_out;
};

// ACT: BatchNorm + ReLU prologue on the input rows; FAST: 64 -> 64, 16-byte aligned rows on both sides (the benchmark's
// shape) - the general element-wise staging / copy-out paths are compiled out of that instantiation (the three roles run
// different code at the same time: dead paths cost instruction-cache reach, as measured on the aggregation kernel).
template <bool ACT, bool FAST>
__global__ void __launch_bounds__(LT_THREADS, 1) linear_tc_kernel(const LinTcParams p) {
    extern __shared__ __align__(1024) unsigned char lt_smem[];
    __shared__ __align__(8) uint64_t bars[2 * LT_STAGES + 4];
    __shared__ uint32_t s_tmem;
    __shared__ int s_abort;
    uint64_t* a_full = bars;
    uint64_t* a_empty = bars + LT_STAGES;
    uint64_t* acc_full = bars + 2 * LT_STAGES;
    uint64_t* acc_empty = acc_full + 2;
    unsigned char* sm_w = lt_smem;                                     // 3 planes x 8 KB
    float* sm_sc = reinterpret_cast<float*>(lt_smem + 3 * LT_W_PLANE);  // in_scale[64] | in_shift[64] | bias[64]
    // per-warp staging tiles: global memory is only touched with coalesced 128-bit accesses (a row-per-lane access
    // costs 32 L1 wavefronts per instruction and made the first version L1-bound)
    float* sm_stg = reinterpret_cast<float*>(lt_smem + 3 * LT_W_PLANE + 3 * LT_F * 4 + 256);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    volatile int* abort_flag = &s_abort;
    const int n_tiles = (p.n_rows + 127) >> 7;
    constexpr bool act = ACT;

    // ---- one-time setup: W planes (K-major core matrices), prologue constants, barriers, tensor memory
    pdl_launch_dependents();
    // (every load of a thread is in flight before the first split: the set-up is one global round trip, not four)
    {
        constexpr int NP = (LT_F * LT_F / 2 + LT_THREADS - 1) / LT_THREADS;
        float w0[NP], w1[NP];
#pragma unroll
        for (int u = 0; u < NP; ++u) {
            const int e = tid + u * LT_THREADS;
            const int n = e >> 5, k = (e & 31) * 2;                   // element pair (n, k), (n, k+1)
            w0[u] = w1[u] = 0.f;
            if (e < LT_F * LT_F / 2 && n < p.n_out) {
                if (k < p.n_in) w0[u] = p.w_is_kn ? __ldg(p.w + (int64_t)k * p.ldw + n) : __ldg(p.w + (int64_t)n * p.ldw + k);
                if (k + 1 < p.n_in) w1[u] = p.w_is_kn ? __ldg(p.w + (int64_t)(k + 1) * p.ldw + n) : __ldg(p.w + (int64_t)n * p.ldw + k + 1);
            }
        }
#pragma unroll
        for (int u = 0; u < NP; ++u) {
            const int e = tid + u * LT_THREADS;
            if (e >= LT_F * LT_F / 2) break;
            const int n = e >> 5, k = (e & 31) * 2;
            uint32_t h, m, l;
            split3x2(w0[u], w1[u], h, m, l);
            const int off = (k >> 3) * LT_W_KCORE + (n >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2;
            *reinterpret_cast<uint32_t*>(sm_w + off) = h;
            *reinterpret_cast<uint32_t*>(sm_w + LT_W_PLANE + off) = m;
            *reinterpret_cast<uint32_t*>(sm_w + 2 * LT_W_PLANE + off) = l;
        }
    }
    // weights above, barriers / tensor memory below need nothing from the kernel in front; the BatchNorm affine of the
    // prologue does (it comes out of that kernel's tail): wait here
    pdl_wait();
    for (int k = tid; k < LT_F; k += LT_THREADS) {
        sm_sc[k] = (act && k < p.n_in) ? p.in_scale[k] : 1.f;
        sm_sc[LT_F + k] = (act && k < p.n_in) ? p.in_shift[k] : 0.f;
        sm_sc[2 * LT_F + k] = (p.bias != nullptr && k < p.n_out) ? p.bias[k] : 0.f;
    }
    if (tid == 0) {
        s_abort = 0;
        for (int i = 0; i < LT_STAGES; ++i) { mbar_init(&a_full[i], 4); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == LT_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;

    if (warp < LT_EPI_WARPS) {
        // ================================ epilogue =========================================================
        // A warp executes its instruction stream serially (well under one instruction per cycle), so one group of
        // four warps per accumulator slot drains every other tile: twice the epilogue throughput of a single group.
        float st1[2] = {0.f, 0.f}, st2[2] = {0.f, 0.f};                      // generic-width path: columns lane, lane + 32
        float cs1[4] = {0.f, 0.f, 0.f, 0.f}, cs2[4] = {0.f, 0.f, 0.f, 0.f};  // 64-wide path: columns 4 (lane & 15) ..
        const uint32_t my_slot = warp >> 2;
        const int q = warp & 3;                       // TMEM lane quarter of this warp
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < n_tiles && !*abort_flag; tile += gridDim.x, ++it) {
            const uint32_t slot = it & 1, ph = (it >> 1) & 1;
            if (slot != my_slot) continue;
            if (!mbar_wait<32>(&acc_full[slot], ph, abort_flag)) break;
            tc_fence_after();
            float* stg = sm_stg + warp * (32 * LT_PITCH);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                float v[32];
#pragma unroll
                for (int c0 = 0; c0 < 32; c0 += 16) {
                    uint32_t t16[16];
                    tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + slot * LT_D_COLS + hf * 32 + c0, t16);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[c0 + j] = __uint_as_float(t16[j]);
                }
                if (hf == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[slot]);   // the accumulator is in registers: free the slot
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const int c = hf * 32 + j;
                    if (c >= p.n_out) break;
                    {
                        const float4 b = *reinterpret_cast<const float4*>(sm_sc + 2 * LT_F + c);   // zero beyond n_out
                        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
                    }
                    *reinterpret_cast<float4*>(stg + lane * LT_PITCH + c) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
            }
            // copy the warp's 32 x 64 tile out with coalesced 128-bit stores (two rows per instruction)
            __syncwarp();
            const int row0 = tile * 128 + q * 32;
            if (FAST || (p.n_out == LT_F && (p.ldy & 3) == 0 && gnm_aligned16(p.y))) {
                const int nvalid = p.n_rows - row0 - (lane >> 4);
                char* dstp = reinterpret_cast<char*>(p.y + (int64_t)(row0 + (lane >> 4)) * p.ldy + (lane & 15) * 4);
                const int64_t step = 2 * p.ldy * (int64_t)sizeof(float);
                const float* sp = stg + (lane >> 4) * LT_PITCH + (lane & 15) * 4;
                // the batch statistics ride on the copy-out: this lane owns four columns of every other row.
                // Loads are issued eight at a time ahead of their uses (a per-row branch would serialise them).
#pragma unroll
                for (int i0 = 0; i0 < 16; i0 += 8) {
                    float4 t[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) t[i] = *reinterpret_cast<const float4*>(sp + (i0 + i) * 2 * LT_PITCH);
                    if (nvalid >= 31) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(dstp + (i0 + i) * step) = t[i];
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            if (2 * (i0 + i) < nvalid) *reinterpret_cast<float4*>(dstp + (i0 + i) * step) = t[i];
                            else t[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        cs1[0] += t[i].x; cs1[1] += t[i].y; cs1[2] += t[i].z; cs1[3] += t[i].w;
                        cs2[0] = fmaf(t[i].x, t[i].x, cs2[0]); cs2[1] = fmaf(t[i].y, t[i].y, cs2[1]);
                        cs2[2] = fmaf(t[i].z, t[i].z, cs2[2]); cs2[3] = fmaf(t[i].w, t[i].w, cs2[3]);
                    }
                }
            } else {
                for (int e = lane; e < 32 * LT_F; e += 32) {
                    const int rr = e >> 6, c = e & 63;          // c = lane or lane + 32
                    if (row0 + rr < p.n_rows && c < p.n_out) {
                        const float t = stg[rr * LT_PITCH + c];
                        p.y[(int64_t)(row0 + rr) * p.ldy + c] = t;
                        st1[c >> 5] += t;
                        st2[c >> 5] = fmaf(t, t, st2[c >> 5]);
                    }
                }
            }
            __syncwarp();
        }
        if (p.col_stats != nullptr) {
            // per-warp partial sums -> shared memory -> one fixed-order sum per CTA -> ONE fp64 atomic per column and
            // CTA (same-address atomics serialise in L2: eight times fewer of them is worth ~8 us per launch)
            float* red = sm_stg + warp * (32 * LT_PITCH);          // this warp's staging tile is free now
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                cs1[u] += __shfl_xor_sync(GNM_FULL_MASK, cs1[u], 16);
                cs2[u] += __shfl_xor_sync(GNM_FULL_MASK, cs2[u], 16);
            }
            __syncwarp();
            // generic-width path holds columns lane / lane + 32 in st1, st2; the 64-wide path 4 (lane & 15) + u in cs
            red[lane] = st1[0]; red[32 + lane] = st1[1]; red[64 + lane] = st2[0]; red[96 + lane] = st2[1];
            __syncwarp();
            if (lane < 16) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    red[lane * 4 + u] += cs1[u];
                    red[64 + lane * 4 + u] += cs2[u];
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(LT_EPI_WARPS * 32) : "memory");
            if (tid < 2 * LT_F) {
                float a = 0.f;
#pragma unroll
                for (int w = 0; w < LT_EPI_WARPS; ++w) a += sm_stg[w * (32 * LT_PITCH) + tid];
                const int c = tid & (LT_F - 1);
                if (c < p.n_out) atomicAdd(&p.col_stats[(tid >> 6) * p.n_out + c], (double)a);
            }
        }
    } else if (warp == LT_MMA_WARP) {
        // ================================ MMA issue (one thread) ============================================
        if (lane == 0) {
            // f32 accumulate, bf16 x bf16, B K-major, N = 64, M = 128 (A comes from tensor memory)
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(LT_F >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint64_t dw_hi = umma_desc(smem_u32(sm_w), LT_W_KCORE, 128);
            const uint64_t dw_mid = dw_hi + (uint64_t)(LT_W_PLANE >> 4);
            const uint64_t dw_lo = dw_hi + (uint64_t)(2 * LT_W_PLANE >> 4);
            const int ksteps = (p.n_in + 15) >> 4;
            uint32_t it = 0;
            bool ok = true;
            for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x, ++it) {
                const uint32_t s = it % LT_STAGES, aph = (it / LT_STAGES) & 1;
                const uint32_t slot = it & 1, ph = (it >> 1) & 1;
                if (!(ok = mbar_wait(&acc_empty[slot], ph ^ 1, abort_flag))) break;
                if (!(ok = mbar_wait(&a_full[s], aph, abort_flag))) break;
                tc_fence_after();
                const uint32_t d = tmem + slot * LT_D_COLS;
                const uint32_t a0 = tmem + LT_A_TMEM0 + s * LT_A_COLS;
                for (int ks = 0; ks < ksteps; ++ks) {
                    const uint64_t kofs = (uint64_t)(ks * 2 * LT_W_KCORE >> 4);
                    const uint32_t ah = a0 + ks * 8, am = ah + 32, al = ah + 64;
                    umma_ts(d, ah, dw_hi + kofs, idesc, ks ? 1u : 0u);      // hi * hi
                    umma_ts(d, ah, dw_mid + kofs, idesc, 1u);               // hi * mid
                    umma_ts(d, am, dw_hi + kofs, idesc, 1u);                // mid * hi
                    umma_ts(d, ah, dw_lo + kofs, idesc, 1u);                // hi * lo
                    umma_ts(d, al, dw_hi + kofs, idesc, 1u);                // lo * hi
                    umma_ts(d, am, dw_mid + kofs, idesc, 1u);               // mid * mid
                }
                umma_commit(&a_empty[s]);
                umma_commit(&acc_full[slot]);
            }
        }
    } else {
        // ================================ producers: one row per thread, two warp groups alternate tiles =========
        // four warps (one per TMEM lane quarter) form a group; the groups take tiles alternately. A warp's 32 rows are
        // fetched with cp.async one own-tile ahead into its staging tile, then read back one row per lane.
        const int grp = (warp - LT_EPI_WARPS) >> 2;
        const int q = warp & 3;                                           // TMEM lane quarter = 32-row slice of the tile
        float* stg = sm_stg + warp * (32 * LT_PITCH);
        const uint32_t stg_u32 = smem_u32(stg);
        const bool fast = FAST || (p.n_in == LT_F && (p.ldx & 3) == 0 && gnm_aligned16(p.x));
        // stage the warp's 32 rows of a tile: cp.async (coalesced 16-byte chunks, rows past the end zero-filled)
        auto stage_rows = [&](int tile) {
            const int row0 = tile * 128 + q * 32;
            if (fast) {
                // lane -> (row pair member lane >> 4, float4 column lane & 15); rows advance by two per copy
                const int nvalid = p.n_rows - row0 - (lane >> 4);        // copy i is in range iff 2 i < nvalid
                const char* src = reinterpret_cast<const char*>(p.x + (int64_t)(row0 + (lane >> 4)) * p.ldx + (lane & 15) * 4);
                const int64_t step = 2 * p.ldx * (int64_t)sizeof(float);
                uint32_t dst = stg_u32 + (uint32_t)(((lane >> 4) * LT_PITCH + (lane & 15) * 4) * 4);
                if (nvalid >= 31) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                        src += step;
                        dst += 2 * LT_PITCH * 4;
                    }
                } else {
#pragma unroll 4
                    for (int i = 0; i < 16; ++i) {
                        const bool okr = 2 * i < nvalid;
                        const int nbytes = okr ? 16 : 0;
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(okr ? src : (const char*)p.x), "r"(nbytes) : "memory");
                        src += step;
                        dst += 2 * LT_PITCH * 4;
                    }
                }
            } else {
                for (int e = lane; e < 32 * LT_F; e += 32) {
                    const int rr = e >> 6, k = e & 63;
                    stg[rr * LT_PITCH + k] = (row0 + rr < p.n_rows && k < p.n_in) ? p.x[(int64_t)(row0 + rr) * p.ldx + k] : 0.f;
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        uint32_t it = 0;
        bool ok = true;
        // first own tile
        {
            const int first_tile = blockIdx.x + grp * gridDim.x;
            if (first_tile < n_tiles) stage_rows(first_tile);
        }
        for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x, ++it) {
            if ((it % LT_GROUPS) != (uint32_t)grp) continue;
            const uint32_t s = it % LT_STAGES, aph = (it / LT_STAGES) & 1;
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + LT_A_TMEM0 + s * LT_A_COLS;
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            bool waited = false;
#pragma unroll
            for (int k0 = 0; k0 < LT_F; k0 += 32) {
                uint32_t hi[16], mid[16], lo[16];
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 t = *reinterpret_cast<const float4*>(stg + lane * LT_PITCH + k0 + j);
                    float a0 = t.x, a1 = t.y, a2 = t.z, a3 = t.w;
                    if (act) {
                        const float4 sc = *reinterpret_cast<const float4*>(sm_sc + k0 + j);
                        const float4 sh = *reinterpret_cast<const float4*>(sm_sc + LT_F + k0 + j);
                        a0 = fmaxf(fmaf(a0, sc.x, sh.x), 0.f); a1 = fmaxf(fmaf(a1, sc.y, sh.y), 0.f);
                        a2 = fmaxf(fmaf(a2, sc.z, sh.z), 0.f); a3 = fmaxf(fmaf(a3, sc.w, sh.w), 0.f);
                    }
                    split3x2(a0, a1, hi[j >> 1], mid[j >> 1], lo[j >> 1]);
                    split3x2(a2, a3, hi[(j >> 1) + 1], mid[(j >> 1) + 1], lo[(j >> 1) + 1]);
                }
                if (k0 == 32) {
                    // the staged rows are in registers now: start fetching this warp's next tile into the same buffer
                    __syncwarp();
                    const int next_tile = tile + LT_GROUPS * gridDim.x;
                    if (next_tile < n_tiles) stage_rows(next_tile);
                }
                if (!waited) {
                    if (!(ok = mbar_wait<32>(&a_empty[s], aph ^ 1, abort_flag))) break;
                    waited = true;
                }
                tmem_st16(taddr + (k0 >> 1), hi);
                tmem_st16(taddr + 32 + (k0 >> 1), mid);
                tmem_st16(taddr + 64 + (k0 >> 1), lo);
            }
            if (!ok) break;
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[s]);
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == LT_MMA_WARP) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
    bn_tail_run(p.tail);
}

}  // namespace

// GNM_OK after launching; GNM_ERR_TOO_LARGE when the shape does not fit this kernel (caller uses the FFMA kernel)
int gnm_launch_linear_tc(const float* x, int64_t ldx, int n_rows, int n_in, const float* w, int64_t ldw, int w_is_kn,
                         const float* bias, const float* in_scale, const float* in_shift, float* y, int64_t ldy,
                         int n_out, double* col_stats, const gnm_bn_tail* tail, cudaStream_t stream) {
    if (n_in > LT_F || n_out > LT_F || n_in < 1 || n_out < 1) return GNM_ERR_TOO_LARGE;
    int dev = 0, sms = 148, major = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) return GNM_ERR_TOO_LARGE;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    LinTcParams p;
    p.x = x; p.ldx = ldx; p.w = w; p.ldw = ldw; p.w_is_kn = w_is_kn; p.bias = bias; p.in_scale = in_scale;
    p.in_shift = in_shift; p.y = y; p.ldy = ldy; p.col_stats = col_stats; p.n_rows = n_rows; p.n_in = n_in; p.n_out = n_out;
    const int trc = bn_tail_args(tail, col_stats, n_out, &p.tail);
    if (trc != GNM_OK) return trc;
    const int tiles = (n_rows + 127) / 128;
    const int grid = tiles < sms ? tiles : sms;
    const bool act = in_scale != nullptr;
    const bool fast = n_in == LT_F && n_out == LT_F && (ldx & 3) == 0 && (ldy & 3) == 0 && gnm_aligned16(x) && gnm_aligned16(y);
    cudaError_t e;
#define GNM_LT_LAUNCH(A, F)                                                                                            \
    do {                                                                                                               \
        e = cudaFuncSetAttribute(linear_tc_kernel<A, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, LT_SMEM);        \
        if (e != cudaSuccess) return (int)e;                                                                           \
        gnm_count_launch(GNM_K_LINEAR_TC);                                                                             \
        e = gnm_launch_pdl<LinTcParams>(linear_tc_kernel<A, F>, grid, LT_THREADS, LT_SMEM, stream, p);                  \
        if (e != cudaSuccess) return (int)e;                                                                           \
    } while (0)
    if (act && fast) GNM_LT_LAUNCH(true, true);
    else if (fast) GNM_LT_LAUNCH(false, true);
    else if (act) GNM_LT_LAUNCH(true, false);
    else GNM_LT_LAUNCH(false, false);
#undef GNM_LT_LAUNCH
    e = cudaGetLastError();
    return e == cudaSuccess ? GNM_OK : (int)e;
}

int gnm_linear_tc_abort_flag(int* aborted) {
    int v = 0, zero = 0;
    cudaError_t e = cudaMemcpyFromSymbol(&v, g_tc_abort, sizeof(int));
    if (e != cudaSuccess) return (int)e;
    e = cudaMemcpyToSymbol(g_tc_abort, &zero, sizeof(int));
    if (aborted) *aborted |= v;
    return e == cudaSuccess ? GNM_OK : (int)e;
}
