// Backward of one Linear -> BatchNorm (-> ReLU) unit on the 5th-generation tensor cores, input-gradient half
// (reference: autograd of models/mlp.py:48-49 and models/graphcnn.py:162-166). Same contract as the dx / stats_in part
// of gnm_linear_bwd (gnm_mlp_bwd.cu); the weight-gradient half is gnm_linear_wgrad_tc below.
//
//   dz = cA*dy + cB*z + cC (BatchNorm backward as an affine map of the two streams), so by linearity
//   dx[m, i] = sum_o dy[m,o] * (cA[o] W[o,i]) + sum_o z[m,o] * (cB[o] W[o,i]) + sum_o cC[o] W[o,i]
// i.e. two GEMMs into the same accumulator with the per-channel coefficients folded into two copies of W, plus a
// constant row. dy and z are split exactly into three bf16 planes each (tensor memory, one row per lane), the two
// scaled W copies into three planes each (shared memory), six MMAs per operand pair and 16-wide k-step.
// Epilogue, per 32-column half: TMEM -> a small shared tile (one row per lane), then a column-layout pass (eight lanes
// per row, float4 columns) next to the cp.async-staged x tile: + constant row, ReLU mask of the unit below
// (a = relu(x*in_scale + in_shift) > 0), its BatchNorm-backward reduction (sum dx, sum dx*xhat) accumulated per lane -
// no cross-lane transposition - and coalesced stores; one fp64 atomic per column and CTA at the end.
#include "gnm_common.cuh"
#include "gnm_tc.cuh"
#include "gnm_bn_tail.cuh"

namespace {

constexpr int BT_F = 64;
constexpr int BT_STAGES = 2;
constexpr int BT_A_COLS = 192;              // per stage: dy hi|mid|lo (96 columns) then z hi|mid|lo (96 columns)
constexpr int BT_D_COLS = 64;
constexpr int BT_A_TMEM0 = 2 * BT_D_COLS;
constexpr int BT_EPI_WARPS = 8, BT_PROD_WARPS = 8;
constexpr int BT_THREADS = (BT_EPI_WARPS + BT_PROD_WARPS + 1) * 32;
constexpr int BT_MMA_WARP = BT_EPI_WARPS + BT_PROD_WARPS;
constexpr int BT_W_PLANE = BT_F * BT_F * 2;
constexpr int BT_W_KCORE = 8 * 128;
constexpr int BT_PITCH = BT_F + 4;
constexpr int BT_STG = 32 * BT_PITCH * 4;
constexpr int BT_OPITCH = 32 + 4;            // epilogue half tile: 32 rows x 32 columns, padded
constexpr int BT_OSTG = 32 * BT_OPITCH * 4;
constexpr int BT_CONST_FLOATS = 5 * BT_F;   // c_row | in_scale | in_shift | in_mean | in_rstd
constexpr int BT_SMEM = 6 * BT_W_PLANE + BT_CONST_FLOATS * 4 + 256 + (BT_EPI_WARPS + BT_PROD_WARPS) * BT_STG + BT_EPI_WARPS * BT_OSTG;

struct LinBwdTcParams {
    const float* dy; int64_t lddy;
    const float* z; int64_t ldz;
    const float* coef;                     // [3][n_out]
    const float* w; int64_t ldw;           // [n_out][n_in]
    const float* x; int64_t ldx;           // input of the unit: previous pre-BN tensor (with in_scale) or plain input
    const float* in_scale; const float* in_shift; const float* in_mean; const float* in_rstd;
    float* dx; int64_t lddx;
    double* stats_in;
    int n_rows, n_out, n_in;
    BnTailDev tail;                        // BatchNorm-backward coefficients of the unit below from stats_in (kind 0: none)
};


// cp.async a warp's 32 rows x 64 floats (rows past n_rows zero-filled) into its padded staging tile
__device__ __forceinline__ void stage_rows_async(const float* base, int64_t ld, int n_rows, int n_cols, int row0, float* stg,
                                                 uint32_t stg_u32, int lane, bool fast) {
    if (fast) {
        // lane -> (row pair member lane >> 4, float4 column lane & 15); rows advance by two per copy
        const int nvalid = n_rows - row0 - (lane >> 4);          // copy i is in range iff 2 i < nvalid
        const char* src = reinterpret_cast<const char*>(base + (int64_t)(row0 + (lane >> 4)) * ld + (lane & 15) * 4);
        const int64_t step = 2 * ld * (int64_t)sizeof(float);
        uint32_t dst = stg_u32 + (uint32_t)(((lane >> 4) * BT_PITCH + (lane & 15) * 4) * 4);
        if (nvalid >= 31) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                src += step;
                dst += 2 * BT_PITCH * 4;
            }
        } else {
#pragma unroll 4
            for (int i = 0; i < 16; ++i) {
                const bool okr = 2 * i < nvalid;
                const int nbytes = okr ? 16 : 0;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(okr ? src : (const char*)base), "r"(nbytes) : "memory");
                src += step;
                dst += 2 * BT_PITCH * 4;
            }
        }
    } else {
        for (int e = lane; e < 32 * BT_F; e += 32) {
            const int rr = e >> 6, k = e & 63;
            stg[rr * BT_PITCH + k] = (row0 + rr < n_rows && k < n_cols) ? base[(int64_t)(row0 + rr) * ld + k] : 0.f;
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// ACT: the unit below is Linear -> BatchNorm -> ReLU (mask + its backward reduction in the epilogue); FAST: 64 x 64,
// aligned, dx requested - the general paths are compiled out (see linear_tc_kernel).
template <bool ACT, bool FAST>
__global__ void __launch_bounds__(BT_THREADS, 1) linear_bwd_dx_tc_kernel(const LinBwdTcParams p) {
    extern __shared__ __align__(1024) unsigned char bt_smem[];
    __shared__ __align__(8) uint64_t bars[2 * BT_STAGES + 4];
    __shared__ uint32_t s_tmem;
    __shared__ int s_abort;
    uint64_t* a_full = bars;
    uint64_t* a_empty = bars + BT_STAGES;
    uint64_t* acc_full = bars + 2 * BT_STAGES;
    uint64_t* acc_empty = acc_full + 2;
    unsigned char* sm_w = bt_smem;                                       // planes: WA hi|mid|lo, WB hi|mid|lo
    float* sm_c = reinterpret_cast<float*>(bt_smem + 6 * BT_W_PLANE);
    float* sm_stg = reinterpret_cast<float*>(bt_smem + 6 * BT_W_PLANE + BT_CONST_FLOATS * 4 + 256);
    float* sm_out = sm_stg + (BT_EPI_WARPS + BT_PROD_WARPS) * (32 * BT_PITCH);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    volatile int* abort_flag = &s_abort;
    const int n_tiles = (p.n_rows + 127) >> 7;
    constexpr bool act = ACT;

    // the B images fold the BatchNorm-backward coefficients, which come out of the tail of the kernel in front: wait first
    pdl_launch_dependents();
    pdl_wait();
    // ---- setup: B images. GEMM k = o (n_out), n = i (n_in); K-major core matrices: (n, k) -> (k/8)*1024 + (n/8)*128 + (n%8)*16 + (k%8)*2
    for (int e = tid; e < BT_F * BT_F / 2; e += BT_THREADS) {
        const int n = e >> 5, k = (e & 31) * 2;
        float w0 = 0.f, w1 = 0.f;
        if (n < p.n_in) {
            if (k < p.n_out) w0 = p.w[(int64_t)k * p.ldw + n];
            if (k + 1 < p.n_out) w1 = p.w[(int64_t)(k + 1) * p.ldw + n];
        }
        const float a0 = k < p.n_out ? p.coef[k] : 0.f, a1 = k + 1 < p.n_out ? p.coef[k + 1] : 0.f;
        const float b0 = k < p.n_out ? p.coef[p.n_out + k] : 0.f, b1 = k + 1 < p.n_out ? p.coef[p.n_out + k + 1] : 0.f;
        const int off = (k >> 3) * BT_W_KCORE + (n >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2;
        uint32_t h, m, l;
        split3x2(a0 * w0, a1 * w1, h, m, l);
        *reinterpret_cast<uint32_t*>(sm_w + off) = h;
        *reinterpret_cast<uint32_t*>(sm_w + BT_W_PLANE + off) = m;
        *reinterpret_cast<uint32_t*>(sm_w + 2 * BT_W_PLANE + off) = l;
        split3x2(b0 * w0, b1 * w1, h, m, l);
        *reinterpret_cast<uint32_t*>(sm_w + 3 * BT_W_PLANE + off) = h;
        *reinterpret_cast<uint32_t*>(sm_w + 4 * BT_W_PLANE + off) = m;
        *reinterpret_cast<uint32_t*>(sm_w + 5 * BT_W_PLANE + off) = l;
    }
    for (int i = tid; i < BT_F; i += BT_THREADS) {
        float c = 0.f;
        if (i < p.n_in)
            for (int o = 0; o < p.n_out; ++o) c = fmaf(p.coef[2 * p.n_out + o], p.w[(int64_t)o * p.ldw + i], c);
        sm_c[i] = c;
        sm_c[BT_F + i] = (act && i < p.n_in) ? p.in_scale[i] : 1.f;
        sm_c[2 * BT_F + i] = (act && i < p.n_in) ? p.in_shift[i] : 0.f;
        sm_c[3 * BT_F + i] = (act && i < p.n_in) ? p.in_mean[i] : 0.f;
        sm_c[4 * BT_F + i] = (act && i < p.n_in) ? p.in_rstd[i] : 0.f;
    }
    if (tid == 0) {
        s_abort = 0;
        for (int i = 0; i < BT_STAGES; ++i) { mbar_init(&a_full[i], BT_PROD_WARPS); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == BT_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;

    if (warp < BT_EPI_WARPS) {
        // ================================ epilogue (two groups of four warps, one per accumulator slot) ============
        // Row pass: the warp's 32 x 32 half of the accumulator goes TMEM -> registers -> a small shared tile, one row
        // per lane. Column pass: the half is re-read eight lanes per row (float4 columns) next to the staged x tile:
        // + constant row, ReLU mask, xhat, the two BatchNorm-backward column sums and the coalesced dx store all
        // happen in that layout, so no cross-lane transposition is needed.
        float cs1[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, cs2[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        const uint32_t my_slot = warp >> 2;
        const int q = warp & 3;
        float* stg = sm_stg + warp * (32 * BT_PITCH);
        float* obuf = sm_out + warp * (32 * BT_OPITCH);
        const uint32_t stg_u32 = smem_u32(stg);
        const bool fast_x = act && (FAST || (p.n_in == BT_F && (p.ldx & 3) == 0 && gnm_aligned16(p.x)));
        const bool fast_o = FAST || (p.dx != nullptr && p.n_in == BT_F && (p.lddx & 3) == 0 && gnm_aligned16(p.dx));
        const int sub = lane >> 3, c4l = (lane & 7) * 4;
        // prefetch x for this group's first tile
        if (act) {
            const int t0 = blockIdx.x + (int)my_slot * gridDim.x;
            if (t0 < n_tiles) stage_rows_async(p.x, p.ldx, p.n_rows, p.n_in, t0 * 128 + q * 32, stg, stg_u32, lane, fast_x);
        }
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < n_tiles && !*abort_flag; tile += gridDim.x, ++it) {
            const uint32_t slot = it & 1, ph = (it >> 1) & 1;
            if (slot != my_slot) continue;
            if (!mbar_wait<32>(&acc_full[slot], ph, abort_flag)) break;
            tc_fence_after();
            const int row0 = tile * 128 + q * 32;
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
                for (int c0 = 0; c0 < 32; c0 += 16) {
                    uint32_t t16[16];
                    tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + slot * BT_D_COLS + hf * 32 + c0, t16);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        *reinterpret_cast<uint4*>(obuf + lane * BT_OPITCH + c0 + j) = make_uint4(t16[j], t16[j + 1], t16[j + 2], t16[j + 3]);
                }
                if (hf == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[slot]);       // the accumulator has left TMEM: free the slot
                }
                __syncwarp();
                const int c = hf * 32 + c4l;
                const float4 cr = *reinterpret_cast<const float4*>(sm_c + c);
                const float4 sc = *reinterpret_cast<const float4*>(sm_c + BT_F + c);
                const float4 sh = *reinterpret_cast<const float4*>(sm_c + 2 * BT_F + c);
                const float4 mu = *reinterpret_cast<const float4*>(sm_c + 3 * BT_F + c);
                const float4 rs = *reinterpret_cast<const float4*>(sm_c + 4 * BT_F + c);
                const int nvalid = p.n_rows - row0 - sub;                  // row 4 i + sub is in range iff 4 i < nvalid
                char* dstp = p.dx != nullptr ? reinterpret_cast<char*>(p.dx + (int64_t)(row0 + sub) * p.lddx + c) : nullptr;
                const int64_t dstep = 4 * p.lddx * (int64_t)sizeof(float);
                // loads first (a per-row branch in front of them would serialise the shared-memory latencies)
#pragma unroll
                for (int i0 = 0; i0 < 8; i0 += 4) {
                float4 dd[4], xx[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    dd[i] = *reinterpret_cast<const float4*>(obuf + (4 * (i0 + i) + sub) * BT_OPITCH + c4l);
                    if (act) xx[i] = *reinterpret_cast<const float4*>(stg + (4 * (i0 + i) + sub) * BT_PITCH + c);
                }
#pragma unroll
                for (int i = i0; i < i0 + 4; ++i) {
                    float4 d = dd[i - i0];
                    const bool okr = 4 * i < nvalid;
                    d.x += cr.x; d.y += cr.y; d.z += cr.z; d.w += cr.w;
                    if (act) {
                        const float4 xv = xx[i - i0];
                        d.x = (okr && fmaf(xv.x, sc.x, sh.x) > 0.f) ? d.x : 0.f;
                        d.y = (okr && fmaf(xv.y, sc.y, sh.y) > 0.f) ? d.y : 0.f;
                        d.z = (okr && fmaf(xv.z, sc.z, sh.z) > 0.f) ? d.z : 0.f;
                        d.w = (okr && fmaf(xv.w, sc.w, sh.w) > 0.f) ? d.w : 0.f;
                        cs1[hf][0] += d.x; cs1[hf][1] += d.y; cs1[hf][2] += d.z; cs1[hf][3] += d.w;
                        cs2[hf][0] = fmaf(d.x, (xv.x - mu.x) * rs.x, cs2[hf][0]);
                        cs2[hf][1] = fmaf(d.y, (xv.y - mu.y) * rs.y, cs2[hf][1]);
                        cs2[hf][2] = fmaf(d.z, (xv.z - mu.z) * rs.z, cs2[hf][2]);
                        cs2[hf][3] = fmaf(d.w, (xv.w - mu.w) * rs.w, cs2[hf][3]);
                    }
                    if (dstp != nullptr && okr) {
                        if (fast_o) {
                            *reinterpret_cast<float4*>(dstp + i * dstep) = d;
                        } else {
                            float* o = reinterpret_cast<float*>(dstp + i * dstep);
                            if (c < p.n_in) o[0] = d.x;
                            if (c + 1 < p.n_in) o[1] = d.y;
                            if (c + 2 < p.n_in) o[2] = d.z;
                            if (c + 3 < p.n_in) o[3] = d.w;
                        }
                    }
                }
                }
                __syncwarp();                                           // the half tile is reused by the next half
            }
            if (act) {
                const int next_tile = tile + 2 * gridDim.x;
                if (next_tile < n_tiles)
                    stage_rows_async(p.x, p.ldx, p.n_rows, p.n_in, next_tile * 128 + q * 32, stg, stg_u32, lane, fast_x);
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (act && p.stats_in != nullptr) {
            // lanes with equal (lane & 7) hold the same columns for different rows: fold them, park the warp's 2 x 64
            // partial sums in its (now free) staging tile, add the eight warps in a fixed order and issue ONE fp64
            // atomic per column and CTA (same-address atomics serialise in L2)
            __syncwarp();
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float a1 = cs1[hf][u], a2 = cs2[hf][u];
                    a1 += __shfl_xor_sync(GNM_FULL_MASK, a1, 8);
                    a2 += __shfl_xor_sync(GNM_FULL_MASK, a2, 8);
                    a1 += __shfl_xor_sync(GNM_FULL_MASK, a1, 16);
                    a2 += __shfl_xor_sync(GNM_FULL_MASK, a2, 16);
                    if (lane < 8) {
                        stg[hf * 32 + c4l + u] = a1;
                        stg[BT_F + hf * 32 + c4l + u] = a2;
                    }
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(BT_EPI_WARPS * 32) : "memory");
            if (tid < 2 * BT_F) {
                float a = 0.f;
#pragma unroll
                for (int w = 0; w < BT_EPI_WARPS; ++w) a += sm_stg[w * (32 * BT_PITCH) + tid];
                const int c = tid & (BT_F - 1);
                if (c < p.n_in) atomicAdd(&p.stats_in[(tid >> 6) * p.n_in + c], (double)a);
            }
        }
    } else if (warp == BT_MMA_WARP) {
        // ================================ MMA issue ==============================================================
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BT_F >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint64_t wa = umma_desc(smem_u32(sm_w), BT_W_KCORE, 128);
            const uint64_t pl = (uint64_t)(BT_W_PLANE >> 4);
            const int ksteps = (p.n_out + 15) >> 4;
            uint32_t it = 0;
            bool ok = true;
            for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x, ++it) {
                const uint32_t s = it % BT_STAGES, aph = (it / BT_STAGES) & 1;
                const uint32_t slot = it & 1, ph = (it >> 1) & 1;
                if (!(ok = mbar_wait(&acc_empty[slot], ph ^ 1, abort_flag))) break;
                if (!(ok = mbar_wait(&a_full[s], aph, abort_flag))) break;
                tc_fence_after();
                const uint32_t d = tmem + slot * BT_D_COLS;
                const uint32_t a0 = tmem + BT_A_TMEM0 + s * BT_A_COLS;
                for (int ks = 0; ks < ksteps; ++ks) {
                    const uint64_t kofs = (uint64_t)(ks * 2 * BT_W_KCORE >> 4);
#pragma unroll
                    for (int op = 0; op < 2; ++op) {             // op 0: dy x (cA W), op 1: z x (cB W)
                        const uint32_t ah = a0 + op * 96 + ks * 8, am = ah + 32, al = ah + 64;
                        const uint64_t wh = wa + (uint64_t)op * 3 * pl + kofs, wm = wh + pl, wl = wh + 2 * pl;
                        umma_ts(d, ah, wh, idesc, (ks | op) ? 1u : 0u);
                        umma_ts(d, ah, wm, idesc, 1u);
                        umma_ts(d, am, wh, idesc, 1u);
                        umma_ts(d, ah, wl, idesc, 1u);
                        umma_ts(d, al, wh, idesc, 1u);
                        umma_ts(d, am, wm, idesc, 1u);
                    }
                }
                umma_commit(&a_empty[s]);
                umma_commit(&acc_full[slot]);
            }
        }
    } else {
        // ================================ producers: warps 8-11 stream dy, warps 12-15 stream z, every tile =========
        const int grp = (warp - BT_EPI_WARPS) >> 2;
        const int q = warp & 3;
        float* stg = sm_stg + warp * (32 * BT_PITCH);
        const uint32_t stg_u32 = smem_u32(stg);
        const float* src = grp == 0 ? p.dy : p.z;
        const int64_t ld = grp == 0 ? p.lddy : p.ldz;
        const bool fast = FAST || (p.n_out == BT_F && (ld & 3) == 0 && gnm_aligned16(src));
        if ((int)blockIdx.x < n_tiles) stage_rows_async(src, ld, p.n_rows, p.n_out, blockIdx.x * 128 + q * 32, stg, stg_u32, lane, fast);
        uint32_t it = 0;
        bool ok = true;
        for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x, ++it) {
            const uint32_t s = it % BT_STAGES, aph = (it / BT_STAGES) & 1;
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + BT_A_TMEM0 + s * BT_A_COLS + grp * 96;
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            bool waited = false;
#pragma unroll
            for (int k0 = 0; k0 < BT_F; k0 += 32) {
                uint32_t hi[16], mid[16], lo[16];
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 t = *reinterpret_cast<const float4*>(stg + lane * BT_PITCH + k0 + j);
                    split3x2(t.x, t.y, hi[j >> 1], mid[j >> 1], lo[j >> 1]);
                    split3x2(t.z, t.w, hi[(j >> 1) + 1], mid[(j >> 1) + 1], lo[(j >> 1) + 1]);
                }
                if (k0 == 32) {
                    __syncwarp();
                    const int next_tile = tile + gridDim.x;
                    if (next_tile < n_tiles)
                        stage_rows_async(src, ld, p.n_rows, p.n_out, next_tile * 128 + q * 32, stg, stg_u32, lane, fast);
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
    if (warp == BT_MMA_WARP) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
    if (ACT) bn_tail_run(p.tail);
}


// ---------------------------------------------------------------------------------------------------------------
// Weight-gradient half: dW[o,i] = sum_m dz[m,o] a[m,i], db[o] = sum_m dz[m,o] with dz = cA*dy + cB*z + cC, i.e.
//   dW = diag(cA) (dy^T a) + diag(cB) (z^T a) + cC (x) colsum(a),   db = cA*colsum(dy) + cB*colsum(z) + cC*n_rows
// One "TN" GEMM with the rows as the reduction axis: A operand = [dy | z]^T (M = 128: dy channels then z channels),
// B operand = a (N = 64), both MN-major in shared memory as three exact bf16 planes, 32 rows per stage. The B
// planes sit side by side (hi | mid | lo, N = 192) so the six kept products take three MMAs per 16-row k-step:
// A_hi x [hi|mid|lo], A_mid x [hi|mid], A_lo x [hi]; the three 64-column groups of the accumulator are summed in
// the epilogue. Column sums are accumulated by the producers on the way.
constexpr int WG_ROWS = 32;                          // rows (reduction length) per stage
constexpr int WG_STAGES = 3;
constexpr int WG_DEPTH = 4;                          // fp32 landing ring: chunks requested ahead of their conversion
constexpr int WG_KCORES = WG_ROWS / 8;
constexpr int WG_CORE_STRIDE = WG_KCORES * 128 + 16; // stride between 8-wide m / n cores (padded: conflict-free fill)
constexpr int WG_A_PLANE = 16 * WG_CORE_STRIDE;      // 128 m = 16 cores
constexpr int WG_B_BYTES = 24 * WG_CORE_STRIDE;      // 192 n = 24 cores (three planes side by side)
constexpr int WG_STAGE_BYTES = 3 * WG_A_PLANE + WG_B_BYTES;
constexpr int WG_PROD_WARPS = 8;
constexpr int WG_RPT = WG_ROWS * 16 / (WG_PROD_WARPS * 32);   // rows per producer thread and chunk
constexpr int WG_PT = WG_PROD_WARPS * 32;
constexpr int WG_THREADS = (WG_PROD_WARPS + 1) * 32;
constexpr int WG_LAND_CHUNK = 3 * WG_RPT * WG_PT * 16;      // per chunk: 3 float4 (three streams) per producer thread and row
constexpr int WG_SMEM = WG_STAGES * WG_STAGE_BYTES + WG_DEPTH * WG_LAND_CHUNK + 1024;

struct WgradTcParams {
    const float* dy; int64_t lddy;
    const float* z; int64_t ldz;
    const float* coef;
    const float* x; int64_t ldx;
    const float* in_scale; const float* in_shift;
    float* dw; int64_t lddw; float* db;
    int n_rows, n_out, n_in;
};

__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

template <bool ACT, bool FAST>
__global__ void __launch_bounds__(WG_THREADS, 1) linear_wgrad_tc_kernel(const WgradTcParams p) {
    extern __shared__ __align__(1024) unsigned char wg_smem[];
    __shared__ __align__(8) uint64_t bars[2 * WG_STAGES + 1];
    __shared__ uint32_t s_tmem;
    __shared__ int s_abort;
    __shared__ float s_sum[3 * BT_F];                  // column sums: dy | z | a
    uint64_t* full = bars;
    uint64_t* empty = bars + WG_STAGES;
    uint64_t* done = bars + 2 * WG_STAGES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    volatile int* abort_flag = &s_abort;
    const int n_chunks = (p.n_rows + WG_ROWS - 1) / WG_ROWS;
    constexpr bool act = ACT;

    pdl_launch_dependents();
    if (tid < 3 * BT_F) s_sum[tid] = 0.f;
    if (tid == 0) {
        s_abort = 0;
        for (int i = 0; i < WG_STAGES; ++i) { mbar_init(&full[i], WG_PROD_WARPS); mbar_init(&empty[i], 1); }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == WG_PROD_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    pdl_wait();

    if (warp < WG_PROD_WARPS) {
        // ================================ producers ============================================================
        // thread -> WG_RPT rows (kr0 + j * WG_PT / 16) of the 32-row chunk, float4 column c4 (fixed per thread), three streams.
        // Global -> shared "landing" slots with cp.async, WG_DEPTH - 1 chunks ahead of their conversion. Each thread
        // reads back exactly the slots it requested, so the only synchronisation is its own cp.async group count -
        // unlike a register prefetch, whose loads all share the warp's scoreboards (measured: a deeper register ring
        // gained nothing), this keeps 3 chunks = 72 KB per SM genuinely in flight.
        const int c4 = tid & 15, kr0 = tid >> 4;
        const bool fast = FAST || (p.n_out == BT_F && p.n_in == BT_F && ((p.lddy | p.ldz | p.ldx) & 3) == 0 &&
                                   gnm_aligned16(p.dy) && gnm_aligned16(p.z) && gnm_aligned16(p.x));
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (act) {
            const int c = c4 * 4;
            if (c < p.n_in) { sc.x = p.in_scale[c]; sh.x = p.in_shift[c]; }
            if (c + 1 < p.n_in) { sc.y = p.in_scale[c + 1]; sh.y = p.in_shift[c + 1]; }
            if (c + 2 < p.n_in) { sc.z = p.in_scale[c + 2]; sh.z = p.in_shift[c + 2]; }
            if (c + 3 < p.n_in) { sc.w = p.in_scale[c + 3]; sh.w = p.in_shift[c + 3]; }
        }
        float4 s_dy = make_float4(0.f, 0.f, 0.f, 0.f), s_z = s_dy, s_a = s_dy;
        unsigned char* land = wg_smem + (size_t)WG_STAGES * WG_STAGE_BYTES;
        float4* my_land = reinterpret_cast<float4*>(land) + tid;                  // slot (d, q) at my_land[(d * 3 * WG_RPT + q) * WG_PT]
        const uint32_t my_land_u32 = smem_u32(my_land);
        const int G = gridDim.x;
        auto request_chunk = [&](int chunk, int d) {
            if (chunk < n_chunks) {
                const int r0 = chunk * WG_ROWS + kr0;
                if (fast) {
#pragma unroll
                    for (int j = 0; j < WG_RPT; ++j) {
                        const int r = r0 + (WG_PT / 16) * j;
                        const bool okr = r < p.n_rows;
                        const int nbytes = okr ? 16 : 0;
                        const int64_t rr = okr ? r : 0;
                        const uint32_t dst = my_land_u32 + (uint32_t)((d * 3 * WG_RPT + j * 3) * WG_PT * 16);
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst),
                                     "l"(p.dy + rr * p.lddy + c4 * 4), "r"(nbytes) : "memory");
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + WG_PT * 16),
                                     "l"(p.z + rr * p.ldz + c4 * 4), "r"(nbytes) : "memory");
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + 2 * WG_PT * 16),
                                     "l"(p.x + rr * p.ldx + c4 * 4), "r"(nbytes) : "memory");
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < WG_RPT; ++j) {
                        const int r = r0 + (WG_PT / 16) * j;
                        const bool okr = r < p.n_rows;
                        float t[12];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int c = c4 * 4 + u;
                            t[u] = (okr && c < p.n_out) ? p.dy[(int64_t)r * p.lddy + c] : 0.f;
                            t[4 + u] = (okr && c < p.n_out) ? p.z[(int64_t)r * p.ldz + c] : 0.f;
                            t[8 + u] = (okr && c < p.n_in) ? p.x[(int64_t)r * p.ldx + c] : 0.f;
                        }
                        my_land[(d * 3 * WG_RPT + j * 3) * WG_PT] = make_float4(t[0], t[1], t[2], t[3]);
                        my_land[(d * 3 * WG_RPT + j * 3 + 1) * WG_PT] = make_float4(t[4], t[5], t[6], t[7]);
                        my_land[(d * 3 * WG_RPT + j * 3 + 2) * WG_PT] = make_float4(t[8], t[9], t[10], t[11]);
                    }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");      // one group per chunk slot, even when empty
        };
#pragma unroll
        for (int d = 0; d < WG_DEPTH - 1; ++d) request_chunk(blockIdx.x + d * G, d);
        uint32_t it = 0;
        bool ok = true;
        for (int chunk = blockIdx.x; chunk < n_chunks && ok; chunk += G, ++it) {
            const uint32_t s = it % WG_STAGES, ph = (it / WG_STAGES) & 1;
            const int d = it % WG_DEPTH;
            unsigned char* st = wg_smem + (size_t)s * WG_STAGE_BYTES;
            // slot (it + DEPTH - 1) % DEPTH was consumed by this thread in the previous iteration: safe to refill
            request_chunk(chunk + (WG_DEPTH - 1) * G, (it + WG_DEPTH - 1) % WG_DEPTH);
            asm volatile("cp.async.wait_group %0;" ::"n"(WG_DEPTH - 1) : "memory");
            float4 vdy[WG_RPT], vz[WG_RPT], va[WG_RPT];
            const int nvalid = p.n_rows - chunk * WG_ROWS - kr0;              // row j of this thread is valid iff (WG_PT / 16) j < nvalid
#pragma unroll
            for (int j = 0; j < WG_RPT; ++j) {
                vdy[j] = my_land[(d * 3 * WG_RPT + j * 3) * WG_PT];
                vz[j] = my_land[(d * 3 * WG_RPT + j * 3 + 1) * WG_PT];
                float4 a = my_land[(d * 3 * WG_RPT + j * 3 + 2) * WG_PT];
                if (act) {
                    const bool okr = (WG_PT / 16) * j < nvalid;
                    a.x = okr ? fmaxf(fmaf(a.x, sc.x, sh.x), 0.f) : 0.f;
                    a.y = okr ? fmaxf(fmaf(a.y, sc.y, sh.y), 0.f) : 0.f;
                    a.z = okr ? fmaxf(fmaf(a.z, sc.z, sh.z), 0.f) : 0.f;
                    a.w = okr ? fmaxf(fmaf(a.w, sc.w, sh.w), 0.f) : 0.f;
                }
                va[j] = a;
            }
            if (!(ok = mbar_wait<32>(&empty[s], ph ^ 1, abort_flag))) break;
#pragma unroll
            for (int j = 0; j < WG_RPT; ++j) {
                const int k = kr0 + (WG_PT / 16) * j;
                const int off = (c4 >> 1) * WG_CORE_STRIDE + (k >> 3) * 128 + (k & 7) * 16 + (c4 & 1) * 8;
                uint32_t h0, m0, l0, h1, m1, l1;
                split3x2(vdy[j].x, vdy[j].y, h0, m0, l0);
                split3x2(vdy[j].z, vdy[j].w, h1, m1, l1);
                *reinterpret_cast<uint2*>(st + off) = make_uint2(h0, h1);
                *reinterpret_cast<uint2*>(st + WG_A_PLANE + off) = make_uint2(m0, m1);
                *reinterpret_cast<uint2*>(st + 2 * WG_A_PLANE + off) = make_uint2(l0, l1);
                split3x2(vz[j].x, vz[j].y, h0, m0, l0);
                split3x2(vz[j].z, vz[j].w, h1, m1, l1);
                const int offz = off + 8 * WG_CORE_STRIDE;                 // z channels: m = 64 + c
                *reinterpret_cast<uint2*>(st + offz) = make_uint2(h0, h1);
                *reinterpret_cast<uint2*>(st + WG_A_PLANE + offz) = make_uint2(m0, m1);
                *reinterpret_cast<uint2*>(st + 2 * WG_A_PLANE + offz) = make_uint2(l0, l1);
                split3x2(va[j].x, va[j].y, h0, m0, l0);
                split3x2(va[j].z, va[j].w, h1, m1, l1);
                unsigned char* sb = st + 3 * WG_A_PLANE + off;
                *reinterpret_cast<uint2*>(sb) = make_uint2(h0, h1);
                *reinterpret_cast<uint2*>(sb + 8 * WG_CORE_STRIDE) = make_uint2(m0, m1);
                *reinterpret_cast<uint2*>(sb + 16 * WG_CORE_STRIDE) = make_uint2(l0, l1);
                s_dy.x += vdy[j].x; s_dy.y += vdy[j].y; s_dy.z += vdy[j].z; s_dy.w += vdy[j].w;
                s_z.x += vz[j].x; s_z.y += vz[j].y; s_z.z += vz[j].z; s_z.w += vz[j].w;
                s_a.x += va[j].x; s_a.y += va[j].y; s_a.z += va[j].z; s_a.w += va[j].w;
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        // column sums: lanes l and l^16 share c4 -> fold, then one shared atomic per column and warp
        float v[12] = {s_dy.x, s_dy.y, s_dy.z, s_dy.w, s_z.x, s_z.y, s_z.z, s_z.w, s_a.x, s_a.y, s_a.z, s_a.w};
#pragma unroll
        for (int u = 0; u < 12; ++u) v[u] += __shfl_xor_sync(GNM_FULL_MASK, v[u], 16);
        if (lane < 16) {
#pragma unroll
            for (int u = 0; u < 12; ++u) atomicAdd(&s_sum[(u >> 2) * BT_F + c4 * 4 + (u & 3)], v[u]);
        }
    } else if (lane == 0) {
        // ================================ MMA issue ==============================================================
        const uint32_t id_base = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t id192 = id_base | ((uint32_t)(192 >> 3) << 17), id128 = id_base | ((uint32_t)(128 >> 3) << 17),
                       id64 = id_base | ((uint32_t)(64 >> 3) << 17);
        uint32_t it = 0;
        bool ok = true;
        for (int chunk = blockIdx.x; chunk < n_chunks && ok; chunk += gridDim.x, ++it) {
            const uint32_t s = it % WG_STAGES, ph = (it / WG_STAGES) & 1;
            if (!(ok = mbar_wait(&full[s], ph, abort_flag))) break;
            tc_fence_after();
            const uint32_t base = smem_u32(wg_smem + (size_t)s * WG_STAGE_BYTES);
            const uint64_t ah = umma_desc(base, 128, WG_CORE_STRIDE);
            const uint64_t am = umma_desc(base + WG_A_PLANE, 128, WG_CORE_STRIDE);
            const uint64_t al = umma_desc(base + 2 * WG_A_PLANE, 128, WG_CORE_STRIDE);
            const uint64_t b = umma_desc(base + 3 * WG_A_PLANE, 128, WG_CORE_STRIDE);
#pragma unroll
            for (int ks = 0; ks < WG_ROWS / 16; ++ks) {
                const uint64_t ko = (uint64_t)(ks * 256 >> 4);
                umma_ss(tmem, ah + ko, b + ko, id192, (it | ks) ? 1u : 0u);
                umma_ss(tmem, am + ko, b + ko, id128, 1u);
                umma_ss(tmem, al + ko, b + ko, id64, 1u);
            }
            umma_commit(&empty[s]);
        }
        if (ok) umma_commit(done);
    }
    // ================================ epilogue (once per CTA) ====================================================
    __syncwarp();
    bool fin = mbar_wait<64>(done, 0, abort_flag);
    tc_fence_after();
    __syncthreads();
    float* tile = reinterpret_cast<float*>(wg_smem);                     // [128][65] floats, stage memory is free now
    if (fin && warp < 4 && (int)blockIdx.x < n_chunks) {
        const int m = warp * 32 + lane;
        const int o = m & 63;
        const float scale = o < p.n_out ? p.coef[(m >> 6) * p.n_out + o] : 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < 64; c0 += 16) {
            uint32_t g0[16], g1[16], g2[16];
            const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + c0;
            tmem_ld16(ta, g0);
            tmem_ld16(ta + 64, g1);
            tmem_ld16(ta + 128, g2);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j)
                tile[m * 65 + c0 + j] = scale * (__uint_as_float(g0[j]) + __uint_as_float(g1[j]) + __uint_as_float(g2[j]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (fin && (int)blockIdx.x < n_chunks) {
        int my_rows = 0;
        for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) my_rows += min(WG_ROWS, p.n_rows - chunk * WG_ROWS);
        for (int e = tid; e < p.n_out * p.n_in; e += WG_THREADS) {
            const int o = e / p.n_in, i = e - o * p.n_in;
            const float v = tile[o * 65 + i] + tile[(64 + o) * 65 + i] + p.coef[2 * p.n_out + o] * s_sum[2 * BT_F + i];
            atomicAdd(&p.dw[(int64_t)o * p.lddw + i], v);
        }
        if (p.db != nullptr) {
            for (int o = tid; o < p.n_out; o += WG_THREADS)
                atomicAdd(&p.db[o], p.coef[o] * s_sum[o] + p.coef[p.n_out + o] * s_sum[BT_F + o] +
                                        p.coef[2 * p.n_out + o] * (float)my_rows);
        }
    }
    if (warp == WG_PROD_WARPS) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------
// One-pass backward unit (64 x 64, aligned, dx requested): dX, dW, db and the ReLU mask / BatchNorm-backward reduction of
// the unit below from ONE read of dy, z and x - 16 M F bytes instead of the 28 M F of the two kernels above.
//   dz = cA*dy + cB*z + cC is formed in fp32 by the producers (one fma pair per element) and split ONCE into three exact
//   bf16 planes; a = relu(x*in_scale + in_shift) (or x) likewise. The dz planes of a 128-row tile live in shared memory
//   as 8 x 16-byte core matrices [8 rows][8 channels] and serve BOTH GEMMs through two descriptors over the same bytes:
//     GEMM 1 (dX, per tile):  D1[row, i]  = sum_o dz[row, o] W[o, i]      A = dz, K-major  (M = 128 rows, K = 64 channels)
//                             B = [W_hi | W_mid | W_lo] side by side (N = 192 / 128 / 64 for the hi / mid / lo plane of dz:
//                             the six kept products in three MMAs per k-step; the three column groups are summed on drain)
//     GEMM 2 (dW, whole CTA): D2[o, i]   += sum_row dz[row, o] a[row, i]  A = dz, MN-major (M = 128: hi channels | mid
//                             channels; a second M = 128 MMA starting at the lo plane, of which only lanes 0-63 count),
//                             B = [a_hi | a_mid | a_lo] (N = 192; 64 for lo), K = 16 rows per MMA, 32-row chunks.
//   Tensor memory: D1 columns 0-191 (one slot: its drain is short and GEMM 1 of the next tile is a whole tile away),
//   D2 columns 192-383 and 384-447, drained once per CTA.
//   Roles: warps 0-3 epilogue (drain D1, + mask / reduction / store in a column layout, as in the dx kernel above),
//   warps 4-11 producers (global -> cp.async landing ring, three 32-row chunks deep -> fma / relu -> split -> planes),
//   warp 12 MMA issue. Barriers: full[4] (chunk c of the tile: a planes + dz planes written), b_empty[2] (a-plane ring),
//   a_free (all MMAs of the tile retired: its dz planes may be overwritten), d1_full / d1_empty, done.
constexpr int OP_A_CS = 16 * 128 + 16;             // bytes between 8-channel cores of a dz plane (128 rows, padded)
constexpr int OP_A_PLANE = 8 * OP_A_CS;
constexpr int OP_A_BYTES = 3 * OP_A_PLANE;
constexpr int OP_B_CS = 4 * 128 + 16;              // bytes between 8-column cores of the a planes (32 rows, padded)
constexpr int OP_B_SLOT = 24 * OP_B_CS;
constexpr int OP_W_KCORE = 24 * 128;               // bytes between 8-channel k-cores of [W_hi | W_mid | W_lo]
constexpr int OP_W_BYTES = 8 * OP_W_KCORE;
constexpr int OP_EPI_WARPS = 4, OP_PROD_WARPS = 8;
constexpr int OP_PT = OP_PROD_WARPS * 32;
constexpr int OP_THREADS = (OP_EPI_WARPS + OP_PROD_WARPS + 1) * 32;
constexpr int OP_MMA_WARP = OP_EPI_WARPS + OP_PROD_WARPS;
constexpr int OP_DEPTH = 3;                        // landing ring: 32-row chunks requested ahead of their conversion
constexpr int OP_LAND_CHUNK = 3 * 2 * OP_PT * 16;  // three streams x two rows per producer thread x 16 B
constexpr int OP_OFF_A = OP_W_BYTES;
constexpr int OP_OFF_B = OP_OFF_A + OP_A_BYTES;
constexpr int OP_OFF_LAND = OP_OFF_B + 2 * OP_B_SLOT;
constexpr int OP_OFF_OBUF = OP_OFF_LAND + OP_DEPTH * OP_LAND_CHUNK;
constexpr int OP_OFF_X = OP_OFF_OBUF + OP_EPI_WARPS * BT_OSTG;
constexpr int OP_OFF_C = OP_OFF_X + OP_EPI_WARPS * BT_STG;
constexpr int OP_SMEM = OP_OFF_C + 4 * BT_F * 4;
constexpr int OP_D2A = 192, OP_D2B = 384;
static_assert(OP_SMEM <= 232448 - 512, "one-pass backward unit: shared memory");
static_assert(128 * 65 * 4 <= OP_DEPTH * OP_LAND_CHUNK, "dW staging tile reuses the landing ring");

struct LinBwd1Params {
    const float* dy; int64_t lddy;
    const float* z; int64_t ldz;
    const float* coef;
    const float* w; int64_t ldw;
    const float* x; int64_t ldx;
    const float* in_scale; const float* in_shift; const float* in_mean; const float* in_rstd;
    float* dx; int64_t lddx;
    float* dw; int64_t lddw; float* db;
    double* stats_in;
    int n_rows;
    BnTailDev tail;
};

// the producers' arithmetic on one chunk: dz = cA*dy + cB*z + cC and a = relu(x*scale + shift) for this thread's two rows
// (landing slots u*3 + {0, 1, 2} = dy, z, x), zero for rows past the end (CHECK), column sums of dz
template <bool ACT, bool CHECK>
__device__ __forceinline__ void op_rows(const float4* land, int row_base, int n_rows, const float4& cA, const float4& cB,
                                        const float4& cC, const float4& sc, const float4& sh, float4 (&vdz)[2],
                                        float4 (&va)[2], float4& s_dz) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const float4 g = land[(u * 3) * OP_PT];
        const float4 zz = land[(u * 3 + 1) * OP_PT];
        float4 a = land[(u * 3 + 2) * OP_PT];
        const bool okr = !CHECK || row_base + 16 * u < n_rows;
        float4 dz;
        dz.x = okr ? fmaf(cA.x, g.x, fmaf(cB.x, zz.x, cC.x)) : 0.f;
        dz.y = okr ? fmaf(cA.y, g.y, fmaf(cB.y, zz.y, cC.y)) : 0.f;
        dz.z = okr ? fmaf(cA.z, g.z, fmaf(cB.z, zz.z, cC.z)) : 0.f;
        dz.w = okr ? fmaf(cA.w, g.w, fmaf(cB.w, zz.w, cC.w)) : 0.f;
        if (ACT) {
            a.x = okr ? fmaxf(fmaf(a.x, sc.x, sh.x), 0.f) : 0.f;
            a.y = okr ? fmaxf(fmaf(a.y, sc.y, sh.y), 0.f) : 0.f;
            a.z = okr ? fmaxf(fmaf(a.z, sc.z, sh.z), 0.f) : 0.f;
            a.w = okr ? fmaxf(fmaf(a.w, sc.w, sh.w), 0.f) : 0.f;
        }
        vdz[u] = dz;
        va[u] = a;
        s_dz.x += dz.x; s_dz.y += dz.y; s_dz.z += dz.z; s_dz.w += dz.w;
    }
}

template <bool ACT>
__global__ void __launch_bounds__(OP_THREADS, 1) linear_bwd_onepass_tc_kernel(const LinBwd1Params p) {
    extern __shared__ __align__(1024) unsigned char op_smem[];
    __shared__ __align__(8) uint64_t bars[4 + 2 + 4];
    __shared__ uint32_t s_tmem;
    __shared__ int s_abort;
    __shared__ float s_sum[BT_F];                                        // column sums of dz (db)
    uint64_t* full = bars;                 // [4]
    uint64_t* b_empty = bars + 4;          // [2]
    uint64_t* a_free = bars + 6;
    uint64_t* d1_full = bars + 7;
    uint64_t* d1_empty = bars + 8;
    uint64_t* done = bars + 9;
    unsigned char* sm_w = op_smem;
    unsigned char* sm_a = op_smem + OP_OFF_A;
    unsigned char* sm_b = op_smem + OP_OFF_B;
    unsigned char* sm_land = op_smem + OP_OFF_LAND;
    float* sm_obuf = reinterpret_cast<float*>(op_smem + OP_OFF_OBUF);
    float* sm_x = reinterpret_cast<float*>(op_smem + OP_OFF_X);
    float* sm_c = reinterpret_cast<float*>(op_smem + OP_OFF_C);          // in_scale | in_shift | in_mean | in_rstd
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    volatile int* abort_flag = &s_abort;
    const int n_tiles = (p.n_rows + 127) >> 7;
    const int G = gridDim.x;

    pdl_launch_dependents();
    pdl_wait();
    // ---- setup: [W_hi | W_mid | W_lo], K-major: (n' = plane * 64 + i, k = o) -> (k/8)*KCORE + (n'/8)*128 + (n'%8)*16 + (k%8)*2
    // (consecutive threads read consecutive i of two rows o, o + 1: coalesced; all loads of a thread are in flight
    // before the first split - the set-up is one global round trip, not five)
    {
        constexpr int NP = (BT_F * BT_F / 2 + OP_THREADS - 1) / OP_THREADS;
        float w0[NP], w1[NP];
#pragma unroll
        for (int u = 0; u < NP; ++u) {
            const int e = tid + u * OP_THREADS;
            const int n = e & 63, k = (e >> 6) * 2;
            w0[u] = w1[u] = 0.f;
            if (e < BT_F * BT_F / 2) {
                w0[u] = __ldg(p.w + (int64_t)k * p.ldw + n);
                w1[u] = __ldg(p.w + (int64_t)(k + 1) * p.ldw + n);
            }
        }
#pragma unroll
        for (int u = 0; u < NP; ++u) {
            const int e = tid + u * OP_THREADS;
            if (e >= BT_F * BT_F / 2) break;
            const int n = e & 63, k = (e >> 6) * 2;
            uint32_t h, m, l;
            split3x2(w0[u], w1[u], h, m, l);
            const int off = (k >> 3) * OP_W_KCORE + (n >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2;
            *reinterpret_cast<uint32_t*>(sm_w + off) = h;
            *reinterpret_cast<uint32_t*>(sm_w + 8 * 128 + off) = m;
            *reinterpret_cast<uint32_t*>(sm_w + 16 * 128 + off) = l;
        }
    }
    for (int i = tid; i < BT_F; i += OP_THREADS) {
        sm_c[i] = ACT ? p.in_scale[i] : 1.f;
        sm_c[BT_F + i] = ACT ? p.in_shift[i] : 0.f;
        sm_c[2 * BT_F + i] = ACT ? p.in_mean[i] : 0.f;
        sm_c[3 * BT_F + i] = ACT ? p.in_rstd[i] : 0.f;
        s_sum[i] = 0.f;
    }
    if (tid == 0) {
        s_abort = 0;
        for (int i = 0; i < 4; ++i) mbar_init(&full[i], OP_PROD_WARPS);
        for (int i = 0; i < 2; ++i) mbar_init(&b_empty[i], 1);
        mbar_init(a_free, 1);
        mbar_init(d1_full, 1);
        mbar_init(d1_empty, OP_EPI_WARPS);
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == OP_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;

    if (warp < OP_EPI_WARPS) {
        // ================================ epilogue: drain D1, mask / reduce / store =================================
        float cs1[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, cs2[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        const int q = warp;
        float* stg = sm_x + warp * (32 * BT_PITCH);
        float* obuf = sm_obuf + warp * (32 * BT_OPITCH);
        const uint32_t stg_u32 = smem_u32(stg);
        const int sub = lane >> 3, c4l = (lane & 7) * 4;
        if (ACT && (int)blockIdx.x < n_tiles)
            stage_rows_async(p.x, p.ldx, p.n_rows, BT_F, blockIdx.x * 128 + q * 32, stg, stg_u32, lane, true);
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < n_tiles && !*abort_flag; tile += G, ++it) {
            if (!mbar_wait<32>(d1_full, it & 1, abort_flag)) break;
            tc_fence_after();
            const int row0 = tile * 128 + q * 32;
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
                for (int c0 = 0; c0 < 32; c0 += 16) {
                    uint32_t g0[16], g1[16], g2[16];
                    const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + hf * 32 + c0;
                    tmem_ld16(ta, g0);
                    tmem_ld16(ta + 64, g1);
                    tmem_ld16(ta + 128, g2);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        float v[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            v[e] = (__uint_as_float(g2[j + e]) + __uint_as_float(g1[j + e])) + __uint_as_float(g0[j + e]);
                        *reinterpret_cast<float4*>(obuf + lane * BT_OPITCH + c0 + j) = make_float4(v[0], v[1], v[2], v[3]);
                    }
                }
                if (hf == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(d1_empty);               // D1 has left tensor memory
                }
                __syncwarp();
                const int c = hf * 32 + c4l;
                const float4 sc = *reinterpret_cast<const float4*>(sm_c + c);
                const float4 sh = *reinterpret_cast<const float4*>(sm_c + BT_F + c);
                const float4 mu = *reinterpret_cast<const float4*>(sm_c + 2 * BT_F + c);
                const float4 rs = *reinterpret_cast<const float4*>(sm_c + 3 * BT_F + c);
                const int nvalid = p.n_rows - row0 - sub;                  // row 4 i + sub is in range iff 4 i < nvalid
                char* dstp = reinterpret_cast<char*>(p.dx + (int64_t)(row0 + sub) * p.lddx + c);
                const int64_t dstep = 4 * p.lddx * (int64_t)sizeof(float);
#pragma unroll
                for (int i0 = 0; i0 < 8; i0 += 4) {
                    float4 dd[4], xx[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        dd[i] = *reinterpret_cast<const float4*>(obuf + (4 * (i0 + i) + sub) * BT_OPITCH + c4l);
                        if (ACT) xx[i] = *reinterpret_cast<const float4*>(stg + (4 * (i0 + i) + sub) * BT_PITCH + c);
                    }
#pragma unroll
                    for (int i = i0; i < i0 + 4; ++i) {
                        float4 d = dd[i - i0];
                        const bool okr = 4 * i < nvalid;
                        if (ACT) {
                            const float4 xv = xx[i - i0];
                            d.x = (okr && fmaf(xv.x, sc.x, sh.x) > 0.f) ? d.x : 0.f;
                            d.y = (okr && fmaf(xv.y, sc.y, sh.y) > 0.f) ? d.y : 0.f;
                            d.z = (okr && fmaf(xv.z, sc.z, sh.z) > 0.f) ? d.z : 0.f;
                            d.w = (okr && fmaf(xv.w, sc.w, sh.w) > 0.f) ? d.w : 0.f;
                            cs1[hf][0] += d.x; cs1[hf][1] += d.y; cs1[hf][2] += d.z; cs1[hf][3] += d.w;
                            cs2[hf][0] = fmaf(d.x, (xv.x - mu.x) * rs.x, cs2[hf][0]);
                            cs2[hf][1] = fmaf(d.y, (xv.y - mu.y) * rs.y, cs2[hf][1]);
                            cs2[hf][2] = fmaf(d.z, (xv.z - mu.z) * rs.z, cs2[hf][2]);
                            cs2[hf][3] = fmaf(d.w, (xv.w - mu.w) * rs.w, cs2[hf][3]);
                        }
                        if (okr) *reinterpret_cast<float4*>(dstp + i * dstep) = d;
                    }
                }
                __syncwarp();                                           // the half tile is reused by the next half
            }
            if (ACT) {
                const int next_tile = tile + G;
                if (next_tile < n_tiles)
                    stage_rows_async(p.x, p.ldx, p.n_rows, BT_F, next_tile * 128 + q * 32, stg, stg_u32, lane, true);
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (ACT && p.stats_in != nullptr) {
            __syncwarp();
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float a1 = cs1[hf][u], a2 = cs2[hf][u];
                    a1 += __shfl_xor_sync(GNM_FULL_MASK, a1, 8);
                    a2 += __shfl_xor_sync(GNM_FULL_MASK, a2, 8);
                    a1 += __shfl_xor_sync(GNM_FULL_MASK, a1, 16);
                    a2 += __shfl_xor_sync(GNM_FULL_MASK, a2, 16);
                    if (lane < 8) {
                        stg[hf * 32 + c4l + u] = a1;
                        stg[BT_F + hf * 32 + c4l + u] = a2;
                    }
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(OP_EPI_WARPS * 32) : "memory");
            if (tid < 2 * BT_F) {
                float a = 0.f;
#pragma unroll
                for (int w = 0; w < OP_EPI_WARPS; ++w) a += sm_x[w * (32 * BT_PITCH) + tid];
                atomicAdd(&p.stats_in[(tid >> 6) * BT_F + (tid & (BT_F - 1))], (double)a);
            }
        }
    } else if (warp == OP_MMA_WARP) {
        // ================================ MMA issue (one thread) ====================================================
        if (lane == 0) {
            const uint32_t id_k = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24);      // A, B K-major
            const uint32_t id_t = id_k | (1u << 15) | (1u << 16);                                           // A, B MN-major
            const uint32_t n192 = (uint32_t)(192 >> 3) << 17, n128 = (uint32_t)(128 >> 3) << 17, n64 = (uint32_t)(64 >> 3) << 17;
            const uint32_t a_base = smem_u32(sm_a), b_base = smem_u32(sm_b), w_base = smem_u32(sm_w);
            uint32_t it = 0, cc = 0;
            bool ok = true;
            for (int tile = blockIdx.x; tile < n_tiles && ok; tile += G, ++it) {
                for (int c = 0; c < 4 && ok; ++c, ++cc) {
                    if (!(ok = mbar_wait(&full[c], it & 1, abort_flag))) break;
                    tc_fence_after();
                    if (c == 3) {
                        // GEMM 1: all 128 rows, four k-steps of 16 channels (two channel cores). Issued BEFORE the last
                        // chunk's GEMM 2 so that a_free fires as early as possible: the producers then overwrite the dz
                        // planes chunk by chunk from row 0, and reach the rows GEMM 2 of this chunk still reads only after
                        // waiting for its a-plane slot (b_empty), i.e. after it has retired.
                        if (!(ok = mbar_wait(d1_empty, (it & 1) ^ 1, abort_flag))) break;
                        tc_fence_after();
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t a_h = umma_desc(a_base + (uint32_t)(ks * 2 * OP_A_CS), OP_A_CS, 128);
                            const uint64_t pl = (uint64_t)(OP_A_PLANE >> 4);
                            const uint64_t b_d = umma_desc(w_base + (uint32_t)(ks * 2 * OP_W_KCORE), OP_W_KCORE, 128);
                            umma_ss(tmem, a_h, b_d, id_k | n192, ks ? 1u : 0u);
                            umma_ss(tmem, a_h + pl, b_d, id_k | n128, 1u);
                            umma_ss(tmem, a_h + 2 * pl, b_d, id_k | n64, 1u);
                        }
                        umma_commit(d1_full);
                        umma_commit(a_free);
                    }
                    const uint32_t slot = cc & 1;
                    // GEMM 2: rows c*32 .. c*32+31 of the tile (row cores 4 c ..), two k-steps of 16 rows
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const uint64_t a_hm = umma_desc(a_base + (uint32_t)((c * 4 + ks * 2) * 128), 128, OP_A_CS);
                        const uint64_t a_lo = umma_desc(a_base + 2 * OP_A_PLANE + (uint32_t)((c * 4 + ks * 2) * 128), 128, OP_A_CS);
                        const uint64_t b_d = umma_desc(b_base + slot * OP_B_SLOT + (uint32_t)(ks * 256), 128, OP_B_CS);
                        const uint32_t acc = (cc | (uint32_t)ks) ? 1u : 0u;
                        umma_ss(tmem + OP_D2A, a_hm, b_d, id_t | n192, acc);
                        umma_ss(tmem + OP_D2B, a_lo, b_d, id_t | n64, acc);
                    }
                    umma_commit(&b_empty[slot]);
                }
                if (!ok) break;
            }
            if (ok) umma_commit(done);
        }
    } else {
        // ================================ producers =================================================================
        // thread -> float4 column c4 (fixed) and rows kr0, kr0 + 16 of every 32-row chunk, three streams; global ->
        // landing ring with cp.async, OP_DEPTH - 1 chunks ahead; every thread reads back exactly what it requested
        const int ptid = tid - OP_EPI_WARPS * 32;
        const int c4 = ptid & 15, kr0 = ptid >> 4;
        const float4 cA = *reinterpret_cast<const float4*>(p.coef + c4 * 4);
        const float4 cB = *reinterpret_cast<const float4*>(p.coef + BT_F + c4 * 4);
        const float4 cC = *reinterpret_cast<const float4*>(p.coef + 2 * BT_F + c4 * 4);
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ACT) {
            sc = *reinterpret_cast<const float4*>(p.in_scale + c4 * 4);
            sh = *reinterpret_cast<const float4*>(p.in_shift + c4 * 4);
        }
        float4 s_dz = make_float4(0.f, 0.f, 0.f, 0.f);
        float4* my_land = reinterpret_cast<float4*>(sm_land) + ptid;      // slot (d, q) at my_land[(d * 6 + q) * OP_PT]
        const uint32_t my_land_u32 = smem_u32(my_land);
        const int n_chunks = ((int)blockIdx.x < n_tiles) ? ((n_tiles - 1 - (int)blockIdx.x) / G + 1) * 4 : 0;   // this CTA's
        auto request_chunk = [&](int j, int d) {
            if (j < n_chunks) {
                const int r0 = ((int)blockIdx.x + (j >> 2) * G) * 128 + (j & 3) * 32 + kr0;
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int r = r0 + 16 * u;
                    const bool okr = r < p.n_rows;
                    const int nbytes = okr ? 16 : 0;
                    const int64_t rr = okr ? r : 0;
                    const uint32_t dst = my_land_u32 + (uint32_t)((d * 6 + u * 3) * OP_PT * 16);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst),
                                 "l"(p.dy + rr * p.lddy + c4 * 4), "r"(nbytes) : "memory");
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + OP_PT * 16),
                                 "l"(p.z + rr * p.ldz + c4 * 4), "r"(nbytes) : "memory");
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + 2 * OP_PT * 16),
                                 "l"(p.x + rr * p.ldx + c4 * 4), "r"(nbytes) : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");          // one group per slot, even when empty
        };
#pragma unroll
        for (int d = 0; d < OP_DEPTH - 1; ++d) request_chunk(d, d);
        bool ok = true;
        for (int j = 0; j < n_chunks && ok; ++j) {
            const int d = j % OP_DEPTH;
            const int c = j & 3;
            // slot (j + DEPTH - 1) % DEPTH was consumed by this thread in the previous iteration: safe to refill
            request_chunk(j + OP_DEPTH - 1, (j + OP_DEPTH - 1) % OP_DEPTH);
            asm volatile("cp.async.wait_group %0;" ::"n"(OP_DEPTH - 1) : "memory");
            const int row_base = ((int)blockIdx.x + (j >> 2) * G) * 128 + c * 32 + kr0;
            float4 vdz[2], va[2];
            // every chunk but the last of the matrix lies wholly inside it: no per-row range checks there
            if (row_base - kr0 + 31 < p.n_rows)
                op_rows<ACT, false>(my_land + (size_t)d * 6 * OP_PT, row_base, p.n_rows, cA, cB, cC, sc, sh, vdz, va, s_dz);
            else
                op_rows<ACT, true>(my_land + (size_t)d * 6 * OP_PT, row_base, p.n_rows, cA, cB, cC, sc, sh, vdz, va, s_dz);
            // a planes -> ring slot (free once GEMM 2 of the chunk two back has retired)
            const uint32_t slot = (uint32_t)j & 1, bph = ((uint32_t)j >> 1) & 1;
            if (!(ok = mbar_wait<32>(&b_empty[slot], bph ^ 1, abort_flag))) break;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int k = kr0 + 16 * u;
                unsigned char* sb = sm_b + slot * OP_B_SLOT + (c4 >> 1) * OP_B_CS + (k >> 3) * 128 + (k & 7) * 16 + (c4 & 1) * 8;
                uint32_t h0, m0, l0, h1, m1, l1;
                split3x2(va[u].x, va[u].y, h0, m0, l0);
                split3x2(va[u].z, va[u].w, h1, m1, l1);
                *reinterpret_cast<uint2*>(sb) = make_uint2(h0, h1);
                *reinterpret_cast<uint2*>(sb + 8 * OP_B_CS) = make_uint2(m0, m1);
                *reinterpret_cast<uint2*>(sb + 16 * OP_B_CS) = make_uint2(l0, l1);
            }
            // dz planes -> the tile buffer (free once every MMA of the previous tile has retired)
            if (c == 0 && !(ok = mbar_wait<32>(a_free, (((uint32_t)j >> 2) & 1) ^ 1, abort_flag))) break;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int k = c * 32 + kr0 + 16 * u;
                unsigned char* sa = sm_a + (c4 >> 1) * OP_A_CS + (k >> 3) * 128 + (k & 7) * 16 + (c4 & 1) * 8;
                uint32_t h0, m0, l0, h1, m1, l1;
                split3x2(vdz[u].x, vdz[u].y, h0, m0, l0);
                split3x2(vdz[u].z, vdz[u].w, h1, m1, l1);
                *reinterpret_cast<uint2*>(sa) = make_uint2(h0, h1);
                *reinterpret_cast<uint2*>(sa + OP_A_PLANE) = make_uint2(m0, m1);
                *reinterpret_cast<uint2*>(sa + 2 * OP_A_PLANE) = make_uint2(l0, l1);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[c]);
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        // column sums of dz: lanes l and l ^ 16 share c4 -> fold, one shared atomic per column and warp
        float v[4] = {s_dz.x, s_dz.y, s_dz.z, s_dz.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] += __shfl_xor_sync(GNM_FULL_MASK, v[u], 16);
        if (lane < 16) {
#pragma unroll
            for (int u = 0; u < 4; ++u) atomicAdd(&s_sum[c4 * 4 + u], v[u]);
        }
    }
    // ================================ dW / db (once per CTA) ==========================================================
    __syncwarp();
    const bool fin = mbar_wait<64>(done, 0, abort_flag);
    tc_fence_after();
    __syncthreads();
    float* tile = reinterpret_cast<float*>(sm_land);                     // [128][65] floats: the landing ring is free now
    const bool have = (int)blockIdx.x < n_tiles;
    if (fin && have && warp < 4) {
        const int m = warp * 32 + lane;
#pragma unroll 1
        for (int c0 = 0; c0 < 64; c0 += 16) {
            uint32_t g0[16], g1[16], g2[16], g3[16];
            const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + OP_D2A + c0;
            tmem_ld16(ta, g0);
            tmem_ld16(ta + 64, g1);
            tmem_ld16(ta + 128, g2);
            if (warp < 2) tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + OP_D2B + c0, g3);   // lo plane: lanes 0-63 only
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float t = (__uint_as_float(g2[j]) + __uint_as_float(g1[j])) + __uint_as_float(g0[j]);
                if (warp < 2) t += __uint_as_float(g3[j]);
                tile[m * 65 + c0 + j] = t;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (fin && have) {
        for (int e = tid; e < BT_F * BT_F; e += OP_THREADS) {
            const int o = e >> 6, i = e & 63;
            atomicAdd(&p.dw[(int64_t)o * p.lddw + i], tile[o * 65 + i] + tile[(64 + o) * 65 + i]);
        }
        if (p.db != nullptr && tid < BT_F) atomicAdd(&p.db[tid], s_sum[tid]);
    }
    if (warp == OP_MMA_WARP) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
    if (ACT) bn_tail_run(p.tail);
}

}  // namespace

int gnm_launch_linear_bwd_dx_tc(const float* dy, int64_t lddy, const float* z, int64_t ldz, const float* coef,
                                const float* x, int64_t ldx, const float* in_scale, const float* in_shift,
                                const float* in_mean, const float* in_rstd, const float* w, int64_t ldw, float* dx,
                                int64_t lddx, double* stats_in, int n_rows, int n_out, int n_in, const gnm_bn_tail* tail,
                                cudaStream_t stream) {
    if (n_in > BT_F || n_out > BT_F || n_in < 1 || n_out < 1) return GNM_ERR_TOO_LARGE;
    if (tail != nullptr && (stats_in == nullptr || in_scale == nullptr)) return GNM_ERR_BAD_ARG;
    int dev = 0, sms = 148, major = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) return GNM_ERR_TOO_LARGE;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    LinBwdTcParams p;
    p.dy = dy; p.lddy = lddy; p.z = z; p.ldz = ldz; p.coef = coef; p.w = w; p.ldw = ldw; p.x = x; p.ldx = ldx;
    p.in_scale = in_scale; p.in_shift = in_shift; p.in_mean = in_mean; p.in_rstd = in_rstd; p.dx = dx; p.lddx = lddx;
    p.stats_in = stats_in; p.n_rows = n_rows; p.n_out = n_out; p.n_in = n_in;
    const int trc = bn_tail_args(tail, stats_in, n_in, &p.tail);
    if (trc != GNM_OK) return trc;
    const int tiles = (n_rows + 127) / 128;
    const int grid = tiles < sms ? tiles : sms;
    const bool act = in_scale != nullptr;
    const bool fast = n_in == BT_F && n_out == BT_F && dx != nullptr && ((lddy | ldz | lddx) & 3) == 0 && gnm_aligned16(dy) &&
                      gnm_aligned16(z) && gnm_aligned16(dx) && (!act || ((ldx & 3) == 0 && gnm_aligned16(x)));
    cudaError_t e;
#define GNM_BT_LAUNCH(A, F)                                                                                                  \
    do {                                                                                                                     \
        e = cudaFuncSetAttribute(linear_bwd_dx_tc_kernel<A, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, BT_SMEM);       \
        if (e != cudaSuccess) return (int)e;                                                                                 \
        gnm_count_launch(GNM_K_LINEAR_BWD_DX_TC);                                                                            \
        e = gnm_launch_pdl<LinBwdTcParams>(linear_bwd_dx_tc_kernel<A, F>, grid, BT_THREADS, BT_SMEM, stream, p);             \
        if (e != cudaSuccess) return (int)e;                                                                                 \
    } while (0)
    if (act && fast) GNM_BT_LAUNCH(true, true);
    else if (fast) GNM_BT_LAUNCH(false, true);
    else if (act) GNM_BT_LAUNCH(true, false);
    else GNM_BT_LAUNCH(false, false);
#undef GNM_BT_LAUNCH
    e = cudaGetLastError();
    return e == cudaSuccess ? GNM_OK : (int)e;
}

int gnm_linear_bwd_tc_abort_flag(int* aborted) {
    int v = 0, zero = 0;
    cudaError_t e = cudaMemcpyFromSymbol(&v, g_tc_abort, sizeof(int));
    if (e != cudaSuccess) return (int)e;
    e = cudaMemcpyToSymbol(g_tc_abort, &zero, sizeof(int));
    if (aborted) *aborted |= v;
    return e == cudaSuccess ? GNM_OK : (int)e;
}

int gnm_launch_linear_wgrad_tc(const float* dy, int64_t lddy, const float* z, int64_t ldz, const float* coef,
                               const float* x, int64_t ldx, const float* in_scale, const float* in_shift, float* dw,
                               int64_t lddw, float* dbias, int n_rows, int n_out, int n_in, cudaStream_t stream) {
    if (n_in > BT_F || n_out > BT_F || n_in < 1 || n_out < 1) return GNM_ERR_TOO_LARGE;
    int dev = 0, sms = 148, major = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) return GNM_ERR_TOO_LARGE;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    WgradTcParams p;
    p.dy = dy; p.lddy = lddy; p.z = z; p.ldz = ldz; p.coef = coef; p.x = x; p.ldx = ldx; p.in_scale = in_scale;
    p.in_shift = in_shift; p.dw = dw; p.lddw = lddw; p.db = dbias; p.n_rows = n_rows; p.n_out = n_out; p.n_in = n_in;
    const int chunks = (n_rows + WG_ROWS - 1) / WG_ROWS;
    const int grid = chunks < sms ? chunks : sms;
    const bool act = in_scale != nullptr;
    const bool fast = n_out == BT_F && n_in == BT_F && ((lddy | ldz | ldx) & 3) == 0 && gnm_aligned16(dy) && gnm_aligned16(z) &&
                      gnm_aligned16(x);
    cudaError_t e;
#define GNM_WG_LAUNCH(A, F)                                                                                                 \
    do {                                                                                                                    \
        e = cudaFuncSetAttribute(linear_wgrad_tc_kernel<A, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM);       \
        if (e != cudaSuccess) return (int)e;                                                                                \
        gnm_count_launch(GNM_K_LINEAR_WGRAD_TC);                                                                            \
        e = gnm_launch_pdl<WgradTcParams>(linear_wgrad_tc_kernel<A, F>, grid, WG_THREADS, WG_SMEM, stream, p);              \
        if (e != cudaSuccess) return (int)e;                                                                                \
    } while (0)
    if (act && fast) GNM_WG_LAUNCH(true, true);
    else if (fast) GNM_WG_LAUNCH(false, true);
    else if (act) GNM_WG_LAUNCH(true, false);
    else GNM_WG_LAUNCH(false, false);
#undef GNM_WG_LAUNCH
    e = cudaGetLastError();
    return e == cudaSuccess ? GNM_OK : (int)e;
}

/* One-pass backward unit (see linear_bwd_onepass_tc_kernel). GNM_ERR_TOO_LARGE, nothing launched, unless the unit is 64 x 64,
 * every matrix is 16-byte aligned with a leading dimension divisible by four, and dx and dw are requested. */
int gnm_launch_linear_bwd_onepass_tc(const float* dy, int64_t lddy, const float* z, int64_t ldz, const float* coef,
                                     const float* x, int64_t ldx, const float* in_scale, const float* in_shift,
                                     const float* in_mean, const float* in_rstd, const float* w, int64_t ldw, float* dw,
                                     int64_t lddw, float* dbias, float* dx, int64_t lddx, double* stats_in, int n_rows,
                                     int n_out, int n_in, const gnm_bn_tail* tail, cudaStream_t stream) {
    if (n_in != BT_F || n_out != BT_F || dx == nullptr || dw == nullptr || n_rows < 1) return GNM_ERR_TOO_LARGE;
    if (((lddy | ldz | ldx | lddx) & 3) != 0 || !gnm_aligned16(dy) || !gnm_aligned16(z) || !gnm_aligned16(x) ||
        !gnm_aligned16(dx) || !gnm_aligned16(coef))
        return GNM_ERR_TOO_LARGE;
    const bool act = in_scale != nullptr;
    if (act && (!gnm_aligned16(in_scale) || !gnm_aligned16(in_shift))) return GNM_ERR_TOO_LARGE;
    if (tail != nullptr && (stats_in == nullptr || !act)) return GNM_ERR_BAD_ARG;
    int dev = 0, sms = 148, major = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) return GNM_ERR_TOO_LARGE;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    LinBwd1Params p;
    p.dy = dy; p.lddy = lddy; p.z = z; p.ldz = ldz; p.coef = coef; p.w = w; p.ldw = ldw; p.x = x; p.ldx = ldx;
    p.in_scale = in_scale; p.in_shift = in_shift; p.in_mean = in_mean; p.in_rstd = in_rstd; p.dx = dx; p.lddx = lddx;
    p.dw = dw; p.lddw = lddw; p.db = dbias; p.stats_in = stats_in; p.n_rows = n_rows;
    const int trc = bn_tail_args(tail, stats_in, n_in, &p.tail);
    if (trc != GNM_OK) return trc;
    const int tiles = (n_rows + 127) / 128;
    const int grid = tiles < sms ? tiles : sms;
    cudaError_t e;
#define GNM_OP_LAUNCH(A)                                                                                                    \
    do {                                                                                                                    \
        e = cudaFuncSetAttribute(linear_bwd_onepass_tc_kernel<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, OP_SMEM);    \
        if (e != cudaSuccess) return (int)e;                                                                                \
        gnm_count_launch(GNM_K_LINEAR_BWD_ONEPASS_TC);                                                                      \
        e = gnm_launch_pdl<LinBwd1Params>(linear_bwd_onepass_tc_kernel<A>, grid, OP_THREADS, OP_SMEM, stream, p);           \
        if (e != cudaSuccess) return (int)e;                                                                                \
    } while (0)
    if (act) GNM_OP_LAUNCH(true);
    else GNM_OP_LAUNCH(false);
#undef GNM_OP_LAUNCH
    e = cudaGetLastError();
    return e == cudaSuccess ? GNM_OK : (int)e;
}
