// Fused backward of one Linear -> BatchNorm (-> ReLU) unit (reference: autograd of models/mlp.py:48-49 and
// models/graphcnn.py:162-166). One pass over the rows replaces four kernels of the unfused path
// (bn_bwd_apply, linear for dX, linear_wgrad for dW, relu_bn_bwd_reduce of the previous unit):
//
//   dz[m,o]  = cA[o]*dy[m,o] + cB[o]*z[m,o] + cC[o]          BatchNorm backward, folded into the load
//   dW[o,i] += sum_m dz[m,o] * a[m,i]                         a = relu(x*in_scale + in_shift) or x
//   db[o]   += sum_m dz[m,o]
//   dx[m,i]  = (sum_o dz[m,o] * W[o,i]) * [a[m,i] > 0]         gradient wrt the unit's input (pre-ReLU-masked)
//   st[i]   += sum_m dx[m,i];  st[Fi+i] += sum_m dx[m,i] * (x[m,i]-mean_in[i])*rstd_in[i]
//
// cA/cB/cC come from gnm_bn_bwd_coeffs (training: cA = g*rstd, cB = -g*rstd^2*m2, cC = g*rstd*(rstd*m2*mean - m1)
// with m1 = sum(dy)/count, m2 = sum(dy*xhat)/count; eval: cA = g*rstd, cB = cC = 0).
// fp32 FFMA; F_out, F_in <= 64 (zero padded); persistent CTAs keep their dW tile in registers across row tiles.
#include "gnm_common.cuh"
#include "gnm_p2p.cuh"

namespace {

constexpr int FB_T = 64;            // feature tile (max F_out, F_in)
constexpr int FB_M = 64;            // rows per tile
constexpr int FB_PAD = 4;
constexpr int FB_LD = FB_T + FB_PAD;    // 68 floats: rows stay 16-byte aligned, float4 column reads conflict-free
constexpr int FB_SMEM_FLOATS = 5 * FB_M * FB_LD + 2 * 2 * FB_T;   // dz, dzT, a, x, W + stat scratch

struct LinBwdParams {
    const float* dy; int64_t lddy;
    const float* z; int64_t ldz;
    const float* coef;              // [3][n_out]
    const float* x; int64_t ldx;
    const float* in_scale; const float* in_shift; const float* in_mean; const float* in_rstd;
    const float* w; int64_t ldw;    // [n_out][n_in]
    float* dw; int64_t lddw; float* db;
    float* dx; int64_t lddx;
    double* stats_in;               // [2*n_in] (nullable)
    int n_rows, n_out, n_in;
};

__global__ void __launch_bounds__(256, 2) linear_bwd_kernel(const LinBwdParams p) {
    extern __shared__ __align__(16) float fb_smem[];
    float* s_dz = fb_smem;                       // [m][o]
    float* s_dzT = s_dz + FB_M * FB_LD;          // [o][m]
    float* s_a = s_dzT + FB_M * FB_LD;           // [m][i]  activated input
    float* s_x = s_a + FB_M * FB_LD;             // [m][i]  raw input (only with in_scale)
    float* s_w = s_x + FB_M * FB_LD;             // [o][i]
    float* s_st = s_w + FB_M * FB_LD;            // [2][FB_T] + [2][FB_T] scratch
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;      // 16 x 16 threads, 4 x 4 outputs each
    const bool act = p.in_scale != nullptr;
    const int n_tiles = (p.n_rows + FB_M - 1) / FB_M;

    // W tile (zero padded), once per CTA
    for (int e = tid; e < FB_T * FB_T; e += 256) {
        const int o = e >> 6, i = e & 63;
        s_w[o * FB_LD + i] = (o < p.n_out && i < p.n_in) ? p.w[(int64_t)o * p.ldw + i] : 0.f;
    }
    // per-thread column constants for the tile loads: this thread always loads columns 4*(tid&15)..+3
    const int c0 = (tid & 15) * 4;
    float cA[4], cB[4], cC[4], isc[4], ish[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int c = c0 + q;
        const bool vo = c < p.n_out, vi = c < p.n_in;
        cA[q] = vo ? p.coef[c] : 0.f;
        cB[q] = vo ? p.coef[p.n_out + c] : 0.f;
        cC[q] = vo ? p.coef[2 * p.n_out + c] : 0.f;
        isc[q] = (act && vi) ? p.in_scale[c] : 1.f;
        ish[q] = (act && vi) ? p.in_shift[c] : 0.f;
    }
    // constants for the dx epilogue: this thread's output columns 4*tx..+3
    float imean[4], irstd[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int c = tx * 4 + q;
        imean[q] = (act && c < p.n_in) ? p.in_mean[c] : 0.f;
        irstd[q] = (act && c < p.n_in) ? p.in_rstd[c] : 0.f;
    }
    float wacc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) wacc[a][b] = 0.f;
    float bacc[4] = {0.f, 0.f, 0.f, 0.f};
    float st1[4] = {0.f, 0.f, 0.f, 0.f}, st2[4] = {0.f, 0.f, 0.f, 0.f};
    const bool vec_in = (p.lddy % 4 == 0) && (p.ldz % 4 == 0) && (p.ldx % 4 == 0) && (p.n_out % 4 == 0) &&
                        (p.n_in % 4 == 0) && gnm_aligned16(p.dy) && gnm_aligned16(p.z) && gnm_aligned16(p.x);
    const bool vec_out = p.dx != nullptr && (p.lddx % 4 == 0) && (p.n_in % 4 == 0) && gnm_aligned16(p.dx);

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m0 = tile * FB_M;
        __syncthreads();                          // previous tile's readers are done (also orders the W fill)
        // ---- load: 64 rows x 16 float4 per operand, 4 per thread
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int mm = (tid >> 4) + 16 * j;
            const int m = m0 + mm;
            float dzv[4] = {0.f, 0.f, 0.f, 0.f}, xv[4] = {0.f, 0.f, 0.f, 0.f}, av[4] = {0.f, 0.f, 0.f, 0.f};
            if (m < p.n_rows) {
                float dyv[4], zv[4];
                if (vec_in) {
                    const float4 t0 = (c0 < p.n_out) ? ld_stream_f4(p.dy + (int64_t)m * p.lddy + c0) : make_float4(0, 0, 0, 0);
                    const float4 t1 = (c0 < p.n_out) ? ld_stream_f4(p.z + (int64_t)m * p.ldz + c0) : make_float4(0, 0, 0, 0);
                    const float4 t2 = (c0 < p.n_in) ? __ldg(reinterpret_cast<const float4*>(p.x + (int64_t)m * p.ldx + c0))
                                                    : make_float4(0, 0, 0, 0);
                    dyv[0] = t0.x; dyv[1] = t0.y; dyv[2] = t0.z; dyv[3] = t0.w;
                    zv[0] = t1.x; zv[1] = t1.y; zv[2] = t1.z; zv[3] = t1.w;
                    xv[0] = t2.x; xv[1] = t2.y; xv[2] = t2.z; xv[3] = t2.w;
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int c = c0 + q;
                        dyv[q] = (c < p.n_out) ? p.dy[(int64_t)m * p.lddy + c] : 0.f;
                        zv[q] = (c < p.n_out) ? p.z[(int64_t)m * p.ldz + c] : 0.f;
                        xv[q] = (c < p.n_in) ? p.x[(int64_t)m * p.ldx + c] : 0.f;
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    dzv[q] = (c0 + q < p.n_out) ? fmaf(cA[q], dyv[q], fmaf(cB[q], zv[q], cC[q])) : 0.f;
                    av[q] = (c0 + q < p.n_in) ? (act ? fmaxf(fmaf(xv[q], isc[q], ish[q]), 0.f) : xv[q]) : 0.f;
                }
            }
            *reinterpret_cast<float4*>(&s_dz[mm * FB_LD + c0]) = make_float4(dzv[0], dzv[1], dzv[2], dzv[3]);
            *reinterpret_cast<float4*>(&s_a[mm * FB_LD + c0]) = make_float4(av[0], av[1], av[2], av[3]);
            if (act) *reinterpret_cast<float4*>(&s_x[mm * FB_LD + c0]) = make_float4(xv[0], xv[1], xv[2], xv[3]);
#pragma unroll
            for (int q = 0; q < 4; ++q) s_dzT[(c0 + q) * FB_LD + mm] = dzv[q];
        }
        __syncthreads();
        // ---- GEMM 2: dW[o = 4ty.., i = 4tx..] += sum_m dz[m][o] * a[m][i];  db
#pragma unroll 8
        for (int m = 0; m < FB_M; ++m) {
            const float4 dv = *reinterpret_cast<const float4*>(&s_dz[m * FB_LD + ty * 4]);
            const float4 av = *reinterpret_cast<const float4*>(&s_a[m * FB_LD + tx * 4]);
            const float d[4] = {dv.x, dv.y, dv.z, dv.w};
            const float a[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
            for (int r = 0; r < 4; ++r) {
#pragma unroll
                for (int c = 0; c < 4; ++c) wacc[r][c] = fmaf(d[r], a[c], wacc[r][c]);
            }
            if (tx == 0) {
#pragma unroll
                for (int r = 0; r < 4; ++r) bacc[r] += d[r];
            }
        }
        // ---- GEMM 1: dx[m = 4ty.., i = 4tx..] = sum_o dz[m][o] * W[o][i]
        if (p.dx != nullptr || p.stats_in != nullptr) {
            float xacc[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) xacc[r][c] = 0.f;
#pragma unroll 8
            for (int o = 0; o < FB_T; ++o) {
                const float4 dv = *reinterpret_cast<const float4*>(&s_dzT[o * FB_LD + ty * 4]);
                const float4 wv = *reinterpret_cast<const float4*>(&s_w[o * FB_LD + tx * 4]);
                const float d[4] = {dv.x, dv.y, dv.z, dv.w};
                const float w[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) xacc[r][c] = fmaf(d[r], w[c], xacc[r][c]);
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int mm = ty * 4 + r;
                const int m = m0 + mm;
                if (m >= p.n_rows) continue;
                float o4[4];
                if (act) {
                    const float4 av = *reinterpret_cast<const float4*>(&s_a[mm * FB_LD + tx * 4]);
                    const float4 xv = *reinterpret_cast<const float4*>(&s_x[mm * FB_LD + tx * 4]);
                    const float a[4] = {av.x, av.y, av.z, av.w};
                    const float x[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float v = (a[c] > 0.f) ? xacc[r][c] : 0.f;
                        o4[c] = v;
                        st1[c] += v;
                        st2[c] = fmaf(v, (x[c] - imean[c]) * irstd[c], st2[c]);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 4; ++c) o4[c] = xacc[r][c];
                }
                if (p.dx != nullptr) {
                    float* out = p.dx + (int64_t)m * p.lddx + tx * 4;
                    if (vec_out && tx * 4 + 3 < p.n_in) {
                        *reinterpret_cast<float4*>(out) = make_float4(o4[0], o4[1], o4[2], o4[3]);
                    } else {
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if (tx * 4 + c < p.n_in) out[c] = o4[c];
                    }
                }
            }
        }
    }
    // ---- flush: dW / db with fp32 atomics (one partial per CTA), BN-backward stats with fp64 atomics
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int o = ty * 4 + r;
        if (o >= p.n_out) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int i = tx * 4 + c;
            if (i < p.n_in) atomicAdd(&p.dw[(int64_t)o * p.lddw + i], wacc[r][c]);
        }
        if (tx == 0 && p.db != nullptr) atomicAdd(&p.db[o], bacc[r]);
    }
    if (p.stats_in != nullptr && act) {
        __syncthreads();
        for (int e = tid; e < 2 * FB_T; e += 256) s_st[e] = 0.f;
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            atomicAdd(&s_st[tx * 4 + c], st1[c]);
            atomicAdd(&s_st[FB_T + tx * 4 + c], st2[c]);
        }
        __syncthreads();
        if (tid < 2 * FB_T) {
            const int which = tid / FB_T, c = tid % FB_T;
            if (c < p.n_in) atomicAdd(&p.stats_in[which * p.n_in + c], (double)s_st[tid]);
        }
    }
}

__global__ void bn_bwd_coeffs_kernel(double* stats, double count, const float* __restrict__ gamma,
                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                     float* __restrict__ coef, int n_feat, const P2PArgs comm) {
    // data parallel: [sum dy, sum dy*xhat] of all ranks, exchanged over peer memory right here (single-CTA launches only)
    if (comm.world > 1 && stats != nullptr) p2p_allreduce_block(stats, 2 * n_feat, comm);
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_feat) return;
    const float g = (gamma ? gamma[c] : 1.f) * rstd[c];
    float a = g, b = 0.f, k = 0.f;
    if (stats != nullptr) {
        const float m1 = (float)(stats[c] / count), m2 = (float)(stats[n_feat + c] / count);
        b = -g * rstd[c] * m2;
        k = g * (rstd[c] * m2 * mean[c] - m1);
    }
    coef[c] = a;
    coef[n_feat + c] = b;
    coef[2 * n_feat + c] = k;
}

}  // namespace

int gnm_p2p_abort_flag_mlp_bwd(int* aborted) {
    int v = 0, zero = 0;
    cudaError_t e = cudaMemcpyFromSymbol(&v, g_p2p_abort, sizeof(int));
    if (e != cudaSuccess) return (int)e;
    e = cudaMemcpyToSymbol(g_p2p_abort, &zero, sizeof(int));
    if (aborted) *aborted |= v;
    return e == cudaSuccess ? GNM_OK : (int)e;
}

extern "C" int gnm_bn_bwd_coeffs(double* stats, double count, const float* gamma, const float* mean,
                                 const float* rstd, float* coef, int n_feat, const gnm_p2p_comm* comm,
                                 gnm_stream_t stream) {
    if (n_feat < 0 || (stats != nullptr && count <= 0.0)) return GNM_ERR_BAD_ARG;
    if (n_feat == 0) return GNM_OK;
    if (!rstd || !coef || (stats != nullptr && !mean)) return GNM_ERR_BAD_ARG;
    P2PArgs pa;
    const int prc = p2p_args(comm, 2 * n_feat, &pa);
    if (prc < 0) return prc;
    gnm_count_launch(GNM_K_OTHER);
    bn_bwd_coeffs_kernel<<<(n_feat + 127) / 128, 128, 0, gnm_cast_stream(stream)>>>(stats, count, gamma, mean, rstd,
                                                                                    coef, n_feat, pa);
    GNM_RETURN_IF_LAUNCH_FAILED();
    return GNM_OK;
}

// tcgen05 implementation (gnm_linear_bwd_tc.cu): input-gradient kernel and weight-gradient kernel
int gnm_launch_linear_bwd_dx_tc(const float* dy, int64_t lddy, const float* z, int64_t ldz, const float* coef,
                                const float* x, int64_t ldx, const float* in_scale, const float* in_shift,
                                const float* in_mean, const float* in_rstd, const float* w, int64_t ldw, float* dx,
                                int64_t lddx, double* stats_in, int n_rows, int n_out, int n_in, const gnm_bn_tail* tail,
                                cudaStream_t stream);
int gnm_launch_linear_wgrad_tc(const float* dy, int64_t lddy, const float* z, int64_t ldz, const float* coef,
                               const float* x, int64_t ldx, const float* in_scale, const float* in_shift, float* dw,
                               int64_t lddw, float* dbias, int n_rows, int n_out, int n_in, cudaStream_t stream);
int gnm_launch_linear_bwd_onepass_tc(const float* dy, int64_t lddy, const float* z, int64_t ldz, const float* coef,
                                     const float* x, int64_t ldx, const float* in_scale, const float* in_shift,
                                     const float* in_mean, const float* in_rstd, const float* w, int64_t ldw, float* dw,
                                     int64_t lddw, float* dbias, float* dx, int64_t lddx, double* stats_in, int n_rows,
                                     int n_out, int n_in, const gnm_bn_tail* tail, cudaStream_t stream);
int gnm_linear_impl_value();     // gnm_mlp.cu: 0 auto, 1 FFMA only, 2 tcgen05 only, 3 tcgen05 two-pass pair only

extern "C" int gnm_linear_bwd(const float* dy, int64_t lddy, const float* z, int64_t ldz, const float* coef,
                              const float* x, int64_t ldx, const float* in_scale, const float* in_shift,
                              const float* in_mean, const float* in_rstd, const float* w, int64_t ldw, float* dw,
                              int64_t lddw, float* dbias, float* dx, int64_t lddx, double* stats_in, int n_rows,
                              int n_out, int n_in, const gnm_bn_tail* tail, gnm_stream_t stream) {
    if (n_rows < 0 || n_out <= 0 || n_in <= 0) return GNM_ERR_BAD_ARG;
    if (tail != nullptr && (dx == nullptr || stats_in == nullptr)) return GNM_ERR_BAD_ARG;
    if (n_out > FB_T || n_in > FB_T) return GNM_ERR_TOO_LARGE;
    if (n_rows == 0) return GNM_OK;
    if (!dy || !z || !coef || !x || !w || !dw) return GNM_ERR_BAD_ARG;
    if (in_scale != nullptr && (!in_shift || !in_mean || !in_rstd)) return GNM_ERR_BAD_ARG;
    if (stats_in != nullptr && in_scale == nullptr) return GNM_ERR_BAD_ARG;
    LinBwdParams p;
    p.dy = dy; p.lddy = lddy; p.z = z; p.ldz = ldz; p.coef = coef; p.x = x; p.ldx = ldx;
    p.in_scale = in_scale; p.in_shift = in_shift; p.in_mean = in_mean; p.in_rstd = in_rstd;
    p.w = w; p.ldw = ldw; p.dw = dw; p.lddw = lddw; p.db = dbias; p.dx = dx; p.lddx = lddx; p.stats_in = stats_in;
    p.n_rows = n_rows; p.n_out = n_out; p.n_in = n_in;
    const int impl = gnm_linear_impl_value();
    if (impl != 1 && (impl >= 2 || n_rows >= 4096)) {
        // one pass over dy, z, x when the unit is 64 x 64 and aligned ...
        int rc = GNM_ERR_TOO_LARGE;
        if (impl != 3 && dx != nullptr)
            rc = gnm_launch_linear_bwd_onepass_tc(dy, lddy, z, ldz, coef, x, ldx, in_scale, in_shift, in_mean, in_rstd, w, ldw,
                                                  dw, lddw, dbias, dx, lddx, stats_in, n_rows, n_out, n_in, tail,
                                                  gnm_cast_stream(stream));
        if (rc != GNM_ERR_TOO_LARGE) return rc;
        // ... else two tensor-core passes over the rows (input gradient, then weight gradient)
        rc = GNM_OK;
        if (dx != nullptr)
            rc = gnm_launch_linear_bwd_dx_tc(dy, lddy, z, ldz, coef, x, ldx, in_scale, in_shift, in_mean, in_rstd, w, ldw,
                                             dx, lddx, stats_in, n_rows, n_out, n_in, tail, gnm_cast_stream(stream));
        if (rc == GNM_OK)
            rc = gnm_launch_linear_wgrad_tc(dy, lddy, z, ldz, coef, x, ldx, in_scale, in_shift, dw, lddw, dbias, n_rows,
                                            n_out, n_in, gnm_cast_stream(stream));
        if (rc == GNM_OK || impl >= 2 || rc != GNM_ERR_TOO_LARGE) return rc;
    }
    if (tail != nullptr) return GNM_ERR_TOO_LARGE;      // the fused FFMA kernel has no tail: nothing launched
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t smem = (size_t)FB_SMEM_FLOATS * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(linear_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int tiles = (n_rows + FB_M - 1) / FB_M;
    int grid = sms * 2;
    if (grid > tiles) grid = tiles;
    gnm_count_launch(GNM_K_LINEAR_BWD_FFMA);
    linear_bwd_kernel<<<grid, 256, smem, gnm_cast_stream(stream)>>>(p);
    GNM_RETURN_IF_LAUNCH_FAILED();
    return GNM_OK;
}
