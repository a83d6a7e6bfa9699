// MLP pieces: Linear (+fused BatchNorm-apply/ReLU prologue, +fused batch-statistics epilogue),
// weight gradient, BatchNorm statistics / finalize / backward.
// Reference: models/mlp.py:40-49, models/graphcnn.py:162-166,185-190 (nn.Linear, nn.BatchNorm1d, ReLU).
#include "gnm_common.cuh"
#include "gnm_p2p.cuh"
#include "gnm_bn_tail.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// y[m, n] = sum_k f(x[m, k]) * W(k, n) + bias[n]      fp32 FFMA, 128 x 64 x 16 tiles
// ------------------------------------------------------------------------------------------
constexpr int LBM = 128, LBN = 64, LBK = 16, LTHREADS = 256;

__global__ void __launch_bounds__(LTHREADS)
linear_kernel(const float* __restrict__ x, int64_t ldx, int n_rows, int n_in, const float* __restrict__ w,
              int64_t ldw, int w_is_kn, const float* __restrict__ bias, const float* __restrict__ in_scale,
              const float* __restrict__ in_shift, float* __restrict__ y, int64_t ldy, int n_out,
              double* __restrict__ col_stats) {
    __shared__ __align__(16) float xs[LBK][LBM + 4];
    __shared__ __align__(16) float ws[LBK][LBN + 4];
    __shared__ float cs[2][LBN];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;          // 16 x 16 threads; thread tile 8 rows x 4 cols
    const int m0 = blockIdx.x * LBM, n0 = blockIdx.y * LBN;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < n_in; k0 += LBK) {
        // x tile: 128 rows x 16 k; thread loads k = tid & 15, rows (tid >> 4) + 16*i
        {
            const int kk = tid & 15;
            const int k = k0 + kk;
            float sc = 1.f, sh = 0.f;
            const bool pro = in_scale != nullptr;
            if (pro && k < n_in) { sc = in_scale[k]; sh = in_shift[k]; }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int mm = (tid >> 4) + 16 * i;
                const int m = m0 + mm;
                float v = 0.f;
                if (m < n_rows && k < n_in) {
                    v = x[(int64_t)m * ldx + k];
                    if (pro) v = fmaxf(fmaf(v, sc, sh), 0.f);
                }
                xs[kk][mm] = v;
            }
        }
        // w tile: 16 k x 64 n
        {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int e = tid + i * LTHREADS;    // 0..1023
                int kk, nn;
                if (w_is_kn) { nn = e & 63; kk = e >> 6; } else { kk = e & 15; nn = e >> 4; }
                const int k = k0 + kk, n = n0 + nn;
                float v = 0.f;
                if (k < n_in && n < n_out) v = w_is_kn ? w[(int64_t)k * ldw + n] : w[(int64_t)n * ldw + k];
                ws[kk][nn] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < LBK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&xs[kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&xs[kk][ty * 8 + 4]);
            const float4 b = *reinterpret_cast<const float4*>(&ws[kk][tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }

    if (col_stats != nullptr) {
        for (int i = tid; i < 2 * LBN; i += LTHREADS) (&cs[0][0])[i] = 0.f;
        __syncthreads();
    }
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx * 4 + j;
        const float bv = (bias != nullptr && n < n_out) ? bias[n] : 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int m = m0 + ty * 8 + i;
            const float v = acc[i][j] + bv;
            acc[i][j] = v;
            if (m < n_rows && n < n_out) { s1[j] += v; s2[j] = fmaf(v, v, s2[j]); }
        }
    }
    const bool vec_ok = (ldy % 4 == 0) && gnm_aligned16(y) && (n0 + tx * 4 + 3 < n_out);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 8 + i;
        if (m >= n_rows) continue;
        float* yr = y + (int64_t)m * ldy + n0 + tx * 4;
        if (vec_ok) {
            *reinterpret_cast<float4*>(yr) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n0 + tx * 4 + j < n_out) yr[j] = acc[i][j];
        }
    }
    if (col_stats != nullptr) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            atomicAdd(&cs[0][tx * 4 + j], s1[j]);
            atomicAdd(&cs[1][tx * 4 + j], s2[j]);
        }
        __syncthreads();
        if (tid < 2 * LBN) {
            const int which = tid / LBN, c = tid % LBN;
            const int n = n0 + c;
            if (n < n_out) atomicAdd(&col_stats[which * n_out + n], (double)cs[which][c]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// dw[o, i] += sum_m dz[m, o] * f(x[m, i]);  dbias[o] += sum_m dz[m, o]
// 64 x 64 output tile per CTA, rows split over blockIdx.z slabs, fp32 atomics to merge slabs.
// ------------------------------------------------------------------------------------------
constexpr int WBO = 64, WBI = 64, WBR = 16;

__global__ void __launch_bounds__(256)
linear_wgrad_kernel(const float* __restrict__ dz, int64_t lddz, const float* __restrict__ x, int64_t ldx, int n_rows,
                    int n_out, int n_in, const float* __restrict__ in_scale, const float* __restrict__ in_shift,
                    float* __restrict__ dw, int64_t lddw, float* __restrict__ dbias, int rows_per_slab) {
    __shared__ __align__(16) float zs[WBR][WBO + 4];
    __shared__ __align__(16) float xs[WBR][WBI + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;     // thread tile: o = ty*4.., i = tx*4..
    const int o0 = blockIdx.x * WBO, i0 = blockIdx.y * WBI;
    const int r0 = blockIdx.z * rows_per_slab, r1 = min(n_rows, r0 + rows_per_slab);
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    float bsum = 0.f;   // threads 0..63 accumulate dbias for column o0 + tid (only blockIdx.y == 0)
    const bool pro = in_scale != nullptr;
    for (int rb = r0; rb < r1; rb += WBR) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int e = tid + t * 256;       // 0..1023 = 16 rows x 64 cols
            const int rr = e >> 6, c = e & 63;
            const int r = rb + rr;
            float vz = 0.f, vx = 0.f;
            if (r < r1) {
                if (o0 + c < n_out) vz = dz[(int64_t)r * lddz + o0 + c];
                if (i0 + c < n_in) {
                    vx = x[(int64_t)r * ldx + i0 + c];
                    if (pro) vx = fmaxf(fmaf(vx, in_scale[i0 + c], in_shift[i0 + c]), 0.f);
                }
            }
            zs[rr][c] = vz;
            xs[rr][c] = vx;
        }
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < WBR; ++rr) {
            const float4 a = *reinterpret_cast<const float4*>(&zs[rr][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&xs[rr][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(av[p], bv[q], acc[p][q]);
        }
        if (dbias != nullptr && blockIdx.y == 0 && tid < WBO) {
#pragma unroll
            for (int rr = 0; rr < WBR; ++rr) bsum += zs[rr][tid];
        }
        __syncthreads();
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int o = o0 + ty * 4 + p;
        if (o >= n_out) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = i0 + tx * 4 + q;
            if (i < n_in) atomicAdd(&dw[(int64_t)o * lddw + i], acc[p][q]);
        }
    }
    if (dbias != nullptr && blockIdx.y == 0 && tid < WBO && o0 + tid < n_out) atomicAdd(&dbias[o0 + tid], bsum);
}

// ------------------------------------------------------------------------------------------
// column statistics
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
col_stats_kernel(const float* __restrict__ x, int64_t ldx, int n_rows, int n_feat, double* __restrict__ col_stats,
                 int rows_per_cta) {
    // thread -> column (tid % cw), row lane (tid / cw); cw = min(n_feat chunk, 256)
    const int f0 = blockIdx.y * 256;
    const int cw = min(256, n_feat - f0);
    const int rl = 256 / cw;
    const int c = threadIdx.x % cw, rr = threadIdx.x / cw;
    __shared__ float sh1[256], sh2[256];
    float s1 = 0.f, s2 = 0.f;
    const int r0 = blockIdx.x * rows_per_cta, r1 = min(n_rows, r0 + rows_per_cta);
    if (rr < rl) {
        for (int r = r0 + rr; r < r1; r += rl) {
            const float v = x[(int64_t)r * ldx + f0 + c];
            s1 += v;
            s2 = fmaf(v, v, s2);
        }
    }
    sh1[threadIdx.x] = s1;
    sh2[threadIdx.x] = s2;
    __syncthreads();
    if (threadIdx.x < cw) {
        double a = 0.0, b = 0.0;
        for (int k = 0; k < rl; ++k) { a += (double)sh1[k * cw + threadIdx.x]; b += (double)sh2[k * cw + threadIdx.x]; }
        atomicAdd(&col_stats[f0 + threadIdx.x], a);
        atomicAdd(&col_stats[n_feat + f0 + threadIdx.x], b);
    }
}

__global__ void bn_finalize_kernel(double* col_stats, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   int64_t* __restrict__ nbt, float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ mean_o, float* __restrict__ rstd_o, int n_feat, const P2PArgs comm) {
    // data parallel: the sums of all ranks are exchanged over peer memory right here (single-CTA launches only)
    if (comm.world > 1) p2p_allreduce_block(col_stats, 2 * n_feat, comm);
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && nbt != nullptr) *nbt += 1;
    if (c >= n_feat) return;
    const double mean = col_stats[c] / count;
    double var = col_stats[n_feat + c] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    const float sc = g * rstd;
    scale[c] = sc;
    shift[c] = b - (float)mean * sc;
    mean_o[c] = (float)mean;
    rstd_o[c] = rstd;
    if (running_mean != nullptr) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    if (running_var != nullptr) {
        const double unb = count > 1.0 ? var * (count / (count - 1.0)) : var;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
    }
}

__global__ void bn_eval_affine_kernel(const float* __restrict__ rm, const float* __restrict__ rv,
                                      const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                      float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_o,
                                      float* __restrict__ rstd_o, int n_feat) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_feat) return;
    const float rstd = 1.f / sqrtf(rv[c] + eps);
    const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    const float sc = g * rstd;
    scale[c] = sc;
    shift[c] = b - rm[c] * sc;
    mean_o[c] = rm[c];
    rstd_o[c] = rstd;
}

// ------------------------------------------------------------------------------------------
// h = relu(z*scale + shift); pooled[g] = pool_scale[g] * sum_rows h        (one CTA per graph x column chunk)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bn_relu_readout_kernel(const float* __restrict__ z, int64_t ldz, int n_feat, const float* __restrict__ scale,
                       const float* __restrict__ shift, float* __restrict__ h, int64_t ldh,
                       const int32_t* __restrict__ node_off, const float* __restrict__ pool_scale,
                       float* __restrict__ pooled, int64_t ld_pooled) {
    const int g = blockIdx.x;
    const int f0 = blockIdx.y * 256;
    const int cw = min(256, n_feat - f0);
    const int rl = 256 / cw;
    const int c = threadIdx.x % cw, rr = threadIdx.x / cw;
    __shared__ float red[256];
    const int r0 = node_off[g], r1 = node_off[g + 1];
    float s = 0.f;
    if (rr < rl) {
        const float sc = scale[f0 + c], sh = shift[f0 + c];
        for (int r = r0 + rr; r < r1; r += rl) {
            const float v = fmaxf(fmaf(z[(int64_t)r * ldz + f0 + c], sc, sh), 0.f);
            if (h != nullptr) h[(int64_t)r * ldh + f0 + c] = v;
            s += v;
        }
    }
    red[threadIdx.x] = s;
    __syncthreads();
    if (pooled != nullptr && threadIdx.x < cw) {
        float t = 0.f;
        for (int k = 0; k < rl; ++k) t += red[k * cw + threadIdx.x];
        if (pool_scale != nullptr) t *= pool_scale[g];
        pooled[(int64_t)g * ld_pooled + f0 + threadIdx.x] = t;
    }
}

// 128-bit version (F % 4 == 0, aligned): F/4 lanes cover a row, 256/(F/4) rows per pass, two passes in flight.
__global__ void __launch_bounds__(256)
bn_relu_readout_vec_kernel(const float* __restrict__ z, int64_t ldz, int n_feat, const float* __restrict__ scale,
                           const float* __restrict__ shift, float* __restrict__ h, int64_t ldh,
                           const int32_t* __restrict__ node_off, const float* __restrict__ pool_scale,
                           float* __restrict__ pooled, int64_t ld_pooled) {
    const int g = blockIdx.x;
    const int lpr = n_feat >> 2;                 // lanes per row (<= 256)
    const int rpp = 256 / lpr;                   // rows per pass
    const int sub = threadIdx.x % lpr, rr = threadIdx.x / lpr;
    __shared__ float4 red[256];
    const int r0 = node_off[g], r1 = node_off[g + 1];
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rr < rpp) {
        const float4 sc = *reinterpret_cast<const float4*>(scale + sub * 4);
        const float4 sh = *reinterpret_cast<const float4*>(shift + sub * 4);
        int r = r0 + rr;
        for (; r + rpp < r1; r += 2 * rpp) {
            const float4 a = ld_stream_f4(z + (int64_t)r * ldz + sub * 4);
            const float4 b = ld_stream_f4(z + (int64_t)(r + rpp) * ldz + sub * 4);
            float4 va, vb;
            va.x = fmaxf(fmaf(a.x, sc.x, sh.x), 0.f); va.y = fmaxf(fmaf(a.y, sc.y, sh.y), 0.f);
            va.z = fmaxf(fmaf(a.z, sc.z, sh.z), 0.f); va.w = fmaxf(fmaf(a.w, sc.w, sh.w), 0.f);
            vb.x = fmaxf(fmaf(b.x, sc.x, sh.x), 0.f); vb.y = fmaxf(fmaf(b.y, sc.y, sh.y), 0.f);
            vb.z = fmaxf(fmaf(b.z, sc.z, sh.z), 0.f); vb.w = fmaxf(fmaf(b.w, sc.w, sh.w), 0.f);
            if (h != nullptr) {
                *reinterpret_cast<float4*>(h + (int64_t)r * ldh + sub * 4) = va;
                *reinterpret_cast<float4*>(h + (int64_t)(r + rpp) * ldh + sub * 4) = vb;
            }
            s.x += va.x + vb.x; s.y += va.y + vb.y; s.z += va.z + vb.z; s.w += va.w + vb.w;
        }
        for (; r < r1; r += rpp) {
            const float4 a = ld_stream_f4(z + (int64_t)r * ldz + sub * 4);
            float4 va;
            va.x = fmaxf(fmaf(a.x, sc.x, sh.x), 0.f); va.y = fmaxf(fmaf(a.y, sc.y, sh.y), 0.f);
            va.z = fmaxf(fmaf(a.z, sc.z, sh.z), 0.f); va.w = fmaxf(fmaf(a.w, sc.w, sh.w), 0.f);
            if (h != nullptr) *reinterpret_cast<float4*>(h + (int64_t)r * ldh + sub * 4) = va;
            s.x += va.x; s.y += va.y; s.z += va.z; s.w += va.w;
        }
    }
    red[threadIdx.x] = s;
    __syncthreads();
    if (pooled != nullptr && threadIdx.x < lpr) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < rpp; ++k) {
            const float4 v = red[k * lpr + threadIdx.x];
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        const float ps = pool_scale ? pool_scale[g] : 1.f;
        float* o = pooled + (int64_t)g * ld_pooled + threadIdx.x * 4;
        o[0] = t.x * ps; o[1] = t.y * ps; o[2] = t.z * ps; o[3] = t.w * ps;
    }
}

// ------------------------------------------------------------------------------------------
// backward of relu(bn(z)), pass 1: assemble the incoming gradient, mask, reduce
// ------------------------------------------------------------------------------------------
// DOUT: a per-row incoming gradient d_out exists (the unfused fallback); without it (the top layer of the training step) the
// kernel keeps fewer rows' operands in registers and three CTAs fit an SM
template <bool DOUT>
__global__ void __launch_bounds__(256, DOUT ? 2 : 3)
relu_bn_bwd_reduce_vec_kernel(const float* __restrict__ z, int64_t ldz, int n_feat, const float* __restrict__ scale,
                              const float* __restrict__ shift, const float* __restrict__ mean,
                              const float* __restrict__ rstd, const float* __restrict__ d_out, int64_t ld_dout,
                              const float* __restrict__ d_pooled, int64_t ld_dpooled,
                              const float* __restrict__ pool_scale, const float* __restrict__ d_score,
                              const float* __restrict__ u, int64_t ldu, const float* __restrict__ d_neg,
                              int64_t ld_dneg, int n_neg, const int32_t* __restrict__ node_off, int n_graphs,
                              float* __restrict__ dy, int64_t lddy, double* __restrict__ stats, const BnTailDev tail) {
    // persistent over graphs: a thread owns the same four columns for every graph, so the column sums stay in
    // registers and each CTA issues ONE set of fp64 atomics (same-address atomics of a CTA per graph serialised in L2)
    const int lpr = n_feat >> 2;
    const int rpp = 256 / lpr;
    const int sub = threadIdx.x % lpr, rr = threadIdx.x / lpr;
    __shared__ float4 red1[256], red2[256];
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
    const int f = sub * 4;
    float4 sc = s1, sh = s1, mu = s1, rs = s1;
    if (rr < rpp) {
        sc = *reinterpret_cast<const float4*>(scale + f); sh = *reinterpret_cast<const float4*>(shift + f);
        mu = *reinterpret_cast<const float4*>(mean + f); rs = *reinterpret_cast<const float4*>(rstd + f);
    }
    for (int g = blockIdx.x; g < n_graphs && rr < rpp; g += gridDim.x) {
        const int r0 = node_off[g], r1 = node_off[g + 1];
        float4 gp = make_float4(0.f, 0.f, 0.f, 0.f), uu = gp;
        if (d_pooled != nullptr) {
            const float ps = pool_scale ? pool_scale[g] : 1.f;
            const float* q = d_pooled + (int64_t)g * ld_dpooled + f;
            gp = make_float4(q[0] * ps, q[1] * ps, q[2] * ps, q[3] * ps);
        }
        if (d_score != nullptr) {
            const float* q = u + (int64_t)g * ldu + f;
            uu = make_float4(q[0], q[1], q[2], q[3]);
        }
        // four rows per thread and iteration, every load issued before the first use (the rows of the shuffled-row gradient
        // d_neg are the first n_neg rows of the batch: loaded one by one they made the CTAs of the first graphs the tail
        // of the launch)
        constexpr int UB = 4;
        for (int rb = r0 + rr; rb < r1; rb += UB * rpp) {
            float4 zv[UB], dv[UB], nv[UB];
            float ds[UB];
#pragma unroll
            for (int k = 0; k < UB; ++k) {
                const int r = rb + k * rpp;
                const bool ok = r < r1;
                zv[k] = nv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (DOUT) dv[k] = zv[k];
                ds[k] = 0.f;
                if (ok) {
                    zv[k] = ld_stream_f4(z + (int64_t)r * ldz + f);
                    if (DOUT) dv[k] = ld_stream_f4(d_out + (int64_t)r * ld_dout + f);
                    if (d_score != nullptr) ds[k] = __ldg(d_score + r);
                    if (d_neg != nullptr && r < n_neg) nv[k] = *reinterpret_cast<const float4*>(d_neg + (int64_t)r * ld_dneg + f);
                }
            }
#pragma unroll
            for (int k = 0; k < UB; ++k) {
                const int r = rb + k * rpp;
                if (r >= r1) continue;
                float4 gr = gp;
                if (DOUT) { gr.x += dv[k].x; gr.y += dv[k].y; gr.z += dv[k].z; gr.w += dv[k].w; }
                gr.x = fmaf(ds[k], uu.x, gr.x); gr.y = fmaf(ds[k], uu.y, gr.y);
                gr.z = fmaf(ds[k], uu.z, gr.z); gr.w = fmaf(ds[k], uu.w, gr.w);
                gr.x += nv[k].x; gr.y += nv[k].y; gr.z += nv[k].z; gr.w += nv[k].w;
                float4 v;
                v.x = (fmaf(zv[k].x, sc.x, sh.x) > 0.f) ? gr.x : 0.f;
                v.y = (fmaf(zv[k].y, sc.y, sh.y) > 0.f) ? gr.y : 0.f;
                v.z = (fmaf(zv[k].z, sc.z, sh.z) > 0.f) ? gr.z : 0.f;
                v.w = (fmaf(zv[k].w, sc.w, sh.w) > 0.f) ? gr.w : 0.f;
                *reinterpret_cast<float4*>(dy + (int64_t)r * lddy + f) = v;
                s1.x += v.x; s1.y += v.y; s1.z += v.z; s1.w += v.w;
                s2.x = fmaf(v.x, (zv[k].x - mu.x) * rs.x, s2.x); s2.y = fmaf(v.y, (zv[k].y - mu.y) * rs.y, s2.y);
                s2.z = fmaf(v.z, (zv[k].z - mu.z) * rs.z, s2.z); s2.w = fmaf(v.w, (zv[k].w - mu.w) * rs.w, s2.w);
            }
        }
    }
    red1[threadIdx.x] = s1;
    red2[threadIdx.x] = s2;
    __syncthreads();
    if (stats != nullptr && threadIdx.x < lpr) {
        double a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
        for (int k = 0; k < rpp; ++k) {
            const float4 v = red1[k * lpr + threadIdx.x], w = red2[k * lpr + threadIdx.x];
            a[0] += v.x; a[1] += v.y; a[2] += v.z; a[3] += v.w;
            b[0] += w.x; b[1] += w.y; b[2] += w.z; b[3] += w.w;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            atomicAdd(&stats[threadIdx.x * 4 + q], a[q]);
            atomicAdd(&stats[n_feat + threadIdx.x * 4 + q], b[q]);
        }
    }
    bn_tail_run(tail);
}

__global__ void __launch_bounds__(256)
relu_bn_bwd_reduce_kernel(const float* __restrict__ z, int64_t ldz, int n_feat, const float* __restrict__ scale,
                          const float* __restrict__ shift, const float* __restrict__ mean,
                          const float* __restrict__ rstd, const float* __restrict__ d_out, int64_t ld_dout,
                          const float* __restrict__ d_pooled, int64_t ld_dpooled, const float* __restrict__ pool_scale,
                          const float* __restrict__ d_score, const float* __restrict__ u, int64_t ldu,
                          const float* __restrict__ d_neg, int64_t ld_dneg, int n_neg,
                          const int32_t* __restrict__ node_off, float* __restrict__ dy, int64_t lddy,
                          double* __restrict__ stats) {
    const int g = blockIdx.x;
    const int f0 = blockIdx.y * 256;
    const int cw = min(256, n_feat - f0);
    const int rl = 256 / cw;
    const int c = threadIdx.x % cw, rr = threadIdx.x / cw;
    __shared__ float sh1[256], sh2[256];
    const int r0 = node_off[g], r1 = node_off[g + 1];
    float s1 = 0.f, s2 = 0.f;
    if (rr < rl) {
        const int f = f0 + c;
        const float sc = scale[f], sh = shift[f], mu = mean[f], rs = rstd[f];
        float gp = 0.f, uu = 0.f;
        if (d_pooled != nullptr) gp = d_pooled[(int64_t)g * ld_dpooled + f] * (pool_scale ? pool_scale[g] : 1.f);
        if (d_score != nullptr) uu = u[(int64_t)g * ldu + f];
        for (int r = r0 + rr; r < r1; r += rl) {
            const float zv = z[(int64_t)r * ldz + f];
            float gr = gp;
            if (d_out != nullptr) gr += d_out[(int64_t)r * ld_dout + f];
            if (d_score != nullptr) gr = fmaf(d_score[r], uu, gr);
            if (d_neg != nullptr && r < n_neg) gr += d_neg[(int64_t)r * ld_dneg + f];
            const float v = (fmaf(zv, sc, sh) > 0.f) ? gr : 0.f;
            dy[(int64_t)r * lddy + f] = v;
            s1 += v;
            s2 = fmaf(v, (zv - mu) * rs, s2);
        }
    }
    sh1[threadIdx.x] = s1;
    sh2[threadIdx.x] = s2;
    __syncthreads();
    if (stats != nullptr && threadIdx.x < cw) {
        double a = 0.0, b = 0.0;
        for (int k = 0; k < rl; ++k) { a += (double)sh1[k * cw + threadIdx.x]; b += (double)sh2[k * cw + threadIdx.x]; }
        atomicAdd(&stats[f0 + threadIdx.x], a);
        atomicAdd(&stats[n_feat + f0 + threadIdx.x], b);
    }
}

__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ z, int64_t ldz, int n_rows, int n_feat, const float* __restrict__ mean,
                    const float* __restrict__ rstd, const float* __restrict__ gamma, const double* __restrict__ stats,
                    double count, float* __restrict__ dy, int64_t lddy) {
    const int64_t total = (int64_t)n_rows * n_feat;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / n_feat), f = (int)(i % n_feat);
        const float gsc = (gamma ? gamma[f] : 1.f) * rstd[f];
        float v = dy[(int64_t)r * lddy + f];
        if (stats != nullptr) {
            const float m1 = (float)(stats[f] / count), m2 = (float)(stats[n_feat + f] / count);
            const float xh = (z[(int64_t)r * ldz + f] - mean[f]) * rstd[f];
            v = v - m1 - xh * m2;
        }
        dy[(int64_t)r * lddy + f] = v * gsc;
    }
}

}  // namespace

// tcgen05 implementation (gnm_linear_tc.cu); GNM_ERR_TOO_LARGE = shape does not fit, use the FFMA kernel
int gnm_launch_linear_tc(const float* x, int64_t ldx, int n_rows, int n_in, const float* w, int64_t ldw, int w_is_kn,
                         const float* bias, const float* in_scale, const float* in_shift, float* y, int64_t ldy,
                         int n_out, double* col_stats, const gnm_bn_tail* tail, cudaStream_t stream);
static int g_linear_impl = 0;     // 0 auto, 1 FFMA kernel, 2 tcgen05 kernels only, 3 = 2 with the two-pass backward pair (A/B switch)

int gnm_linear_impl_value() { return g_linear_impl; }

extern "C" int gnm_set_linear_impl(int impl) {
    if (impl < 0 || impl > 3) return GNM_ERR_BAD_ARG;
    g_linear_impl = impl;
    return GNM_OK;
}

extern "C" int gnm_linear(const float* x, int64_t ldx, int n_rows, int n_in, const float* w, int64_t ldw, int w_is_kn,
                          const float* bias, const float* in_scale, const float* in_shift, float* y, int64_t ldy,
                          int n_out, double* col_stats, const gnm_bn_tail* tail, gnm_stream_t stream) {
    if (n_rows < 0 || n_in < 0 || n_out < 0) return GNM_ERR_BAD_ARG;
    if (n_rows == 0 || n_out == 0) return tail ? GNM_ERR_BAD_ARG : GNM_OK;
    if (!x || !w || !y) return GNM_ERR_BAD_ARG;
    if ((in_scale == nullptr) != (in_shift == nullptr)) return GNM_ERR_BAD_ARG;
    // large-M 64-wide layers go to the tensor cores; tiny problems are not worth a persistent 148-CTA launch
    if (g_linear_impl != 1 && (g_linear_impl >= 2 || n_rows >= 4096)) {
        const int rc = gnm_launch_linear_tc(x, ldx, n_rows, n_in, w, ldw, w_is_kn, bias, in_scale, in_shift, y, ldy, n_out,
                                            col_stats, tail, gnm_cast_stream(stream));
        if (rc == GNM_OK || g_linear_impl >= 2 || rc != GNM_ERR_TOO_LARGE) return rc;
    }
    if (tail != nullptr) return GNM_ERR_TOO_LARGE;      // the FFMA kernel has no tail: nothing launched, call gnm_bn_finalize
    dim3 grid((n_rows + LBM - 1) / LBM, (n_out + LBN - 1) / LBN);
    gnm_count_launch(GNM_K_LINEAR_FFMA);
    linear_kernel<<<grid, LTHREADS, 0, gnm_cast_stream(stream)>>>(x, ldx, n_rows, n_in, w, ldw, w_is_kn, bias, in_scale,
                                                                  in_shift, y, ldy, n_out, col_stats);
    GNM_RETURN_IF_LAUNCH_FAILED();
    return GNM_OK;
}

extern "C" int gnm_linear_wgrad(const float* dz, int64_t lddz, const float* x, int64_t ldx, int n_rows, int n_out,
                                int n_in, const float* in_scale, const float* in_shift, float* dw, int64_t lddw,
                                float* dbias, gnm_stream_t stream) {
    if (n_rows < 0 || n_in < 0 || n_out < 0) return GNM_ERR_BAD_ARG;
    if (n_rows == 0 || n_out == 0) return GNM_OK;
    if (!dz || (n_in > 0 && (!x || !dw))) return GNM_ERR_BAD_ARG;
    if ((in_scale == nullptr) != (in_shift == nullptr)) return GNM_ERR_BAD_ARG;
    const int to = (n_out + WBO - 1) / WBO, ti = n_in > 0 ? (n_in + WBI - 1) / WBI : 1;
    int slabs = (148 * 4) / (to * ti);
    if (slabs < 1) slabs = 1;
    int rows_per_slab = (n_rows + slabs - 1) / slabs;
    rows_per_slab = ((rows_per_slab + WBR - 1) / WBR) * WBR;
    if (rows_per_slab < 256) rows_per_slab = 256;
    slabs = (n_rows + rows_per_slab - 1) / rows_per_slab;
    dim3 grid(to, ti, slabs);
    gnm_count_launch(GNM_K_LINEAR_WGRAD_FFMA);
    linear_wgrad_kernel<<<grid, 256, 0, gnm_cast_stream(stream)>>>(dz, lddz, x, ldx, n_rows, n_out, n_in, in_scale,
                                                                   in_shift, dw, lddw, dbias, rows_per_slab);
    GNM_RETURN_IF_LAUNCH_FAILED();
    return GNM_OK;
}

extern "C" int gnm_col_stats(const float* x, int64_t ldx, int n_rows, int n_feat, double* col_stats,
                             gnm_stream_t stream) {
    if (n_rows < 0 || n_feat < 0) return GNM_ERR_BAD_ARG;
    if (n_rows == 0 || n_feat == 0) return GNM_OK;
    if (!x || !col_stats) return GNM_ERR_BAD_ARG;
    const int fparts = (n_feat + 255) / 256;
    int ctas = 148 * 4;
    int rows_per_cta = (n_rows + ctas - 1) / ctas;
    if (rows_per_cta < 32) rows_per_cta = 32;
    ctas = (n_rows + rows_per_cta - 1) / rows_per_cta;
    dim3 grid(ctas, fparts);
    gnm_count_launch(GNM_K_OTHER);
    col_stats_kernel<<<grid, 256, 0, gnm_cast_stream(stream)>>>(x, ldx, n_rows, n_feat, col_stats, rows_per_cta);
    GNM_RETURN_IF_LAUNCH_FAILED();
    return GNM_OK;
}

int gnm_p2p_abort_flag_mlp(int* aborted) {
    int v = 0, zero = 0;
    cudaError_t e = cudaMemcpyFromSymbol(&v, g_p2p_abort, sizeof(int));
    if (e != cudaSuccess) return (int)e;
    e = cudaMemcpyToSymbol(g_p2p_abort, &zero, sizeof(int));
    if (aborted) *aborted |= v;
    return e == cudaSuccess ? GNM_OK : (int)e;
}

extern "C" int gnm_bn_finalize(double* col_stats, double count, const float* gamma, const float* beta, float eps,
                               float momentum, float* running_mean, float* running_var, int64_t* num_batches_tracked,
                               float* scale, float* shift, float* mean, float* rstd, int n_feat,
                               const gnm_p2p_comm* comm, gnm_stream_t stream) {
    if (n_feat < 0 || count <= 0.0) return GNM_ERR_BAD_ARG;
    if (n_feat == 0) return GNM_OK;
    if (!col_stats || !scale || !shift || !mean || !rstd) return GNM_ERR_BAD_ARG;
    P2PArgs pa;
    const int prc = p2p_args(comm, 2 * n_feat, &pa);      // 2 n_feat <= 256 doubles: the launch below is one CTA
    if (prc < 0) return prc;
    gnm_count_launch(GNM_K_OTHER);
    bn_finalize_kernel<<<(n_feat + 127) / 128, 128, 0, gnm_cast_stream(stream)>>>(
        col_stats, count, gamma, beta, eps, momentum, running_mean, running_var, num_batches_tracked, scale, shift, mean,
        rstd, n_feat, pa);
    GNM_RETURN_IF_LAUNCH_FAILED();
    return GNM_OK;
}

extern "C" int gnm_bn_eval_affine(const float* running_mean, const float* running_var, const float* gamma,
                                  const float* beta, float eps, float* scale, float* shift, float* mean, float* rstd,
                                  int n_feat, gnm_stream_t stream) {
    if (n_feat < 0) return GNM_ERR_BAD_ARG;
    if (n_feat == 0) return GNM_OK;
    if (!running_mean || !running_var || !scale || !shift || !mean || !rstd) return GNM_ERR_BAD_ARG;
    gnm_count_launch(GNM_K_OTHER);
    bn_eval_affine_kernel<<<(n_feat + 127) / 128, 128, 0, gnm_cast_stream(stream)>>>(
        running_mean, running_var, gamma, beta, eps, scale, shift, mean, rstd, n_feat);
    GNM_RETURN_IF_LAUNCH_FAILED();
    return GNM_OK;
}

extern "C" int gnm_bn_relu_readout(const float* z, int64_t ldz, int n_rows, int n_feat, const float* scale,
                                   const float* shift, float* h, int64_t ldh, const int32_t* node_off, int n_graphs,
                                   const float* pool_scale, float* pooled, int64_t ld_pooled, gnm_stream_t stream) {
    if (n_rows < 0 || n_feat < 0 || n_graphs < 0) return GNM_ERR_BAD_ARG;
    if (n_rows == 0 || n_feat == 0 || n_graphs == 0) return GNM_OK;
    if (!z || !scale || !shift || !node_off) return GNM_ERR_BAD_ARG;
    const bool vec = (n_feat % 4 == 0) && n_feat <= 1024 && (ldz % 4 == 0) && gnm_aligned16(z) && gnm_aligned16(scale) &&
                     gnm_aligned16(shift) && (h == nullptr || ((ldh % 4 == 0) && gnm_aligned16(h)));
    if (vec) {
        gnm_count_launch(GNM_K_OTHER);
        bn_relu_readout_vec_kernel<<<n_graphs, 256, 0, gnm_cast_stream(stream)>>>(z, ldz, n_feat, scale, shift, h, ldh,
                                                                                  node_off, pool_scale, pooled,
                                                                                  ld_pooled);
        GNM_RETURN_IF_LAUNCH_FAILED();
        return GNM_OK;
    }
    dim3 grid(n_graphs, (n_feat + 255) / 256);
    gnm_count_launch(GNM_K_OTHER);
    bn_relu_readout_kernel<<<grid, 256, 0, gnm_cast_stream(stream)>>>(z, ldz, n_feat, scale, shift, h, ldh, node_off,
                                                                      pool_scale, pooled, ld_pooled);
    GNM_RETURN_IF_LAUNCH_FAILED();
    return GNM_OK;
}

extern "C" int gnm_relu_bn_bwd_reduce(const float* z, int64_t ldz, int n_rows, int n_feat, const float* scale,
                                      const float* shift, const float* mean, const float* rstd, const float* d_out,
                                      int64_t ld_dout, const float* d_pooled, int64_t ld_dpooled,
                                      const float* pool_scale, const float* d_score, const float* u, int64_t ldu,
                                      const float* d_neg, int64_t ld_dneg, int n_neg, const int32_t* node_off,
                                      int n_graphs, float* dy, int64_t lddy, double* stats, const gnm_bn_tail* tail,
                                      gnm_stream_t stream) {
    if (n_rows < 0 || n_feat < 0 || n_graphs < 0) return GNM_ERR_BAD_ARG;
    if (n_rows == 0 || n_feat == 0 || n_graphs == 0) return GNM_OK;
    if (!z || !scale || !shift || !mean || !rstd || !node_off || !dy) return GNM_ERR_BAD_ARG;
    if (d_score != nullptr && u == nullptr) return GNM_ERR_BAD_ARG;
    const bool vec = (n_feat % 4 == 0) && n_feat <= 1024 && (ldz % 4 == 0) && (lddy % 4 == 0) && gnm_aligned16(z) &&
                     gnm_aligned16(dy) && gnm_aligned16(scale) && gnm_aligned16(shift) && gnm_aligned16(mean) &&
                     gnm_aligned16(rstd) && (d_out == nullptr || ((ld_dout % 4 == 0) && gnm_aligned16(d_out))) &&
                     (d_neg == nullptr || ((ld_dneg % 4 == 0) && gnm_aligned16(d_neg)));
    BnTailDev td;
    const int trc = bn_tail_args(tail, stats, n_feat, &td);
    if (trc != GNM_OK) return trc;
    if (vec) {
        gnm_count_launch(GNM_K_OTHER);
        const int per_sm = d_out != nullptr ? 2 : 3;                 // resident CTAs per SM (launch bounds)
        const int grid = n_graphs < 148 * per_sm ? n_graphs : 148 * per_sm;
        if (d_out != nullptr)
            relu_bn_bwd_reduce_vec_kernel<true><<<grid, 256, 0, gnm_cast_stream(stream)>>>(
                z, ldz, n_feat, scale, shift, mean, rstd, d_out, ld_dout, d_pooled, ld_dpooled, pool_scale, d_score, u, ldu,
                d_neg, ld_dneg, n_neg, node_off, n_graphs, dy, lddy, stats, td);
        else
            relu_bn_bwd_reduce_vec_kernel<false><<<grid, 256, 0, gnm_cast_stream(stream)>>>(
                z, ldz, n_feat, scale, shift, mean, rstd, d_out, ld_dout, d_pooled, ld_dpooled, pool_scale, d_score, u, ldu,
                d_neg, ld_dneg, n_neg, node_off, n_graphs, dy, lddy, stats, td);
        GNM_RETURN_IF_LAUNCH_FAILED();
        return GNM_OK;
    }
    if (tail != nullptr) return GNM_ERR_TOO_LARGE;      // the scalar kernel has no tail: nothing launched
    dim3 grid(n_graphs, (n_feat + 255) / 256);
    gnm_count_launch(GNM_K_OTHER);
    relu_bn_bwd_reduce_kernel<<<grid, 256, 0, gnm_cast_stream(stream)>>>(
        z, ldz, n_feat, scale, shift, mean, rstd, d_out, ld_dout, d_pooled, ld_dpooled, pool_scale, d_score, u, ldu,
        d_neg, ld_dneg, n_neg, node_off, dy, lddy, stats);
    GNM_RETURN_IF_LAUNCH_FAILED();
    return GNM_OK;
}

extern "C" int gnm_bn_bwd_apply(const float* z, int64_t ldz, int n_rows, int n_feat, const float* mean,
                                const float* rstd, const float* gamma, const double* stats, double count, float* dy,
                                int64_t lddy, gnm_stream_t stream) {
    if (n_rows < 0 || n_feat < 0) return GNM_ERR_BAD_ARG;
    if (n_rows == 0 || n_feat == 0) return GNM_OK;
    if (!rstd || !dy || (stats != nullptr && (!z || !mean || count <= 0.0))) return GNM_ERR_BAD_ARG;
    const int64_t total = (int64_t)n_rows * n_feat;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    gnm_count_launch(GNM_K_OTHER);
    bn_bwd_apply_kernel<<<(int)blocks, 256, 0, gnm_cast_stream(stream)>>>(z, ldz, n_rows, n_feat, mean, rstd, gamma,
                                                                          stats, count, dy, lddy);
    GNM_RETURN_IF_LAUNCH_FAILED();
    return GNM_OK;
}
