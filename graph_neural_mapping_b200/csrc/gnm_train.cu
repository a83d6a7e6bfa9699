// The [B, L*F]-sized remainder of a training step: everything main.py:31-41 does around the encoder that
// is not an M-sized tensor operation - prediction heads + dropout + CrossEntropy (graphcnn.py:228-231, main.py:35),
// BCEWithLogits over the DGI scores (main.py:34), the summary / bilinear glue of the Discriminator
// (graphcnn.py:238-239, discriminator.py:28-29 refactored as u_g = W c_g) and torch.optim.Adam (main.py:39-41,136).
// On torch these are ~110 tiny kernels per step (13 % of the step's GPU time at B = 1024); here they are 8 launches.
// All of it is latency-bound fp32 CUDA-core work: a few hundred KB per launch.
#include <math.h>

#include "gnm_common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------------------------
// prediction heads + dropout + CrossEntropy, forward AND backward in one pass over g_f
// ------------------------------------------------------------------------------------------------------------------
constexpr int HD_MAX_LAYERS = 16;
constexpr int HD_MAX_CLASSES = 8;
constexpr int HD_WARPS = 8;

struct HeadsParams {
    const float* g_f; int64_t ldg;                 // [B, L*F] readout features
    const float* w[HD_MAX_LAYERS];                 // linears_prediction[l].weight [C, F]
    const float* b[HD_MAX_LAYERS];                 // linears_prediction[l].bias   [C]
    float* dw[HD_MAX_LAYERS];                      // gradients (written, not accumulated)
    float* db[HD_MAX_LAYERS];
    const float* mask;                             // nullable [L, B, C]: dropout keep mask already scaled by 1/(1-p)
    const int64_t* labels;                         // [B]
    float inv_count;                               // 1 / B (CrossEntropyLoss mean)
    float* c_logit;                                // [B, C]
    double* loss_acc;                              // += sum_b CE_b * inv_count
    float* d_gf; int64_t ldd;                      // [B, L*F] gradient of the CE term wrt g_f (written)
    float* ws;                                     // workspace: gridDim.x x (L*C*F + L*C) partial sums
    unsigned int* counter;                         // zero on entry; reset to zero by the last CTA
    int n_graphs, n_layers, n_feat, n_classes;
    const float* d_logit_in;                       // HD_BWD only: d loss / d c_logit [B, C] from the caller's loss
};

// HD_FUSED: heads + CrossEntropy forward and backward; HD_FWD: c_logit only; HD_BWD: backward for a given d c_logit
enum { HD_FUSED = 0, HD_FWD = 1, HD_BWD = 2 };

// One warp per graph row; lanes own features f = lane, lane + 32, ... Per-warp gradient partials live in shared
// memory (owner-lane updates, no atomics); CTAs write their partial to the workspace and the last CTA to finish adds
// them in CTA order: deterministic, one launch, no pre-zeroed outputs.
template <int MODE>
__global__ void __launch_bounds__(HD_WARPS * 32) heads_ce_kernel(const HeadsParams p) {
    extern __shared__ float hd_smem[];
    __shared__ float s_logit[HD_WARPS][HD_MAX_CLASSES];
    __shared__ bool s_last;
    const int L = p.n_layers, C = p.n_classes, F = p.n_feat;
    const int gsz = L * C * F + L * C;             // dW entries then db entries
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* acc = hd_smem + (size_t)warp * gsz;
    if (MODE != HD_FWD) {
        for (int i = lane; i < gsz; i += 32) acc[i] = 0.f;
        __syncwarp();
    }
    double loss_w = 0.0;
    for (int row = blockIdx.x * HD_WARPS + warp; row < p.n_graphs; row += gridDim.x * HD_WARPS) {
        const float* g = p.g_f + (int64_t)row * p.ldg;
        // ---- forward: logit[c] = sum_l mask[l,row,c] * (<W_l[c], g_l> + b_l[c]) -------------------------------
        float logit[HD_MAX_CLASSES];
#pragma unroll
        for (int c = 0; c < HD_MAX_CLASSES; ++c) logit[c] = 0.f;
        for (int l = 0; l < L && MODE != HD_BWD; ++l) {
#pragma unroll
            for (int c = 0; c < HD_MAX_CLASSES; ++c) {
                if (c >= C) break;
                float d = 0.f;
                for (int f = lane; f < F; f += 32) d = fmaf(__ldg(p.w[l] + c * F + f), g[l * F + f], d);
                d = warp_sum(d) + __ldg(p.b[l] + c);
                const float mk = p.mask ? p.mask[((int64_t)l * p.n_graphs + row) * C + c] : 1.f;
                logit[c] += mk * d;
            }
        }
        float dl[HD_MAX_CLASSES];
        if (MODE == HD_FWD) {
#pragma unroll
            for (int c = 0; c < HD_MAX_CLASSES; ++c)
                if (c < C && lane == 0) p.c_logit[(int64_t)row * C + c] = logit[c];
            continue;
        }
        if (MODE == HD_BWD) {
#pragma unroll
            for (int c = 0; c < HD_MAX_CLASSES; ++c) dl[c] = c < C ? __ldg(p.d_logit_in + (int64_t)row * C + c) : 0.f;
        } else {
        // ---- CrossEntropy (mean over the batch) and its gradient ------------------------------------------------
        float mx = logit[0];
#pragma unroll
        for (int c = 1; c < HD_MAX_CLASSES; ++c) if (c < C) mx = fmaxf(mx, logit[c]);
        float se = 0.f;
#pragma unroll
        for (int c = 0; c < HD_MAX_CLASSES; ++c) if (c < C) se += expf(logit[c] - mx);
        const float lse = mx + logf(se);
        const int lab = (int)p.labels[row];
        float picked = 0.f;
#pragma unroll
        for (int c = 0; c < HD_MAX_CLASSES; ++c) {
            dl[c] = 0.f;
            if (c < C) {
                dl[c] = (expf(logit[c] - lse) - (c == lab ? 1.f : 0.f)) * p.inv_count;
                if (c == lab) picked = logit[c];
                if (lane == 0) p.c_logit[(int64_t)row * C + c] = logit[c];
            }
        }
        if (lane == 0) loss_w += (double)((lse - picked) * p.inv_count);
        }
        // ---- backward: d g_f, dW, db --------------------------------------------------------------------------------
        for (int l = 0; l < L; ++l) {
            float dm[HD_MAX_CLASSES];
#pragma unroll
            for (int c = 0; c < HD_MAX_CLASSES; ++c) {
                dm[c] = 0.f;
                if (c < C) dm[c] = dl[c] * (p.mask ? p.mask[((int64_t)l * p.n_graphs + row) * C + c] : 1.f);
            }
            for (int f = lane; f < F; f += 32) {
                const float gv = g[l * F + f];
                float dg = 0.f;
#pragma unroll
                for (int c = 0; c < HD_MAX_CLASSES; ++c) {
                    if (c >= C) break;
                    dg = fmaf(dm[c], __ldg(p.w[l] + c * F + f), dg);
                    acc[(l * C + c) * F + f] = fmaf(dm[c], gv, acc[(l * C + c) * F + f]);
                }
                p.d_gf[(int64_t)row * p.ldd + l * F + f] = dg;
            }
            if (lane < C) acc[L * C * F + l * C + lane] += dm[lane];      // dm[] is warp-uniform
        }
    }
    if (MODE == HD_FWD) return;
    __syncthreads();
    // CTA partial = sum of its warps (fixed order) -> workspace
    float* mine = p.ws + (size_t)blockIdx.x * gsz;
    for (int i = threadIdx.x; i < gsz; i += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < HD_WARPS; ++w) s += hd_smem[(size_t)w * gsz + i];
        mine[i] = s;
    }
    if (MODE == HD_FUSED) {
        loss_w = warp_sum_d(loss_w);
        if (lane == 0 && loss_w != 0.0) atomicAdd(p.loss_acc, loss_w);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(p.counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int i = threadIdx.x; i < gsz; i += blockDim.x) {
        float s = 0.f;
        for (unsigned int b = 0; b < gridDim.x; ++b) s += __ldcg(p.ws + (size_t)b * gsz + i);
        if (i < L * C * F) {
            const int l = i / (C * F);
            p.dw[l][i - l * C * F] = s;
        } else {
            const int j = i - L * C * F, l = j / C;
            p.db[l][j - l * C] = s;
        }
    }
    if (threadIdx.x == 0) *p.counter = 0u;
}

// ------------------------------------------------------------------------------------------------------------------
// BCEWithLogits over the DGI scores (targets: first n_pos rows 1, the rest 0; main.py:32-34), mean reduction
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bce_logits_kernel(const float* __restrict__ x, int64_t n, int64_t n_pos, float grad_scale,
                                                         double loss_scale, double* __restrict__ loss_acc,
                                                         float* __restrict__ dx) {
    __shared__ double part[8];
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = x[i];
        const float y = i < n_pos ? 1.f : 0.f;
        // torch's stable form: max(x, 0) - x*y + log1p(exp(-|x|))
        const float l = fmaxf(v, 0.f) - v * y + log1pf(expf(-fabsf(v)));
        acc += (double)l;
        if (dx) dx[i] = (1.f / (1.f + expf(-v)) - y) * grad_scale;
    }
    acc = warp_sum_d(acc);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += part[i];
        atomicAdd(loss_acc, t * loss_scale);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// small strided fp32 GEMM for the [B, L*F] glue:  C[m,n] = epi( sum_k pro(A[m,k]) * B[k,n] )
//   A element (m,k) at a[m*sam + k*sak], B element (k,n) at b[k*sbk + n*sbn] - covers NT / TN / NN without copies.
//   SIGMOID_A: pro = sigmoid (and the activated A is also written to a_out, row-major [M, K]);
//   DSIG_EPI:  C = add[m,n] (nullable) + acc * s[m,n] * (1 - s[m,n])   (backward through c = sigmoid(g_f)).
// ------------------------------------------------------------------------------------------------------------------
constexpr int SG_T = 64, SG_K = 16, SG_P = SG_T + 4;      // 64 x 64 tile, 16-deep k slices, rows padded to 68 floats

// One operand tile (SG_K x SG_T, k-major in shared memory) from a strided matrix: element (i, k) at p[i*si + k*sk],
// i = row of A / column of B. 128-bit loads along whichever axis is contiguous when the tile is interior and aligned.
template <bool SIGMOID>
__device__ __forceinline__ void sg_load_tile(float (*dst)[SG_P], const float* __restrict__ p, int64_t si, int64_t sk, int i0,
                                             int k0, int n_i, int n_k, float* __restrict__ act_out, int64_t ld_act,
                                             bool write_act) {
    const int t = threadIdx.x;
    const bool interior = i0 + SG_T <= n_i && k0 + SG_K <= n_k;
    if (sk == 1 && interior && (si & 3) == 0 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
        const int i = t >> 2, k4 = (t & 3) * 4;                    // 64 rows x 4 float4 along k
        float4 v = *reinterpret_cast<const float4*>(p + (int64_t)(i0 + i) * si + k0 + k4);
        if (SIGMOID) {
            v.x = 1.f / (1.f + expf(-v.x)); v.y = 1.f / (1.f + expf(-v.y));
            v.z = 1.f / (1.f + expf(-v.z)); v.w = 1.f / (1.f + expf(-v.w));
            if (write_act) *reinterpret_cast<float4*>(act_out + (int64_t)(i0 + i) * ld_act + k0 + k4) = v;
        }
        dst[k4][i] = v.x; dst[k4 + 1][i] = v.y; dst[k4 + 2][i] = v.z; dst[k4 + 3][i] = v.w;
    } else if (si == 1 && interior && (sk & 3) == 0 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0) && !SIGMOID) {
        const int k = t >> 4, i4 = (t & 15) * 4;                   // 16 k x 16 float4 along i
        *reinterpret_cast<float4*>(&dst[k][i4]) = *reinterpret_cast<const float4*>(p + (int64_t)(k0 + k) * sk + i0 + i4);
    } else {
        for (int e = t; e < SG_K * SG_T; e += 256) {
            int k, i;
            if (sk == 1) { k = e % SG_K; i = e / SG_K; } else { i = e % SG_T; k = e / SG_T; }
            float v = 0.f;
            if (i0 + i < n_i && k0 + k < n_k) {
                v = p[(int64_t)(i0 + i) * si + (int64_t)(k0 + k) * sk];
                if (SIGMOID) {
                    v = 1.f / (1.f + expf(-v));
                    if (write_act) act_out[(int64_t)(i0 + i) * ld_act + k0 + k] = v;
                }
            }
            dst[k][i] = v;
        }
    }
}

// grid = (n tiles, m tiles, k splits); with more than one k split the partial products are added with atomics into a
// C that the host entry zeroed (the epilogue variants run unsplit).
template <bool SIGMOID_A, bool DSIG_EPI>
__global__ void __launch_bounds__(256) small_gemm_kernel(const float* __restrict__ a, int64_t sam, int64_t sak,
                                                         const float* __restrict__ b, int64_t sbk, int64_t sbn,
                                                         float* __restrict__ c, int64_t ldc, int m_tot, int n_tot, int k_tot,
                                                         int k_per_split, float* __restrict__ a_out, int64_t lda_out,
                                                         const float* __restrict__ epi_s, int64_t lds,
                                                         const float* __restrict__ epi_add, int64_t ldadd) {
    __shared__ __align__(16) float sa[SG_K][SG_P];
    __shared__ __align__(16) float sb[SG_K][SG_P];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * SG_T, n0 = blockIdx.x * SG_T;
    const int k_begin = blockIdx.z * k_per_split, k_end = min(k_tot, k_begin + k_per_split);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    // Interior tiles of aligned operands (every tile of the training step's three products): the next k slice's global
    // loads are issued into registers BEFORE the current slice is multiplied - the unpipelined loop below pays one global
    // round trip per 16-deep slice, 20 of them for K = 320, which was most of these kernels' 30-40 us.
    const int t = threadIdx.x;
    const int mode_a = (sak == 1 && (sam & 3) == 0) ? 0 : ((sam == 1 && (sak & 3) == 0 && !SIGMOID_A) ? 1 : -1);
    const int mode_b = (sbk == 1 && (sbn & 3) == 0) ? 0 : ((sbn == 1 && (sbk & 3) == 0) ? 1 : -1);
    const bool piped = mode_a >= 0 && mode_b >= 0 && m0 + SG_T <= m_tot && n0 + SG_T <= n_tot && k_begin < k_end &&
                       ((k_end - k_begin) % SG_K) == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
    if (piped) {
        // mode 0: contiguous along k - thread -> row t >> 2, float4 (t & 3) along k; mode 1: contiguous along the row /
        // column index - thread -> k = t >> 4, float4 (t & 15) along i
        const float* pa = mode_a == 0 ? a + (int64_t)(m0 + (t >> 2)) * sam + (t & 3) * 4 : a + (int64_t)(t >> 4) * sak + m0 + (t & 15) * 4;
        const float* pb = mode_b == 0 ? b + (int64_t)(n0 + (t >> 2)) * sbn + (t & 3) * 4 : b + (int64_t)(t >> 4) * sbk + n0 + (t & 15) * 4;
        const int64_t step_a = mode_a == 0 ? SG_K : SG_K * sak, step_b = mode_b == 0 ? SG_K : SG_K * sbk;
        pa += (int64_t)k_begin * (mode_a == 0 ? 1 : sak);
        pb += (int64_t)k_begin * (mode_b == 0 ? 1 : sbk);
        float4 ra = *reinterpret_cast<const float4*>(pa), rb = *reinterpret_cast<const float4*>(pb);
        for (int k0 = k_begin; k0 < k_end; k0 += SG_K) {
            if (SIGMOID_A) {
                ra.x = 1.f / (1.f + expf(-ra.x)); ra.y = 1.f / (1.f + expf(-ra.y));
                ra.z = 1.f / (1.f + expf(-ra.z)); ra.w = 1.f / (1.f + expf(-ra.w));
                if (blockIdx.x == 0)            // mode_a == 0 here (the sigmoid variant has no mode 1)
                    *reinterpret_cast<float4*>(a_out + (int64_t)(m0 + (t >> 2)) * lda_out + k0 + (t & 3) * 4) = ra;
            }
            if (mode_a == 0) {
                const int i = t >> 2, k4 = (t & 3) * 4;
                sa[k4][i] = ra.x; sa[k4 + 1][i] = ra.y; sa[k4 + 2][i] = ra.z; sa[k4 + 3][i] = ra.w;
            } else {
                *reinterpret_cast<float4*>(&sa[t >> 4][(t & 15) * 4]) = ra;
            }
            if (mode_b == 0) {
                const int i = t >> 2, k4 = (t & 3) * 4;
                sb[k4][i] = rb.x; sb[k4 + 1][i] = rb.y; sb[k4 + 2][i] = rb.z; sb[k4 + 3][i] = rb.w;
            } else {
                *reinterpret_cast<float4*>(&sb[t >> 4][(t & 15) * 4]) = rb;
            }
            __syncthreads();
            if (k0 + SG_K < k_end) {
                pa += step_a;
                pb += step_b;
                ra = *reinterpret_cast<const float4*>(pa);
                rb = *reinterpret_cast<const float4*>(pb);
            }
#pragma unroll
            for (int kk = 0; kk < SG_K; ++kk) {
                // four consecutive rows / columns per thread: one 128-bit shared-memory load per operand and k
                const float4 a4 = *reinterpret_cast<const float4*>(&sa[kk][ty * 4]);
                const float4 b4 = *reinterpret_cast<const float4*>(&sb[kk][tx * 4]);
                const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
    for (int k0 = k_begin; k0 < k_end && !piped; k0 += SG_K) {
        sg_load_tile<SIGMOID_A>(sa, a, sam, sak, m0, k0, m_tot, k_end, a_out, lda_out, blockIdx.x == 0);
        sg_load_tile<false>(sb, b, sbn, sbk, n0, k0, n_tot, k_end, nullptr, 0, false);
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < SG_K; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4*>(&sa[kk][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&sb[kk][tx * 4]);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= m_tot) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= n_tot) continue;
            float v = acc[i][j];
            if (DSIG_EPI) {
                const float s = epi_s[(int64_t)m * lds + n];
                v = v * s * (1.f - s);
                if (epi_add) v += epi_add[(int64_t)m * ldadd + n];
            }
            if (gridDim.z > 1) atomicAdd(&c[(int64_t)m * ldc + n], v);
            else c[(int64_t)m * ldc + n] = v;
        }
    }
}

// d_neg[j, :] = sum_{g : neg_idx[g] == j} s2[g] * u[g, :]   (one CTA per row j; deterministic, rows nobody names are zero)
__global__ void __launch_bounds__(128) dgi_neg_grad_kernel(const int32_t* __restrict__ neg_idx, const float* __restrict__ s2,
                                                           const float* __restrict__ u, int64_t ldu, int n_graphs, int width,
                                                           float* __restrict__ d_neg, int64_t ldn) {
    __shared__ int s_hits[64];
    __shared__ int s_n;
    const int j = blockIdx.x;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (int g = threadIdx.x; g < n_graphs; g += blockDim.x)
        if (neg_idx[g] == j) {
            const int k = atomicAdd(&s_n, 1);
            if (k < 64) s_hits[k] = g;
        }
    __syncthreads();
    const int n = s_n < 64 ? s_n : 64;
    // order the (few) hits by graph id so that the sum does not depend on the atomics' arrival order
    if (threadIdx.x == 0)
        for (int a = 1; a < n; ++a) {
            const int v = s_hits[a];
            int b2 = a - 1;
            while (b2 >= 0 && s_hits[b2] > v) { s_hits[b2 + 1] = s_hits[b2]; --b2; }
            s_hits[b2 + 1] = v;
        }
    __syncthreads();
    for (int f = threadIdx.x; f < width; f += blockDim.x) {
        float acc = 0.f;
        for (int k = 0; k < n; ++k) {
            const int g = s_hits[k];
            acc = fmaf(s2[g], u[(int64_t)g * ldu + f], acc);
        }
        d_neg[(int64_t)j * ldn + f] = acc;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Adam over a table of tensors (torch.optim.Adam defaults: no amsgrad, no maximize; main.py:136), one launch
// ------------------------------------------------------------------------------------------------------------------
constexpr int AD_MAX_TENSORS = 64;
constexpr int AD_CHUNK = 2048;          // elements per CTA work item

struct AdamTable {
    float* p[AD_MAX_TENSORS];
    const float* g[AD_MAX_TENSORS];
    int numel[AD_MAX_TENSORS];
    int state_off[AD_MAX_TENSORS];      // offset of the tensor's exp_avg / exp_avg_sq in the flat state buffers
    int chunk0[AD_MAX_TENSORS + 1];     // prefix of per-tensor chunk counts
    int n_tensors;
};

__global__ void __launch_bounds__(256) adam_kernel(const AdamTable t, float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                                                   float* __restrict__ step, const float* __restrict__ lr, double beta1_d,
                                                   double beta2_d, float eps, float weight_decay, float grad_scale,
                                                   const double* __restrict__ loss_terms, int n_loss_terms,
                                                   float* __restrict__ loss_out) {
    // every CTA derives the same step number from the (not yet incremented) device counter; the LAST chunk's CTA
    // would race with readers if it incremented in place, so the counter is double-buffered: step[0] is read,
    // step[1] receives step[0] + 1 and a tiny tail of this kernel's launch (CTA 0 of the next launch) copies it back
    // scalar arithmetic in double, as torch's Python-side bias corrections (1 - beta ** step, lr / bc1, sqrt(bc2)) and
    // its `1 - beta` weights are: (float)(1 - 0.999) and 1.f - 0.999f differ by 5e-5 relative
    __shared__ float s_sc[6];
    const float t_new = step[0] + 1.f;
    if (threadIdx.x == 0) {
        const double bc1 = 1.0 - pow(beta1_d, (double)t_new), bc2 = 1.0 - pow(beta2_d, (double)t_new);
        s_sc[0] = (float)((double)lr[0] / bc1);
        s_sc[1] = (float)sqrt(bc2);
        s_sc[2] = (float)(1.0 - beta1_d);
        s_sc[3] = (float)(1.0 - beta2_d);
        s_sc[4] = (float)beta2_d;
    }
    __syncthreads();
    const float step_size = s_sc[0], bc2_sqrt = s_sc[1], omb1 = s_sc[2], omb2 = s_sc[3], beta2 = s_sc[4];
    const int total_chunks = t.chunk0[t.n_tensors];
    for (int item = blockIdx.x; item < total_chunks; item += gridDim.x) {
        int ti = 0;
        while (ti + 1 < t.n_tensors && t.chunk0[ti + 1] <= item) ++ti;
        const int e0 = (item - t.chunk0[ti]) * AD_CHUNK;
        const int e1 = min(e0 + AD_CHUNK, t.numel[ti]);
        float* pp = t.p[ti];
        const float* gg = t.g[ti];
        float* m = exp_avg + t.state_off[ti];
        float* v = exp_avg_sq + t.state_off[ti];
        for (int e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
            float g = gg[e] * grad_scale;
            const float w = pp[e];
            if (weight_decay != 0.f) g = fmaf(weight_decay, w, g);
            // torch: exp_avg.lerp_(grad, 1 - beta1); exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
            const float mm = m[e] + (g - m[e]) * omb1;
            const float vv = fmaf(omb2, g * g, v[e] * beta2);
            m[e] = mm;
            v[e] = vv;
            // torch (capturable): denom = exp_avg_sq.sqrt() / bias_correction2_sqrt + eps; param.addcdiv_(exp_avg, denom, -step_size)
            const float denom = sqrtf(vv) / bc2_sqrt + eps;
            pp[e] = w - step_size * (mm / denom);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        step[1] = t_new;
        if (loss_out != nullptr) {
            double s = 0.0;
            for (int i = 0; i < n_loss_terms; ++i) s += loss_terms[i];
            *loss_out = (float)s;
        }
    }
}

__global__ void adam_commit_step_kernel(float* step) { step[0] = step[1]; }

}  // namespace

extern "C" int gnm_heads_ce_workspace(int n_graphs, int n_layers, int n_feat, int n_classes) {
    int ctas = (n_graphs + HD_WARPS - 1) / HD_WARPS;
    if (ctas > 64) ctas = 64;
    if (ctas < 1) ctas = 1;
    return ctas * (n_layers * n_classes * n_feat + n_layers * n_classes);
}

extern "C" int gnm_heads_ce(const float* g_f, int64_t ldg, int n_graphs, int n_layers, int n_feat, int n_classes,
                            const float* const* weights, const float* const* biases, const float* mask,
                            const int64_t* labels, float inv_count, float* c_logit, double* loss_acc, float* d_gf,
                            int64_t ldd, float* const* d_weights, float* const* d_biases, float* workspace,
                            int64_t workspace_floats, unsigned int* counter, gnm_stream_t stream) {
    if (n_graphs < 0 || n_layers < 1 || n_feat < 1 || n_classes < 1) return GNM_ERR_BAD_ARG;
    if (n_layers > HD_MAX_LAYERS || n_classes > HD_MAX_CLASSES) return GNM_ERR_TOO_LARGE;
    if (n_graphs == 0) return GNM_OK;
    if (!g_f || !weights || !biases || !labels || !c_logit || !loss_acc || !d_gf || !d_weights || !d_biases || !workspace ||
        !counter)
        return GNM_ERR_BAD_ARG;
    const int gsz = n_layers * n_classes * n_feat + n_layers * n_classes;
    const size_t smem = (size_t)HD_WARPS * gsz * sizeof(float);
    if (smem > 200 * 1024) return GNM_ERR_TOO_LARGE;
    int ctas = (n_graphs + HD_WARPS - 1) / HD_WARPS;
    if (ctas > 64) ctas = 64;
    if (workspace_floats < (int64_t)ctas * gsz) return GNM_ERR_BAD_ARG;
    HeadsParams p;
    p.g_f = g_f; p.ldg = ldg; p.mask = mask; p.labels = labels; p.inv_count = inv_count; p.c_logit = c_logit;
    p.loss_acc = loss_acc; p.d_gf = d_gf; p.ldd = ldd; p.ws = workspace; p.counter = counter;
    p.n_graphs = n_graphs; p.n_layers = n_layers; p.n_feat = n_feat; p.n_classes = n_classes;
    for (int l = 0; l < HD_MAX_LAYERS; ++l) {
        const bool on = l < n_layers;
        p.w[l] = on ? weights[l] : nullptr; p.b[l] = on ? biases[l] : nullptr;
        p.dw[l] = on ? d_weights[l] : nullptr; p.db[l] = on ? d_biases[l] : nullptr;
        if (on && (!p.w[l] || !p.b[l] || !p.dw[l] || !p.db[l])) return GNM_ERR_BAD_ARG;
    }
    p.d_logit_in = nullptr;
    cudaError_t e = cudaFuncSetAttribute(heads_ce_kernel<HD_FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    gnm_count_launch(GNM_K_OTHER);
    heads_ce_kernel<HD_FUSED><<<ctas, HD_WARPS * 32, smem, gnm_cast_stream(stream)>>>(p);
    GNM_RETURN_IF_LAUNCH_FAILED();
    return GNM_OK;
}

extern "C" int gnm_heads_fwd(const float* g_f, int64_t ldg, int n_graphs, int n_layers, int n_feat, int n_classes,
                             const float* const* weights, const float* const* biases, const float* mask, float* c_logit,
                             gnm_stream_t stream) {
    if (n_graphs < 0 || n_layers < 1 || n_feat < 1 || n_classes < 1) return GNM_ERR_BAD_ARG;
    if (n_layers > HD_MAX_LAYERS || n_classes > HD_MAX_CLASSES) return GNM_ERR_TOO_LARGE;
    if (n_graphs == 0) return GNM_OK;
    if (!g_f || !weights || !biases || !c_logit) return GNM_ERR_BAD_ARG;
    HeadsParams p = {};
    p.g_f = g_f; p.ldg = ldg; p.mask = mask; p.c_logit = c_logit;
    p.n_graphs = n_graphs; p.n_layers = n_layers; p.n_feat = n_feat; p.n_classes = n_classes;
    for (int l = 0; l < n_layers; ++l) {
        p.w[l] = weights[l]; p.b[l] = biases[l];
        if (!p.w[l] || !p.b[l]) return GNM_ERR_BAD_ARG;
    }
    int ctas = (n_graphs + HD_WARPS - 1) / HD_WARPS;
    if (ctas > 148 * 2) ctas = 148 * 2;
    gnm_count_launch(GNM_K_OTHER);
    heads_ce_kernel<HD_FWD><<<ctas, HD_WARPS * 32, 0, gnm_cast_stream(stream)>>>(p);
    GNM_RETURN_IF_LAUNCH_FAILED();
    return GNM_OK;
}

extern "C" int gnm_heads_bwd(const float* g_f, int64_t ldg, int n_graphs, int n_layers, int n_feat, int n_classes,
                             const float* const* weights, const float* mask, const float* d_logit, float* d_gf, int64_t ldd,
                             float* const* d_weights, float* const* d_biases, float* workspace, int64_t workspace_floats,
                             unsigned int* counter, gnm_stream_t stream) {
    if (n_graphs < 0 || n_layers < 1 || n_feat < 1 || n_classes < 1) return GNM_ERR_BAD_ARG;
    if (n_layers > HD_MAX_LAYERS || n_classes > HD_MAX_CLASSES) return GNM_ERR_TOO_LARGE;
    if (n_graphs == 0) return GNM_OK;
    if (!g_f || !weights || !d_logit || !d_gf || !d_weights || !d_biases || !workspace || !counter) return GNM_ERR_BAD_ARG;
    const int gsz = n_layers * n_classes * n_feat + n_layers * n_classes;
    const size_t smem = (size_t)HD_WARPS * gsz * sizeof(float);
    if (smem > 200 * 1024) return GNM_ERR_TOO_LARGE;
    int ctas = (n_graphs + HD_WARPS - 1) / HD_WARPS;
    if (ctas > 64) ctas = 64;
    if (workspace_floats < (int64_t)ctas * gsz) return GNM_ERR_BAD_ARG;
    HeadsParams p = {};
    p.g_f = g_f; p.ldg = ldg; p.mask = mask; p.d_logit_in = d_logit; p.d_gf = d_gf; p.ldd = ldd; p.ws = workspace;
    p.counter = counter; p.n_graphs = n_graphs; p.n_layers = n_layers; p.n_feat = n_feat; p.n_classes = n_classes;
    for (int l = 0; l < n_layers; ++l) {
        p.w[l] = weights[l]; p.dw[l] = d_weights[l]; p.db[l] = d_biases[l];
        if (!p.w[l] || !p.dw[l] || !p.db[l]) return GNM_ERR_BAD_ARG;
    }
    cudaError_t e = cudaFuncSetAttribute(heads_ce_kernel<HD_BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    gnm_count_launch(GNM_K_OTHER);
    heads_ce_kernel<HD_BWD><<<ctas, HD_WARPS * 32, smem, gnm_cast_stream(stream)>>>(p);
    GNM_RETURN_IF_LAUNCH_FAILED();
    return GNM_OK;
}

extern "C" int gnm_bce_logits(const float* logits, int64_t n, int64_t n_pos, float grad_scale, double loss_scale,
                              double* loss_acc, float* d_logits, gnm_stream_t stream) {
    if (n < 0 || n_pos < 0 || n_pos > n) return GNM_ERR_BAD_ARG;
    if (n == 0) return GNM_OK;
    if (!logits || !loss_acc) return GNM_ERR_BAD_ARG;
    int64_t blocks = (n + 256 * 8 - 1) / (256 * 8);
    if (blocks > 148 * 4) blocks = 148 * 4;
    gnm_count_launch(GNM_K_OTHER);
    bce_logits_kernel<<<(int)blocks, 256, 0, gnm_cast_stream(stream)>>>(logits, n, n_pos, grad_scale, loss_scale, loss_acc,
                                                                        d_logits);
    GNM_RETURN_IF_LAUNCH_FAILED();
    return GNM_OK;
}

extern "C" int gnm_small_gemm(const float* a, int64_t sam, int64_t sak, const float* b, int64_t sbk, int64_t sbn, float* c,
                              int64_t ldc, int m, int n, int k, int sigmoid_a, float* a_out, int64_t lda_out,
                              const float* dsig_s, int64_t lds, const float* dsig_add, int64_t ldadd, gnm_stream_t stream) {
    if (m < 0 || n < 0 || k < 0) return GNM_ERR_BAD_ARG;
    if (m == 0 || n == 0) return GNM_OK;
    if (!a || !b || !c) return GNM_ERR_BAD_ARG;
    if (sigmoid_a && dsig_s) return GNM_ERR_BAD_ARG;
    cudaStream_t st = gnm_cast_stream(stream);
    const int tiles = ((n + SG_T - 1) / SG_T) * ((m + SG_T - 1) / SG_T);
    // few output tiles and a long reduction (dW = du^T c: 25 tiles, K = B): split k across CTAs, partials merged by atomics
    int splits = 1;
    if (!sigmoid_a && !dsig_s && tiles < 74 && k >= 8 * SG_K) {
        splits = 148 / tiles;
        const int max_splits = k / (4 * SG_K);
        if (splits > max_splits) splits = max_splits;
        if (splits < 1) splits = 1;
    }
    int k_per_split = ((k + splits - 1) / splits + SG_K - 1) / SG_K * SG_K;
    splits = (k + k_per_split - 1) / k_per_split;
    if (splits < 1) { splits = 1; k_per_split = SG_K; }
    if (splits > 1) {
        if (ldc != n) {
            cudaError_t e = cudaMemset2DAsync(c, (size_t)ldc * 4, 0, (size_t)n * 4, (size_t)m, st);
            if (e != cudaSuccess) return (int)e;
        } else {
            cudaError_t e = cudaMemsetAsync(c, 0, (size_t)m * n * 4, st);
            if (e != cudaSuccess) return (int)e;
        }
    }
    dim3 grid((n + SG_T - 1) / SG_T, (m + SG_T - 1) / SG_T, splits);
    gnm_count_launch(GNM_K_OTHER);
    if (sigmoid_a)
        small_gemm_kernel<true, false><<<grid, 256, 0, st>>>(a, sam, sak, b, sbk, sbn, c, ldc, m, n, k, k_per_split, a_out, lda_out,
                                                             nullptr, 0, nullptr, 0);
    else if (dsig_s)
        small_gemm_kernel<false, true><<<grid, 256, 0, st>>>(a, sam, sak, b, sbk, sbn, c, ldc, m, n, k, k_per_split, nullptr, 0,
                                                             dsig_s, lds, dsig_add, ldadd);
    else
        small_gemm_kernel<false, false><<<grid, 256, 0, st>>>(a, sam, sak, b, sbk, sbn, c, ldc, m, n, k, k_per_split, nullptr, 0,
                                                              nullptr, 0, nullptr, 0);
    GNM_RETURN_IF_LAUNCH_FAILED();
    return GNM_OK;
}

extern "C" int gnm_dgi_neg_grad(const int32_t* neg_idx, const float* s2, const float* u, int64_t ldu, int n_graphs, int width,
                                float* d_neg, int64_t ldn, int n_neg, gnm_stream_t stream) {
    if (n_graphs < 0 || width < 0 || n_neg < 0) return GNM_ERR_BAD_ARG;
    if (n_neg == 0 || width == 0) return GNM_OK;
    if (!neg_idx || !s2 || !u || !d_neg) return GNM_ERR_BAD_ARG;
    gnm_count_launch(GNM_K_OTHER);
    dgi_neg_grad_kernel<<<n_neg, 128, 0, gnm_cast_stream(stream)>>>(neg_idx, s2, u, ldu, n_graphs, width, d_neg, ldn);
    GNM_RETURN_IF_LAUNCH_FAILED();
    return GNM_OK;
}

extern "C" int gnm_adam_step(float* const* params, const float* const* grads, const int32_t* numel, const int32_t* state_off,
                             int n_tensors, float* exp_avg, float* exp_avg_sq, float* step, const float* lr, double beta1,
                             double beta2, float eps, float weight_decay, float grad_scale, const double* loss_terms,
                             int n_loss_terms, float* loss_out, gnm_stream_t stream) {
    if (n_tensors < 0 || n_loss_terms < 0) return GNM_ERR_BAD_ARG;
    if (!step || !lr || (n_tensors > 0 && (!params || !grads || !numel || !state_off || !exp_avg || !exp_avg_sq)))
        return GNM_ERR_BAD_ARG;
    if (loss_out != nullptr && n_loss_terms > 0 && !loss_terms) return GNM_ERR_BAD_ARG;
    cudaStream_t st = gnm_cast_stream(stream);
    // tables of AD_MAX_TENSORS tensors per launch; only the last launch advances the step counter
    int done = 0;
    do {
        AdamTable t;
        const int cnt = (n_tensors - done) < AD_MAX_TENSORS ? (n_tensors - done) : AD_MAX_TENSORS;
        t.n_tensors = cnt;
        t.chunk0[0] = 0;
        for (int i = 0; i < AD_MAX_TENSORS; ++i) {
            const bool on = i < cnt;
            t.p[i] = on ? params[done + i] : nullptr;
            t.g[i] = on ? grads[done + i] : nullptr;
            t.numel[i] = on ? numel[done + i] : 0;
            t.state_off[i] = on ? state_off[done + i] : 0;
            if (on && (!t.p[i] || !t.g[i] || t.numel[i] < 0)) return GNM_ERR_BAD_ARG;
            t.chunk0[i + 1] = t.chunk0[i] + (t.numel[i] + AD_CHUNK - 1) / AD_CHUNK;
        }
        int grid = t.chunk0[cnt];
        if (grid > 148 * 4) grid = 148 * 4;
        if (grid < 1) grid = 1;
        const bool last = done + cnt >= n_tensors;
        gnm_count_launch(GNM_K_OTHER);
        adam_kernel<<<grid, 256, 0, st>>>(t, exp_avg, exp_avg_sq, step, lr, beta1, beta2, eps, weight_decay, grad_scale,
                                          loss_terms, last ? n_loss_terms : 0, last ? loss_out : nullptr);
        GNM_RETURN_IF_LAUNCH_FAILED();
        done += cnt;
    } while (done < n_tensors);
    gnm_count_launch(GNM_K_OTHER);
    adam_commit_step_kernel<<<1, 1, 0, st>>>(step);
    GNM_RETURN_IF_LAUNCH_FAILED();
    return GNM_OK;
}
