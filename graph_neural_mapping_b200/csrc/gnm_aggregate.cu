// Neighbour aggregation over the block-diagonal CSR (reference: models/graphcnn.py:154-161, :178-182)
// and the small row-wise helpers that go with it (eps gradient, layer-0 table gradient).
#include "gnm_common.cuh"

namespace {

template <int VEC>
struct Vec;
template <>
struct Vec<4> {
    float4 v;
    __device__ __forceinline__ void zero() { v = make_float4(0.f, 0.f, 0.f, 0.f); }
    __device__ __forceinline__ void load(const float* p) { v = __ldg(reinterpret_cast<const float4*>(p)); }
    __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = v; }
    __device__ __forceinline__ void fma(float w, const Vec<4>& o) {
        v.x = fmaf(w, o.v.x, v.x); v.y = fmaf(w, o.v.y, v.y); v.z = fmaf(w, o.v.z, v.z); v.w = fmaf(w, o.v.w, v.w);
    }
    __device__ __forceinline__ void div(float d) { v.x /= d; v.y /= d; v.z /= d; v.w /= d; }
    __device__ __forceinline__ void xor_add(int o) {
        v.x += __shfl_xor_sync(GNM_FULL_MASK, v.x, o); v.y += __shfl_xor_sync(GNM_FULL_MASK, v.y, o);
        v.z += __shfl_xor_sync(GNM_FULL_MASK, v.z, o); v.w += __shfl_xor_sync(GNM_FULL_MASK, v.w, o);
    }
};
template <>
struct Vec<1> {
    float v;
    __device__ __forceinline__ void zero() { v = 0.f; }
    __device__ __forceinline__ void load(const float* p) { v = __ldg(p); }
    __device__ __forceinline__ void store(float* p) const { *p = v; }
    __device__ __forceinline__ void fma(float w, const Vec<1>& o) { v = fmaf(w, o.v, v); }
    __device__ __forceinline__ void div(float d) { v /= d; }
    __device__ __forceinline__ void xor_add(int o) { v += __shfl_xor_sync(GNM_FULL_MASK, v, o); }
};

// Warp per destination row. The warp is split into G = 32/LPR groups of LPR lanes; each group
// owns one neighbour at a time and its lanes cover LPR*VEC consecutive features of that
// neighbour's row with one (128-bit when VEC == 4) coalesced load. Column indices are read
// 32 at a time (coalesced) and broadcast with shuffles; four neighbours per group are in
// flight per iteration. blockIdx.y tiles feature columns when F > LPR*VEC.
template <int VEC, int LPR>
__global__ void __launch_bounds__(256)
aggregate_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, int n_rows,
                 const float* __restrict__ src, int64_t ld_src, const int32_t* __restrict__ src_map,
                 float* __restrict__ dst, int64_t ld_dst, int n_feat, int mode, const float* __restrict__ eps,
                 const float* __restrict__ bias) {
    constexpr int G = 32 / LPR;
    const int row = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
    if (row >= n_rows) return;
    const int lane = threadIdx.x & 31;
    const int grp = lane / LPR, sub = lane % LPR;
    const int col = blockIdx.y * (LPR * VEC) + sub * VEC;
    const bool active = col < n_feat;
    const int cc = active ? col : 0;
    const int s = rowptr[row], e = rowptr[row + 1];
    Vec<VEC> acc;
    acc.zero();
    for (int k = s; k < e; k += 32) {
        const int idx = (k + lane < e) ? colidx[k + lane] : -1;
        const int cnt = min(32, e - k);
        for (int t = 0; t < cnt; t += 4 * G) {
            int j[4];
            float w[4];
            Vec<VEC> x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int sl = t + u * G + grp;
                j[u] = __shfl_sync(GNM_FULL_MASK, idx, sl & 31);
                if (sl >= 32) j[u] = -1;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int jj = max(j[u], 0);
                w[u] = (j[u] >= 0) ? 1.f : 0.f;
                if (mode == 2) {
                    const int dj = rowptr[jj + 1] - rowptr[jj];
                    w[u] = (j[u] >= 0) ? 1.f / (float)dj : 0.f;
                }
                const int64_t r = src_map ? (int64_t)src_map[jj] : (int64_t)jj;
                x[u].load(src + r * ld_src + cc);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) acc.fma(w[u], x[u]);
        }
    }
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) acc.xor_add(o);
    if (grp == 0 && active) {
        if (mode == 1) acc.div((float)(e - s));
        if (eps != nullptr) {
            const float c = 1.f + __ldg(eps);
            const int64_t r = src_map ? (int64_t)src_map[row] : (int64_t)row;
            Vec<VEC> self;
            self.load(src + r * ld_src + col);
            acc.fma(c, self);
        }
        if (bias != nullptr) {
            Vec<VEC> bv;
            bv.load(bias + col);
            acc.fma(1.f, bv);
        }
        acc.store(dst + (int64_t)row * ld_dst + col);
    }
}

template <int VEC, int LPR>
int launch_aggregate(const int32_t* rowptr, const int32_t* colidx, int n_rows, const float* src, int64_t ld_src,
                     const int32_t* src_map, float* dst, int64_t ld_dst, int n_feat, int mode, const float* eps,
                     const float* bias, cudaStream_t st) {
    const int rows_per_block = 8;
    dim3 grid((n_rows + rows_per_block - 1) / rows_per_block, (n_feat + LPR * VEC - 1) / (LPR * VEC));
    gnm_count_launch(GNM_K_AGG_CSR);
    aggregate_kernel<VEC, LPR><<<grid, rows_per_block * 32, 0, st>>>(rowptr, colidx, n_rows, src, ld_src, src_map, dst,
                                                                      ld_dst, n_feat, mode, eps, bias);
    GNM_RETURN_IF_LAUNCH_FAILED();
    return GNM_OK;
}

// sum_i <a[i], b[map(i)]>: one warp per row slice, block reduce, one double atomic per block.
__global__ void __launch_bounds__(256)
dot_rows_kernel(const float* __restrict__ a, int64_t lda, const float* __restrict__ b, int64_t ldb,
                const int32_t* __restrict__ b_map, int n_rows, int n_feat, double* __restrict__ out) {
    __shared__ double part[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wg = blockIdx.x * 8 + warp, nw = gridDim.x * 8;
    float acc = 0.f;
    double dacc = 0.0;
    for (int r = wg; r < n_rows; r += nw) {
        const int64_t rb = b_map ? (int64_t)b_map[r] : (int64_t)r;
        const float* pa = a + (int64_t)r * lda;
        const float* pb = b + rb * ldb;
        for (int c = lane; c < n_feat; c += 32) acc = fmaf(pa[c], __ldg(pb + c), acc);
        dacc += (double)acc;   // flush the fp32 partial every row: keeps long sums accurate
        acc = 0.f;
    }
    dacc = warp_sum_d(dacc);
    if (lane == 0) part[warp] = dacc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += part[i];
        atomicAdd(out, t);
    }
}

// table_grad[tags[r], :] += g[r, :]. Each CTA owns a contiguous slab of rows and a feature
// chunk; it accumulates into a shared-memory copy of the (table rows x chunk) tile and then
// flushes once with global atomics: (#CTAs x tile) atomics instead of one per element of g.
__global__ void __launch_bounds__(256)
scatter_rows_add_kernel(const float* __restrict__ g, int64_t ldg, const int32_t* __restrict__ tags, int n_rows,
                        int n_feat, float* __restrict__ table_grad, int64_t ldt, int n_table_rows, int fchunk,
                        int rows_per_cta, float* __restrict__ workspace) {
    extern __shared__ __align__(16) float tile[];   // [n_table_rows][fchunk]
    const int f0 = blockIdx.y * fchunk;
    const int fw = min(fchunk, n_feat - f0);
    for (int i = threadIdx.x; i < n_table_rows * fchunk; i += blockDim.x) tile[i] = 0.f;
    __syncthreads();
    const int r0 = blockIdx.x * rows_per_cta, r1 = min(n_rows, r0 + rows_per_cta);
    // threads cover (row, feature) pairs with features fastest: coalesced reads of g
    const int tpr = min(fw, (int)blockDim.x);          // threads per row
    const int rstep = blockDim.x / tpr;
    const int tf = threadIdx.x % tpr, tr = threadIdx.x / tpr;
    if (tr < rstep) {
        for (int r = r0 + tr; r < r1; r += rstep) {
            const int t = tags[r];
            if (t < 0 || t >= n_table_rows) continue;
            for (int f = tf; f < fw; f += tpr) atomicAdd(&tile[t * fchunk + f], g[(int64_t)r * ldg + f0 + f]);
        }
    }
    __syncthreads();
    if (workspace != nullptr) {
        // deterministic two-stage merge: this CTA's tile goes to its own workspace slice (plain stores)
        float* mine = workspace + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (size_t)n_table_rows * fchunk;
        for (int i = threadIdx.x; i < n_table_rows * fchunk; i += blockDim.x) mine[i] = tile[i];
        return;
    }
    for (int i = threadIdx.x; i < n_table_rows * fchunk; i += blockDim.x) {
        const int t = i / fchunk, f = i % fchunk;
        const float v = tile[i];
        if (f < fw && v != 0.f) atomicAdd(&table_grad[(int64_t)t * ldt + f0 + f], v);
    }
}

// second stage: table_grad[t, f0 + f] += sum over the CTAs' partial tiles (fixed order: deterministic)
__global__ void __launch_bounds__(256)
scatter_rows_merge_kernel(const float* __restrict__ workspace, int n_ctas, int n_table_rows, int fchunk, int n_feat,
                          float* __restrict__ table_grad, int64_t ldt) {
    const int fpart = blockIdx.y;
    const int f0 = fpart * fchunk;
    const int fw = min(fchunk, n_feat - f0);
    const size_t tile = (size_t)n_table_rows * fchunk;
    const float* base = workspace + (size_t)fpart * n_ctas * tile;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_table_rows * fchunk; i += gridDim.x * blockDim.x) {
        const int t = i / fchunk, f = i % fchunk;
        if (f >= fw) continue;
        float a = 0.f;
        for (int c = 0; c < n_ctas; ++c) a += base[(size_t)c * tile + i];
        table_grad[(int64_t)t * ldt + f0 + f] += a;
    }
}

// out[tag[t], :] += sum_k g[k * period + t, :]: the gathered-table gradient when every graph of the batch carries the
// same injective tag sequence (util.py:106-116: one tag per ROI, the same ROI order in every subject). Two deterministic stages: each (row tile, split)
// CTA sums its share of the graphs with 128-bit loads, eight rows in flight per thread; the splits are then added
// in a fixed order.
__global__ void __launch_bounds__(256)
rows_period_sum_kernel(const float* __restrict__ g, int64_t ldg, int n_periods, int period, int n_feat4,
                       float* __restrict__ partial) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;         // (t, c4)
    if (e >= period * n_feat4) return;
    const int t = e / n_feat4, c4 = e - t * n_feat4;
    const int split = blockIdx.y, n_splits = gridDim.y;
    const int k0 = (int)((int64_t)n_periods * split / n_splits), k1 = (int)((int64_t)n_periods * (split + 1) / n_splits);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* base = g + (int64_t)t * ldg + c4 * 4;
    const int64_t step = (int64_t)period * ldg;
    int k = k0;
    for (; k + 8 <= k1; k += 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(base + (int64_t)(k + u) * step));
#pragma unroll
        for (int u = 0; u < 8; ++u) { a.x += v[u].x; a.y += v[u].y; a.z += v[u].z; a.w += v[u].w; }
    }
    for (; k < k1; ++k) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(base + (int64_t)k * step));
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    reinterpret_cast<float4*>(partial)[(int64_t)split * period * n_feat4 + e] = a;
}

__global__ void __launch_bounds__(256)
rows_period_merge_kernel(const float* __restrict__ partial, int n_splits, int period, int n_feat4,
                         const int32_t* __restrict__ tags, int n_table_rows, float* __restrict__ out, int64_t ldo) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= period * n_feat4) return;
    const int t = e / n_feat4, c4 = e - t * n_feat4;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < n_splits; ++s) {
        const float4 v = reinterpret_cast<const float4*>(partial)[(int64_t)s * period * n_feat4 + e];
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    const int row = tags != nullptr ? tags[t] : t;
    if (row < 0 || row >= n_table_rows) return;
    float* o = out + (int64_t)row * ldo + c4 * 4;
    o[0] += a.x; o[1] += a.y; o[2] += a.z; o[3] += a.w;
}

// ---- max pooling over neighbours (graphcnn.py:55-81, 137-143) ----------------------------------------------------
// The reference gathers a padded neighbour list (pads point at a dummy row = column-wise minimum of h over the whole
// batch) and takes torch.max over it. Equivalent here: maximum over the CSR row (which holds the self loop when the
// reference appends the node itself, i.e. learn_eps == False); a row without any entry yields the dummy row. The
// arg-max (source row, or n_rows for the dummy) is kept for the backward pass. Ties keep the lowest column id - they
// only occur at exact zeros behind a ReLU (whose backward masks them) or on measure-zero inputs.

__device__ __forceinline__ unsigned int float_order_key(float v) {
    const unsigned int u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float float_from_order_key(unsigned int k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// packed[f] = min over rows of (order_key(h[r, f]) << 32 | r): column minimum and its first row. Caller presets ~0.
__global__ void __launch_bounds__(128)
col_min_kernel(const float* __restrict__ h, int64_t ldh, int n_rows, int n_feat, int rows_per_cta,
               unsigned long long* __restrict__ packed) {
    const int f = blockIdx.y * blockDim.x + threadIdx.x;
    if (f >= n_feat) return;
    const int r0 = blockIdx.x * rows_per_cta, r1 = min(n_rows, r0 + rows_per_cta);
    unsigned long long best = ~0ull;
    for (int r = r0; r < r1; ++r) {
        const unsigned long long k = ((unsigned long long)float_order_key(h[(int64_t)r * ldh + f]) << 32) | (unsigned int)r;
        best = k < best ? k : best;
    }
    if (best != ~0ull) atomicMin(&packed[f], best);
}

__global__ void __launch_bounds__(256)
aggregate_max_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, int n_rows,
                     const float* __restrict__ h, int64_t ldh, int n_feat, const unsigned long long* __restrict__ col_min,
                     const float* __restrict__ eps, float* __restrict__ out, int64_t ld_out, int32_t* __restrict__ argmax) {
    const int row = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
    if (row >= n_rows) return;
    const int lane = threadIdx.x & 31;
    const int s = rowptr[row], e = rowptr[row + 1];
    const float self_c = eps != nullptr ? 1.f + __ldg(eps) : 0.f;
    for (int f0 = 0; f0 < n_feat; f0 += 64) {
        const int fa = f0 + lane, fb = f0 + lane + 32;
        const bool oka = fa < n_feat, okb = fb < n_feat;
        float ba = -INFINITY, bb = -INFINITY;
        int ia = -1, ib = -1;
        for (int k = s; k < e; ++k) {
            const int j = colidx[k];
            const float* hj = h + (int64_t)j * ldh;
            const float va = oka ? hj[fa] : -INFINITY, vb = okb ? hj[fb] : -INFINITY;
            if (va > ba || ia < 0) { if (oka) { ba = va; ia = j; } }
            if (vb > bb || ib < 0) { if (okb) { bb = vb; ib = j; } }
        }
        if (oka) {
            if (ia < 0) { ba = float_from_order_key((unsigned int)(col_min[fa] >> 32)); ia = n_rows; }
            if (eps != nullptr) ba = fmaf(self_c, h[(int64_t)row * ldh + fa], ba);
            out[(int64_t)row * ld_out + fa] = ba;
            argmax[(int64_t)row * n_feat + fa] = ia;
        }
        if (okb) {
            if (ib < 0) { bb = float_from_order_key((unsigned int)(col_min[fb] >> 32)); ib = n_rows; }
            if (eps != nullptr) bb = fmaf(self_c, h[(int64_t)row * ldh + fb], bb);
            out[(int64_t)row * ld_out + fb] = bb;
            argmax[(int64_t)row * n_feat + fb] = ib;
        }
    }
}

// d_h[j, f] = sum over the rows i of CSR row j (symmetric structure: i lists j iff j lists i) whose arg-max at f is j,
// + (1 + eps) d_out[j, f]. Pull form: deterministic, no atomics.
__global__ void __launch_bounds__(256)
aggregate_max_bwd_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, int n_rows,
                         const float* __restrict__ d_out, int64_t ld_dout, int n_feat,
                         const int32_t* __restrict__ argmax, const float* __restrict__ eps, float* __restrict__ d_h,
                         int64_t ld_dh) {
    const int row = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
    if (row >= n_rows) return;
    const int lane = threadIdx.x & 31;
    const int s = rowptr[row], e = rowptr[row + 1];
    const float self_c = eps != nullptr ? 1.f + __ldg(eps) : 0.f;
    for (int f0 = 0; f0 < n_feat; f0 += 64) {
        const int fa = f0 + lane, fb = f0 + lane + 32;
        const bool oka = fa < n_feat, okb = fb < n_feat;
        float aa = 0.f, ab = 0.f;
        for (int k = s; k < e; ++k) {
            const int i = colidx[k];
            const int32_t* am = argmax + (int64_t)i * n_feat;
            const float* di = d_out + (int64_t)i * ld_dout;
            if (oka && am[fa] == row) aa += di[fa];
            if (okb && am[fb] == row) ab += di[fb];
        }
        if (oka) d_h[(int64_t)row * ld_dh + fa] = fmaf(self_c, d_out[(int64_t)row * ld_dout + fa], aa);
        if (okb) d_h[(int64_t)row * ld_dh + fb] = fmaf(self_c, d_out[(int64_t)row * ld_dout + fb], ab);
    }
}

// rows that took the dummy (no neighbour at all): their gradient belongs to the row that supplied the column minimum
__global__ void __launch_bounds__(256)
aggregate_max_dummy_bwd_kernel(int n_rows, const float* __restrict__ d_out, int64_t ld_dout, int n_feat,
                               const int32_t* __restrict__ argmax, const unsigned long long* __restrict__ col_min,
                               float* __restrict__ d_h, int64_t ld_dh) {
    const int64_t total = (int64_t)n_rows * n_feat;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        if (argmax[idx] != n_rows) continue;
        const int i = (int)(idx / n_feat), f = (int)(idx - (int64_t)i * n_feat);
        const int r = (int)(col_min[f] & 0xffffffffull);
        atomicAdd(&d_h[(int64_t)r * ld_dh + f], d_out[(int64_t)i * ld_dout + f]);
    }
}

}  // namespace

extern "C" int gnm_aggregate(const int32_t* rowptr, const int32_t* colidx, int n_rows, const float* src,
                             int64_t ld_src, const int32_t* src_map, float* dst, int64_t ld_dst, int n_feat, int mode,
                             const float* eps, const float* bias, gnm_stream_t stream) {
    if (n_rows < 0 || n_feat < 0 || mode < 0 || mode > 2) return GNM_ERR_BAD_ARG;
    if (n_rows == 0 || n_feat == 0) return GNM_OK;
    if (!rowptr || !src || !dst) return GNM_ERR_BAD_ARG;
    cudaStream_t st = gnm_cast_stream(stream);
    const bool vec4 = (n_feat % 4 == 0) && (ld_src % 4 == 0) && (ld_dst % 4 == 0) && gnm_aligned16(src) &&
                      gnm_aligned16(dst) && (bias == nullptr || gnm_aligned16(bias));
#define GNM_AGG(V, L) \
    return launch_aggregate<V, L>(rowptr, colidx, n_rows, src, ld_src, src_map, dst, ld_dst, n_feat, mode, eps, bias, st)
    if (vec4) {
        const int q = n_feat / 4;
        if (q <= 1) GNM_AGG(4, 1);
        if (q <= 2) GNM_AGG(4, 2);
        if (q <= 4) GNM_AGG(4, 4);
        if (q <= 8) GNM_AGG(4, 8);
        if (q <= 16) GNM_AGG(4, 16);
        GNM_AGG(4, 32);
    } else {
        if (n_feat <= 2) GNM_AGG(1, 2);
        if (n_feat <= 4) GNM_AGG(1, 4);
        if (n_feat <= 8) GNM_AGG(1, 8);
        if (n_feat <= 16) GNM_AGG(1, 16);
        GNM_AGG(1, 32);
    }
#undef GNM_AGG
}

extern "C" int gnm_dot_rows(const float* a, int64_t lda, const float* b, int64_t ldb, const int32_t* b_map,
                            int n_rows, int n_feat, double* out, gnm_stream_t stream) {
    if (n_rows < 0 || n_feat < 0) return GNM_ERR_BAD_ARG;
    if (n_rows == 0 || n_feat == 0) return GNM_OK;
    if (!a || !b || !out) return GNM_ERR_BAD_ARG;
    int blocks = (n_rows + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    gnm_count_launch(GNM_K_OTHER);
    dot_rows_kernel<<<blocks, 256, 0, gnm_cast_stream(stream)>>>(a, lda, b, ldb, b_map, n_rows, n_feat, out);
    GNM_RETURN_IF_LAUNCH_FAILED();
    return GNM_OK;
}

extern "C" int gnm_scatter_rows_add(const float* g, int64_t ldg, const int32_t* tags, int n_rows, int n_feat,
                                    float* table_grad, int64_t ldt, int n_table_rows, float* workspace,
                                    int64_t workspace_floats, gnm_stream_t stream) {
    if (n_rows < 0 || n_feat < 0 || n_table_rows < 0) return GNM_ERR_BAD_ARG;
    if (n_rows == 0 || n_feat == 0 || n_table_rows == 0) return GNM_OK;
    if (!g || !tags || !table_grad) return GNM_ERR_BAD_ARG;
    // feature chunk so that the tile fits in ~96 KB of shared memory
    int fchunk = n_feat;
    while ((int64_t)n_table_rows * fchunk * 4 > 96 * 1024 && fchunk > 1) fchunk = (fchunk + 1) / 2;
    if ((int64_t)n_table_rows * fchunk * 4 > 200 * 1024) return GNM_ERR_TOO_LARGE;
    const int fparts = (n_feat + fchunk - 1) / fchunk;
    int ctas = 148 * 2 / fparts;
    if (ctas < 1) ctas = 1;
    int rows_per_cta = (n_rows + ctas - 1) / ctas;
    if (rows_per_cta < 64) rows_per_cta = 64;
    ctas = (n_rows + rows_per_cta - 1) / rows_per_cta;
    const size_t smem = (size_t)n_table_rows * fchunk * 4;
    cudaError_t e = cudaFuncSetAttribute(scatter_rows_add_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int64_t need = (int64_t)ctas * fparts * n_table_rows * fchunk;
    float* ws = (workspace != nullptr && workspace_floats >= need) ? workspace : nullptr;
    dim3 grid(ctas, fparts);
    gnm_count_launch(GNM_K_OTHER);
    scatter_rows_add_kernel<<<grid, 256, smem, gnm_cast_stream(stream)>>>(g, ldg, tags, n_rows, n_feat, table_grad, ldt,
                                                                          n_table_rows, fchunk, rows_per_cta, ws);
    GNM_RETURN_IF_LAUNCH_FAILED();
    if (ws != nullptr) {
        int mb = (n_table_rows * fchunk + 255) / 256;
        if (mb > 148 * 4) mb = 148 * 4;
        dim3 mgrid(mb, fparts);
        gnm_count_launch(GNM_K_OTHER);
        scatter_rows_merge_kernel<<<mgrid, 256, 0, gnm_cast_stream(stream)>>>(ws, ctas, n_table_rows, fchunk, n_feat,
                                                                              table_grad, ldt);
        GNM_RETURN_IF_LAUNCH_FAILED();
    }
    return GNM_OK;
}

/* Workspace size (floats) for the deterministic two-stage path of gnm_scatter_rows_add. */
extern "C" int64_t gnm_scatter_rows_workspace(int n_rows, int n_feat, int n_table_rows) {
    if (n_rows <= 0 || n_feat <= 0 || n_table_rows <= 0) return 0;
    int fchunk = n_feat;
    while ((int64_t)n_table_rows * fchunk * 4 > 96 * 1024 && fchunk > 1) fchunk = (fchunk + 1) / 2;
    const int fparts = (n_feat + fchunk - 1) / fchunk;
    int ctas = 148 * 2 / fparts;
    if (ctas < 1) ctas = 1;
    int rows_per_cta = (n_rows + ctas - 1) / ctas;
    if (rows_per_cta < 64) rows_per_cta = 64;
    ctas = (n_rows + rows_per_cta - 1) / rows_per_cta;
    return (int64_t)ctas * fparts * n_table_rows * fchunk;
}

/* Splits (workspace = splits * period * n_feat floats) used by gnm_rows_period_sum. */
static int rows_period_splits(int n_periods) { return n_periods < 32 ? (n_periods < 1 ? 1 : n_periods) : 32; }

extern "C" int64_t gnm_rows_period_workspace(int n_rows, int n_feat, int period) {
    if (n_rows <= 0 || n_feat <= 0 || period <= 0) return 0;
    return (int64_t)rows_period_splits(n_rows / period) * period * n_feat;
}

extern "C" int gnm_rows_period_sum(const float* g, int64_t ldg, int n_rows, int n_feat, int period, const int32_t* tags,
                                   float* out, int64_t ldo, int n_table_rows, float* workspace,
                                   int64_t workspace_floats, gnm_stream_t stream) {
    if (n_rows < 0 || n_feat < 0 || period <= 0 || n_rows % period != 0 || n_table_rows < 0) return GNM_ERR_BAD_ARG;
    if (n_rows == 0 || n_feat == 0) return GNM_OK;
    if (!g || !out || !workspace) return GNM_ERR_BAD_ARG;
    if ((n_feat & 3) || (ldg & 3) || !gnm_aligned16(g) || !gnm_aligned16(workspace)) return GNM_ERR_ALIGN;
    if (workspace_floats < gnm_rows_period_workspace(n_rows, n_feat, period)) return GNM_ERR_BAD_ARG;
    const int n_periods = n_rows / period, splits = rows_period_splits(n_periods), f4 = n_feat / 4;
    const int blocks = (period * f4 + 255) / 256;
    gnm_count_launch(GNM_K_OTHER);
    rows_period_sum_kernel<<<dim3(blocks, splits), 256, 0, gnm_cast_stream(stream)>>>(g, ldg, n_periods, period, f4, workspace);
    GNM_RETURN_IF_LAUNCH_FAILED();
    gnm_count_launch(GNM_K_OTHER);
    rows_period_merge_kernel<<<blocks, 256, 0, gnm_cast_stream(stream)>>>(workspace, splits, period, f4, tags, n_table_rows, out,
                                                                           ldo);
    GNM_RETURN_IF_LAUNCH_FAILED();
    return GNM_OK;
}

extern "C" int gnm_col_min(const float* h, int64_t ldh, int n_rows, int n_feat, unsigned long long* packed,
                           gnm_stream_t stream) {
    if (n_rows < 0 || n_feat < 0) return GNM_ERR_BAD_ARG;
    if (n_rows == 0 || n_feat == 0) return GNM_OK;
    if (!h || !packed) return GNM_ERR_BAD_ARG;
    const int rows_per_cta = 256;
    dim3 grid((n_rows + rows_per_cta - 1) / rows_per_cta, (n_feat + 127) / 128);
    gnm_count_launch(GNM_K_OTHER);
    col_min_kernel<<<grid, 128, 0, gnm_cast_stream(stream)>>>(h, ldh, n_rows, n_feat, rows_per_cta, packed);
    GNM_RETURN_IF_LAUNCH_FAILED();
    return GNM_OK;
}

extern "C" int gnm_aggregate_max(const int32_t* rowptr, const int32_t* colidx, int n_rows, const float* h, int64_t ldh,
                                 int n_feat, const unsigned long long* col_min, const float* eps, float* out,
                                 int64_t ld_out, int32_t* argmax, gnm_stream_t stream) {
    if (n_rows < 0 || n_feat < 0) return GNM_ERR_BAD_ARG;
    if (n_rows == 0 || n_feat == 0) return GNM_OK;
    if (!rowptr || !h || !col_min || !out || !argmax) return GNM_ERR_BAD_ARG;     // colidx may be NULL when nnz == 0
    gnm_count_launch(GNM_K_OTHER);
    aggregate_max_kernel<<<(n_rows + 7) / 8, 256, 0, gnm_cast_stream(stream)>>>(rowptr, colidx, n_rows, h, ldh, n_feat,
                                                                               col_min, eps, out, ld_out, argmax);
    GNM_RETURN_IF_LAUNCH_FAILED();
    return GNM_OK;
}

extern "C" int gnm_aggregate_max_bwd(const int32_t* rowptr, const int32_t* colidx, int n_rows, const float* d_out,
                                     int64_t ld_dout, int n_feat, const int32_t* argmax,
                                     const unsigned long long* col_min, const float* eps, float* d_h, int64_t ld_dh,
                                     gnm_stream_t stream) {
    if (n_rows < 0 || n_feat < 0) return GNM_ERR_BAD_ARG;
    if (n_rows == 0 || n_feat == 0) return GNM_OK;
    if (!rowptr || !d_out || !argmax || !col_min || !d_h) return GNM_ERR_BAD_ARG;   // colidx may be NULL when nnz == 0
    gnm_count_launch(GNM_K_OTHER);
    aggregate_max_bwd_kernel<<<(n_rows + 7) / 8, 256, 0, gnm_cast_stream(stream)>>>(rowptr, colidx, n_rows, d_out, ld_dout,
                                                                                   n_feat, argmax, eps, d_h, ld_dh);
    GNM_RETURN_IF_LAUNCH_FAILED();
    gnm_count_launch(GNM_K_OTHER);
    aggregate_max_dummy_bwd_kernel<<<148 * 4, 256, 0, gnm_cast_stream(stream)>>>(n_rows, d_out, ld_dout, n_feat, argmax,
                                                                                col_min, d_h, ld_dh);
    GNM_RETURN_IF_LAUNCH_FAILED();
    return GNM_OK;
}
