// Block-diagonal CSR construction and batch assembly (reference: models/graphcnn.py:84-134).
#include "gnm_common.cuh"

namespace {

constexpr int kBuildThreads = 256;

// One CTA per graph. Shared memory (ints): cnt[n_max+1] | cur[n_max] | bins[sort_warps][n_max].
//  1. histogram of source rows (+1 per row for the self loop)          -> row lengths
//  2. block exclusive scan                                             -> row starts (= rowptr)
//  3. scatter destinations to their row (shared cursors; order within a row is arbitrary)
//  4. per-row counting sort over the graph's N column ids (one warp per row) -> canonical
//     row-major order, duplicates kept. Deterministic output regardless of step 3's order.
__global__ void __launch_bounds__(kBuildThreads)
csr_build_kernel(const int64_t* __restrict__ edges, int64_t e_total, const int64_t* __restrict__ edge_off,
                 const int32_t* __restrict__ node_off, int n_graphs, int n_max, int self_loops, int local_cols,
                 int sort_warps, int32_t* __restrict__ rowptr, int32_t* __restrict__ colidx,
                 int32_t* __restrict__ status) {
    extern __shared__ int smem_i[];
    __shared__ int warp_tot[kBuildThreads / 32];
    const int g = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n0 = node_off[g];
    const int n = node_off[g + 1] - n0;
    const int64_t e0 = edge_off[g], e1 = edge_off[g + 1];
    const int64_t out0 = e0 + (self_loops ? (int64_t)n0 : 0);
    int* cnt = smem_i;
    int* cur = smem_i + n_max + 1;
    int* bins = cur + n_max;
    const int64_t* src = edges;
    const int64_t* dst = edges + e_total;

    for (int i = tid; i <= n; i += kBuildThreads) cnt[i] = (self_loops && i < n) ? 1 : 0;
    __syncthreads();
    bool bad = false;
    for (int64_t e = e0 + tid; e < e1; e += kBuildThreads) {
        const int64_t s = src[e], d = dst[e];
        if (s < 0 || s >= n || d < 0 || d >= n) { bad = true; continue; }
        atomicAdd(&cnt[(int)s], 1);
    }
    if (bad) atomicOr(status, 1);
    __syncthreads();

    // exclusive scan of cnt[0..n] (n+1 entries)
    const int per = (n + 1 + kBuildThreads - 1) / kBuildThreads;
    const int lo = min(tid * per, n + 1), hi = min(lo + per, n + 1);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += cnt[i];
    int incl = warp_inclusive_scan(s, lane);
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = (lane < kBuildThreads / 32) ? warp_tot[lane] : 0;
        int wi = warp_inclusive_scan(w, lane);
        if (lane < kBuildThreads / 32) warp_tot[lane] = wi - w;
    }
    __syncthreads();
    int base = warp_tot[warp] + incl - s;
    for (int i = lo; i < hi; ++i) {
        const int c = cnt[i];
        cnt[i] = base;
        base += c;
    }
    __syncthreads();
    for (int i = tid; i < n; i += kBuildThreads) {
        const int p = cnt[i];
        rowptr[n0 + i] = (int32_t)(out0 + p);
        if (self_loops) colidx[out0 + p] = i;   // the self loop takes the first slot of its row
        cur[i] = p + (self_loops ? 1 : 0);
    }
    if (g == n_graphs - 1 && tid == 0) rowptr[n0 + n] = (int32_t)(out0 + cnt[n]);
    __syncthreads();
    for (int64_t e = e0 + tid; e < e1; e += kBuildThreads) {
        const int64_t sr = src[e], d = dst[e];
        if (sr < 0 || sr >= n || d < 0 || d >= n) continue;
        const int p = atomicAdd(&cur[(int)sr], 1);
        colidx[out0 + p] = (int32_t)d;
    }
    __syncthreads();

    const int col_base = local_cols ? 0 : n0;
    if (warp < sort_warps) {
        int* b = bins + (size_t)warp * n_max;
        for (int i = warp; i < n; i += sort_warps) {
            const int rs = cnt[i], len = cnt[i + 1] - rs;
            int32_t* seg = colidx + out0 + rs;
            if (len == 0) continue;
            if (len == 1) {
                if (lane == 0) seg[0] += col_base;
                continue;
            }
            for (int c = lane; c < n; c += 32) b[c] = 0;
            __syncwarp();
            for (int q = lane; q < len; q += 32) atomicAdd(&b[seg[q]], 1);
            __syncwarp();
            int off = 0;
            for (int c0 = 0; c0 < n; c0 += 32) {
                const int c = (c0 + lane < n) ? b[c0 + lane] : 0;
                const int inc = warp_inclusive_scan(c, lane);
                const int exc = inc - c;
                for (int t = 0; t < c; ++t) seg[off + exc + t] = c0 + lane + col_base;
                off += __shfl_sync(GNM_FULL_MASK, inc, 31);
            }
            __syncwarp();
        }
    }
}

// One CTA per batch slot: copy a stored graph's local CSR into the batch CSR.
__global__ void __launch_bounds__(256)
csr_batch_gather_kernel(const int64_t* __restrict__ src_rp_addr, const int64_t* __restrict__ src_ci_addr,
                        const int64_t* __restrict__ src_tag_addr, const int32_t* __restrict__ node_off,
                        const int64_t* __restrict__ nnz_off, int n_graphs, int32_t* __restrict__ rowptr,
                        int32_t* __restrict__ colidx, int32_t* __restrict__ tags) {
    const int g = blockIdx.x;
    const int32_t* rp = reinterpret_cast<const int32_t*>(src_rp_addr[g]);
    const int32_t* ci = reinterpret_cast<const int32_t*>(src_ci_addr[g]);
    const int n0 = node_off[g], n = node_off[g + 1] - n0;
    const int64_t z0 = nnz_off[g];
    const int base = rp[0];
    const int nnz = rp[n] - base;
    const int part = blockIdx.y, parts = gridDim.y;
    if (part == 0) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) rowptr[n0 + i] = (int32_t)(rp[i] - base + z0);
        if (g == n_graphs - 1 && threadIdx.x == 0) rowptr[n0 + n] = (int32_t)(nnz + z0);
        if (tags != nullptr && src_tag_addr != nullptr) {
            const int32_t* tg = reinterpret_cast<const int32_t*>(src_tag_addr[g]);
            for (int i = threadIdx.x; i < n; i += blockDim.x) tags[n0 + i] = tg[i];
        }
    }
    if (colidx == nullptr) return;       // structure-only gather: the dense-block kernels read the stored bitmaps
    const int chunk = (nnz + parts - 1) / parts;
    const int q0 = part * chunk, q1 = min(nnz, q0 + chunk);
    const int32_t* s = ci;   // address of this graph's own column segment
    int32_t* d = colidx + z0;
    for (int q = q0 + threadIdx.x; q < q1; q += blockDim.x) d[q] = s[q] + n0;
}

}  // namespace

extern "C" int gnm_csr_build(const int64_t* edges, int64_t e_total, const int64_t* edge_off, const int32_t* node_off,
                             int n_graphs, int n_max, int add_self_loops, int local_cols, int32_t* rowptr,
                             int32_t* colidx, int32_t* status, gnm_stream_t stream) {
    if (n_graphs < 0 || n_max < 0 || e_total < 0) return GNM_ERR_BAD_ARG;
    if (n_graphs == 0) return GNM_OK;
    if (!edge_off || !node_off || !rowptr || !status || (e_total > 0 && (!edges || !colidx))) return GNM_ERR_BAD_ARG;
    int dev = 0;
    cudaGetDevice(&dev);
    int smem_cap = 0;
    cudaDeviceGetAttribute(&smem_cap, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const int64_t fixed = (int64_t)(2 * (int64_t)n_max + 1) * 4;
    int64_t avail = (int64_t)smem_cap - 1024 - fixed;
    int sort_warps = n_max > 0 ? (int)(avail / ((int64_t)n_max * 4)) : kBuildThreads / 32;
    if (sort_warps < 1) return GNM_ERR_TOO_LARGE;
    if (sort_warps > kBuildThreads / 32) sort_warps = kBuildThreads / 32;
    const size_t smem = (size_t)(fixed + (int64_t)sort_warps * n_max * 4);
    cudaError_t e = cudaFuncSetAttribute(csr_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    gnm_count_launch(GNM_K_OTHER);
    csr_build_kernel<<<n_graphs, kBuildThreads, smem, gnm_cast_stream(stream)>>>(
        edges, e_total, edge_off, node_off, n_graphs, n_max, add_self_loops, local_cols, sort_warps, rowptr, colidx,
        status);
    GNM_RETURN_IF_LAUNCH_FAILED();
    return GNM_OK;
}

extern "C" int gnm_csr_batch_gather(const int64_t* src_rowptr_addr, const int64_t* src_colidx_addr,
                                    const int64_t* src_tag_addr, const int32_t* node_off, const int64_t* nnz_off,
                                    int n_graphs, int32_t* rowptr, int32_t* colidx, int32_t* tags,
                                    gnm_stream_t stream) {
    if (n_graphs < 0) return GNM_ERR_BAD_ARG;
    if (n_graphs == 0) return GNM_OK;
    if (!src_rowptr_addr || !src_colidx_addr || !node_off || !nnz_off || !rowptr) return GNM_ERR_BAD_ARG;
    // a handful of CTAs per graph so that small batches still fill the 148 SMs
    int parts = 1;
    if (n_graphs < 1184) parts = (1184 + n_graphs - 1) / n_graphs;
    if (parts > 16) parts = 16;
    if (colidx == nullptr) parts = 1;
    dim3 grid(n_graphs, parts);
    gnm_count_launch(GNM_K_OTHER);
    csr_batch_gather_kernel<<<grid, 256, 0, gnm_cast_stream(stream)>>>(src_rowptr_addr, src_colidx_addr, src_tag_addr,
                                                                       node_off, nnz_off, n_graphs, rowptr, colidx,
                                                                       tags);
    GNM_RETURN_IF_LAUNCH_FAILED();
    return GNM_OK;
}
