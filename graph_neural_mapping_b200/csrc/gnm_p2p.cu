// Peer-memory exchange buffers (CUDA IPC) and the standalone small all-reduce; the fused users are gnm_bn_finalize
// (gnm_mlp.cu) and gnm_bn_bwd_coeffs (gnm_mlp_bwd.cu). Protocol: gnm_p2p.cuh.
#include "gnm_p2p.cuh"

namespace {

constexpr size_t P2P_BYTES = (size_t)2 * GNM_P2P_MAX_WORLD * GNM_P2P_MAX_DOUBLES * sizeof(double) +
                             (size_t)2 * GNM_P2P_MAX_WORLD * sizeof(unsigned int) + 256;

__global__ void __launch_bounds__(256) p2p_allreduce_kernel(double* data, int n, const P2PArgs a) {
    p2p_allreduce_block(data, n, a);
}

// dst[r] + byte_offset <- src for every peer region r (remote stores over NVLink; the local copy is a plain store)
__global__ void __launch_bounds__(256) p2p_push_kernel(const float4* __restrict__ src, int64_t n, void* const* __restrict__ regions,
                                                       int world, int64_t byte_offset) {
    const int64_t n4 = n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = src[i];
        for (int r = 0; r < world; ++r)
            reinterpret_cast<float4*>(reinterpret_cast<char*>(regions[r]) + byte_offset)[i] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {          // tail
        const int64_t e = n4 * 4 + threadIdx.x;
        const float v = reinterpret_cast<const float*>(src)[e];
        for (int r = 0; r < world; ++r)
            reinterpret_cast<float*>(reinterpret_cast<char*>(regions[r]) + byte_offset)[e] = v;
    }
}

// out = scale * sum_r slot[r]  (slot r at base + r * stride floats), fixed rank order: bit-identical on every rank
__global__ void __launch_bounds__(256) sum_slots_kernel(const float4* __restrict__ base, int world, int64_t stride4, int64_t n,
                                                        float scale, float4* __restrict__ out) {
    const int64_t n4 = n >> 2;
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {          // tail
        const int64_t e = n4 * 4 + threadIdx.x;
        float s = 0.f;
        for (int r = 0; r < world; ++r) s += __ldcg(reinterpret_cast<const float*>(base) + (int64_t)r * stride4 * 4 + e);
        reinterpret_cast<float*>(out)[e] = s * scale;
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < world; ++r) {
            const float4 v = __ldcg(base + (int64_t)r * stride4 + i);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        out[i] = make_float4(s.x * scale, s.y * scale, s.z * scale, s.w * scale);
    }
}

// dst[idx[g], :] = scale[g] * src[g, :]   (idx injective: a pure scatter; dst may be peer memory)
__global__ void __launch_bounds__(128) scatter_scaled_rows_kernel(const int32_t* __restrict__ idx, const float* __restrict__ scale,
                                                                  const float* __restrict__ src, int64_t lds, int width,
                                                                  float* __restrict__ dst, int64_t ldd, int n_dst) {
    const int g = blockIdx.x;
    const int j = idx[g];
    if (j < 0 || j >= n_dst) return;
    const float s = scale[g];
    for (int f = threadIdx.x; f < width; f += blockDim.x) dst[(int64_t)j * ldd + f] = s * src[(int64_t)g * lds + f];
}

}  // namespace

extern "C" int64_t gnm_p2p_buffer_bytes(void) { return (int64_t)P2P_BYTES; }

extern "C" int gnm_p2p_alloc(void** buf, unsigned char* handle) { return gnm_p2p_alloc_bytes(buf, handle, (int64_t)P2P_BYTES); }

extern "C" int gnm_p2p_alloc_bytes(void** buf, unsigned char* handle, int64_t bytes) {
    if (!buf || !handle || bytes <= 0) return GNM_ERR_BAD_ARG;
    cudaError_t e = cudaMalloc(buf, (size_t)bytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(*buf, 0, (size_t)bytes);
    if (e != cudaSuccess) return (int)e;
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, *buf);
    if (e != cudaSuccess) return (int)e;
    static_assert(sizeof(cudaIpcMemHandle_t) == GNM_P2P_HANDLE_BYTES, "IPC handle size");
    memcpy(handle, &h, sizeof(h));
    return GNM_OK;
}

extern "C" int gnm_p2p_open(const unsigned char* handle, void** buf) {
    if (!buf || !handle) return GNM_ERR_BAD_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    cudaError_t e = cudaIpcOpenMemHandle(buf, h, cudaIpcMemLazyEnablePeerAccess);
    return e == cudaSuccess ? GNM_OK : (int)e;
}

extern "C" int gnm_p2p_close(void* buf, int owner) {
    if (!buf) return GNM_OK;
    cudaError_t e = owner ? cudaFree(buf) : cudaIpcCloseMemHandle(buf);
    return e == cudaSuccess ? GNM_OK : (int)e;
}

extern "C" int gnm_p2p_allreduce(double* data, int n, const gnm_p2p_comm* comm, gnm_stream_t stream) {
    if (n < 0) return GNM_ERR_BAD_ARG;
    if (n == 0) return GNM_OK;
    if (!data) return GNM_ERR_BAD_ARG;
    P2PArgs a;
    const int rc = p2p_args(comm, n, &a);
    if (rc == 1) return GNM_OK;
    if (rc != GNM_OK) return rc;
    gnm_count_launch(GNM_K_OTHER);
    p2p_allreduce_kernel<<<1, 256, 0, gnm_cast_stream(stream)>>>(data, n, a);
    GNM_RETURN_IF_LAUNCH_FAILED();
    return GNM_OK;
}

extern "C" int gnm_p2p_push(const float* src, int64_t n, void* const* regions, int world, int64_t byte_offset,
                            gnm_stream_t stream) {
    if (n < 0 || world < 1 || byte_offset < 0) return GNM_ERR_BAD_ARG;
    if (n == 0) return GNM_OK;
    if (!src || !regions) return GNM_ERR_BAD_ARG;
    if ((byte_offset & 15) || !gnm_aligned16(src)) return GNM_ERR_ALIGN;
    int64_t blocks = (n / 4 + 255) / 256;
    if (blocks > 148 * 2) blocks = 148 * 2;
    gnm_count_launch(GNM_K_OTHER);
    p2p_push_kernel<<<(int)blocks, 256, 0, gnm_cast_stream(stream)>>>(reinterpret_cast<const float4*>(src), n, regions, world,
                                                                      byte_offset);
    GNM_RETURN_IF_LAUNCH_FAILED();
    return GNM_OK;
}

extern "C" int gnm_sum_slots(const float* base, int world, int64_t stride, int64_t n, float scale, float* out,
                             gnm_stream_t stream) {
    if (n < 0 || world < 1 || stride < 0) return GNM_ERR_BAD_ARG;
    if (n == 0) return GNM_OK;
    if (!base || !out) return GNM_ERR_BAD_ARG;
    if ((stride & 3) || !gnm_aligned16(base) || !gnm_aligned16(out)) return GNM_ERR_ALIGN;
    int64_t blocks = (n / 4 + 255) / 256;
    if (blocks > 148 * 2) blocks = 148 * 2;
    gnm_count_launch(GNM_K_OTHER);
    sum_slots_kernel<<<(int)blocks, 256, 0, gnm_cast_stream(stream)>>>(reinterpret_cast<const float4*>(base), world, stride / 4,
                                                                       n, scale, reinterpret_cast<float4*>(out));
    GNM_RETURN_IF_LAUNCH_FAILED();
    return GNM_OK;
}

extern "C" int gnm_scatter_scaled_rows(const int32_t* idx, const float* scale, const float* src, int64_t lds, int n_src,
                                       int width, float* dst, int64_t ldd, int n_dst, gnm_stream_t stream) {
    if (n_src < 0 || width < 0 || n_dst < 0) return GNM_ERR_BAD_ARG;
    if (n_src == 0 || width == 0) return GNM_OK;
    if (!idx || !scale || !src || !dst) return GNM_ERR_BAD_ARG;
    gnm_count_launch(GNM_K_OTHER);
    scatter_scaled_rows_kernel<<<n_src, 128, 0, gnm_cast_stream(stream)>>>(idx, scale, src, lds, width, dst, ldd, n_dst);
    GNM_RETURN_IF_LAUNCH_FAILED();
    return GNM_OK;
}

int gnm_p2p_abort_flag_mlp(int* aborted);        // gnm_mlp.cu
int gnm_p2p_abort_flag_mlp_bwd(int* aborted);    // gnm_mlp_bwd.cu

/* *aborted = 1 if any peer exchange since the last call gave up waiting for a peer (its result is then invalid).
 * Synchronises the device; for tests and shutdown checks. */
extern "C" int gnm_p2p_status(int* aborted) {
    int v = 0, zero = 0;
    cudaError_t e = cudaMemcpyFromSymbol(&v, g_p2p_abort, sizeof(int));
    if (e != cudaSuccess) return (int)e;
    e = cudaMemcpyToSymbol(g_p2p_abort, &zero, sizeof(int));
    if (e != cudaSuccess) return (int)e;
    if (aborted) *aborted = v;
    int rc = gnm_p2p_abort_flag_mlp(aborted);
    if (rc != GNM_OK) return rc;
    return gnm_p2p_abort_flag_mlp_bwd(aborted);
}
