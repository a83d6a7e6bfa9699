// Peer-memory exchange buffers (CUDA IPC) and the standalone small all-reduce; the fused users are gnm_bn_finalize
// (gnm_mlp.cu) and gnm_bn_bwd_coeffs (gnm_mlp_bwd.cu). Protocol: gnm_p2p.cuh.
#include "gnm_p2p.cuh"

namespace {

constexpr size_t P2P_BYTES = (size_t)2 * GNM_P2P_MAX_WORLD * GNM_P2P_MAX_DOUBLES * sizeof(double) +
                             (size_t)2 * GNM_P2P_MAX_WORLD * sizeof(unsigned int) + 256;

__global__ void __launch_bounds__(256) p2p_allreduce_kernel(double* data, int n, const P2PArgs a) {
    p2p_allreduce_block(data, n, a);
}

}  // namespace

extern "C" int64_t gnm_p2p_buffer_bytes(void) { return (int64_t)P2P_BYTES; }

extern "C" int gnm_p2p_alloc(void** buf, unsigned char* handle) {
    if (!buf || !handle) return GNM_ERR_BAD_ARG;
    cudaError_t e = cudaMalloc(buf, P2P_BYTES);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(*buf, 0, P2P_BYTES);
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
