// All-reduce of a few hundred doubles across the GPUs of one NVLink / NVSwitch domain, done by the CONSUMING kernel
// itself over peer memory (no NCCL launch, no separate collective kernel): the synchronised BatchNorm sums of the
// data-parallel step ([sum, sumsq] forward, [sum dy, sum dy*xhat] backward; 20 exchanges per step, each on the
// critical path) - see gnm_bn_finalize / gnm_bn_bwd_coeffs / gnm_p2p_allreduce.
//
// Every rank owns one exchange buffer (cudaMalloc + CUDA IPC, mapped by all peers):
//   double   slot[2][GNM_P2P_MAX_WORLD][GNM_P2P_MAX_DOUBLES]   payload of call parity p from source rank s
//   unsigned flag[2][GNM_P2P_MAX_WORLD]                        sequence number of the last payload written there
// Call k (k = 1, 2, ... counted per rank in device memory, so CUDA-graph replays stay in step): rank r stores its
// payload into slot[k&1][r] of EVERY rank (remote stores over NVLink), fences at system scope, stores k into the
// matching flags, then waits until all its own flags of that parity show k and adds the slots in rank order - every
// rank gets the bit-identical sum. The wait is bounded (~10 s: long enough for the rank-to-rank skew of a CUDA-graph
// capture in front of the first replay, short enough never to hang a box). Two parities suffice: a peer can only start call k+2 after it has seen my flag of
// call k+1, which I write after I finished reading call k.
#pragma once
#include "gnm_common.cuh"

namespace {

__device__ int g_p2p_abort = 0;      // raised when a peer did not show up within ~10 s (per translation unit)
constexpr long long P2P_TIMEOUT_CYCLES = 20000000000LL;

struct P2PArgs {
    void* const* peers;        // device array [world]: exchange buffer of every rank as mapped in THIS process
    unsigned int* counter;     // this rank's call counter (device memory)
    int rank, world;
};

// host: validate a communicator for an n-double exchange. 0 = ok, 1 = single process (nothing to do), < 0 = error
inline int p2p_args(const gnm_p2p_comm* comm, int n, P2PArgs* out) {
    out->peers = nullptr; out->counter = nullptr; out->rank = 0; out->world = 1;
    if (comm == nullptr || comm->world <= 1) return 1;
    if (comm->world > GNM_P2P_MAX_WORLD || comm->rank < 0 || comm->rank >= comm->world || !comm->peers || !comm->counter)
        return GNM_ERR_BAD_ARG;
    if (n > GNM_P2P_MAX_DOUBLES) return GNM_ERR_TOO_LARGE;
    out->peers = comm->peers;
    out->counter = comm->counter;
    out->rank = comm->rank;
    out->world = comm->world;
    return GNM_OK;
}

__device__ __forceinline__ double* p2p_slot(void* base, int par, int src) {
    return reinterpret_cast<double*>(base) + ((size_t)par * GNM_P2P_MAX_WORLD + src) * GNM_P2P_MAX_DOUBLES;
}
__device__ __forceinline__ unsigned int* p2p_flag(void* base, int par, int src) {
    return reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(base) +
                                           (size_t)2 * GNM_P2P_MAX_WORLD * GNM_P2P_MAX_DOUBLES * sizeof(double)) +
           par * GNM_P2P_MAX_WORLD + src;
}

// Whole-CTA call (blockDim.x >= world, n <= GNM_P2P_MAX_DOUBLES); data[0..n) is replaced by the sum over all ranks.
// Exactly ONE CTA per rank may execute it per call.
__device__ __forceinline__ void p2p_allreduce_block(double* data, int n, const P2PArgs a) {
    __shared__ unsigned int s_seq;
    __shared__ int s_gave_up;
    const int tid = threadIdx.x;
    if (tid == 0) {
        s_gave_up = 0;
        s_seq = *a.counter + 1;
        *a.counter = s_seq;
    }
    __syncthreads();
    const unsigned int seq = s_seq;
    const int par = (int)(seq & 1u);
    for (int i = tid; i < n * a.world; i += blockDim.x) {
        const int p = i / n, e = i - p * n;
        p2p_slot(a.peers[p], par, a.rank)[e] = data[e];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < a.world) {
        unsigned int* f = p2p_flag(a.peers[tid], par, a.rank);
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(seq) : "memory");
        const unsigned int* mine = p2p_flag(a.peers[a.rank], par, tid);
        unsigned int v;
        const long long t0 = clock64();
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
            if (v != seq) {
                __nanosleep(100);
                if (clock64() - t0 > P2P_TIMEOUT_CYCLES) {      // bounded: never hang the GPU on a missing peer
                    g_p2p_abort = 1;
                    s_gave_up = 1;
                    break;
                }
            }
        } while (v != seq);
    }
    __syncthreads();
    void* local = a.peers[a.rank];
    // a peer that never showed up must not yield a plausible-looking partial sum: poison the result, so that the
    // BatchNorm statistics, the loss and every gradient of this step turn NaN (and gnm_p2p_status reports why)
    const bool gave_up = s_gave_up != 0;
    for (int e = tid; e < n; e += blockDim.x) {
        double s = gave_up ? __longlong_as_double(0x7ff8000000000000LL) : 0.0;
        for (int p = 0; p < a.world; ++p) {
            const volatile double* q = p2p_slot(local, par, p);
            s += q[e];
        }
        data[e] = s;
    }
    __syncthreads();
}

}  // namespace
