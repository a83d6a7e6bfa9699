// BatchNorm finalisation folded into the kernel that PRODUCES the batch statistics (include/gnm.h: gnm_bn_tail).
// gnm_bn_finalize / gnm_bn_bwd_coeffs are 64-channel kernels of a few microseconds each, 20 per training step, every one
// of them a serial link between two big kernels; here the last CTA of the producer to finish (atomic ticket) does the
// same arithmetic - including, data parallel, the peer-memory all-reduce of the sums, which needs exactly one CTA per
// rank - so the link disappears from the launch list.
#pragma once
#include "gnm_p2p.cuh"

namespace {

struct BnTailDev {
    int kind;                       // 0 = none, 1 = forward finalize, 2 = backward coefficients
    int n_feat;
    double* stats;                  // [2 * n_feat] sums the producer accumulates with atomics
    double count;
    const float* gamma; const float* beta;
    float eps, momentum;
    float* running_mean; float* running_var; int64_t* nbt;
    float* scale; float* shift; float* mean; float* rstd;
    float* coef;
    unsigned int* counter;
    P2PArgs comm;
};

// host: translate the C-ABI struct; 0 = ok (out->kind == 0 when tail == NULL), < 0 = error
inline int bn_tail_args(const gnm_bn_tail* tail, double* stats, int n_feat, BnTailDev* out) {
    out->kind = 0; out->n_feat = n_feat; out->stats = stats; out->comm.world = 1; out->comm.rank = 0;
    out->comm.peers = nullptr; out->comm.counter = nullptr;
    if (tail == nullptr) return GNM_OK;
    if (tail->kind != GNM_BN_TAIL_FINALIZE && tail->kind != GNM_BN_TAIL_BWD_COEFFS) return GNM_ERR_BAD_ARG;
    if (!stats || !tail->counter || tail->count <= 0.0 || !tail->mean || !tail->rstd) return GNM_ERR_BAD_ARG;
    if (tail->kind == GNM_BN_TAIL_FINALIZE && (!tail->scale || !tail->shift)) return GNM_ERR_BAD_ARG;
    if (tail->kind == GNM_BN_TAIL_BWD_COEFFS && !tail->coef) return GNM_ERR_BAD_ARG;
    const int prc = p2p_args(tail->comm, 2 * n_feat, &out->comm);
    if (prc < 0) return prc;
    out->kind = tail->kind; out->count = tail->count; out->gamma = tail->gamma; out->beta = tail->beta;
    out->eps = tail->eps; out->momentum = tail->momentum; out->running_mean = tail->running_mean;
    out->running_var = tail->running_var; out->nbt = tail->num_batches_tracked; out->scale = tail->scale;
    out->shift = tail->shift; out->mean = tail->mean; out->rstd = tail->rstd; out->coef = tail->coef;
    out->counter = tail->counter;
    return GNM_OK;
}

__device__ __forceinline__ void bn_finalize_channel(int c, int n_feat, const double* stats, double count, const float* gamma,
                                                    const float* beta, float eps, float momentum, float* running_mean,
                                                    float* running_var, float* scale, float* shift, float* mean_o,
                                                    float* rstd_o) {
    const double mean = __ldcg(stats + c) / count;          // L2: the sums were produced by other CTAs' atomics
    double var = __ldcg(stats + n_feat + c) / count - mean * mean;
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

__device__ __forceinline__ void bn_bwd_coeffs_channel(int c, int n_feat, const double* stats, double count, const float* gamma,
                                                      const float* mean, const float* rstd, float* coef) {
    const float g = (gamma ? gamma[c] : 1.f) * rstd[c];
    float a = g, b = 0.f, k = 0.f;
    if (stats != nullptr) {
        const float m1 = (float)(__ldcg(stats + c) / count), m2 = (float)(__ldcg(stats + n_feat + c) / count);
        b = -g * rstd[c] * m2;
        k = g * (rstd[c] * m2 * mean[c] - m1);
    }
    coef[c] = a;
    coef[n_feat + c] = b;
    coef[2 * n_feat + c] = k;
}

// Called by EVERY thread of EVERY CTA as the last thing the kernel does, after the CTA's own atomics on t.stats were
// issued. The last CTA to arrive finalises. blockDim.x >= 32 (>= the data-parallel world size).
__device__ __forceinline__ void bn_tail_run(const BnTailDev& t) {
    if (t.kind == 0) return;
    __shared__ int s_bn_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int total = gridDim.x * gridDim.y * gridDim.z;
        s_bn_last = atomicAdd(t.counter, 1u) == total - 1;
    }
    __syncthreads();
    if (!s_bn_last) return;
    __threadfence();
    if (t.comm.world > 1) p2p_allreduce_block(t.stats, 2 * t.n_feat, t.comm);
    const double* st = t.stats;
    for (int c = threadIdx.x; c < t.n_feat; c += blockDim.x) {
        if (t.kind == 1)
            bn_finalize_channel(c, t.n_feat, st, t.count, t.gamma, t.beta, t.eps, t.momentum,
                                t.running_mean, t.running_var, t.scale, t.shift, t.mean, t.rstd);
        else
            bn_bwd_coeffs_channel(c, t.n_feat, st, t.count, t.gamma, t.mean, t.rstd, t.coef);
    }
    if (threadIdx.x == 0) {
        if (t.kind == 1 && t.nbt != nullptr) *t.nbt += 1;
        *t.counter = 0u;
    }
}

}  // namespace
