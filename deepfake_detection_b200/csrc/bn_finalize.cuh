// BatchNorm finalisation as device functions + descriptors, so that the LAST CTA of the kernel that produced a layer's
// statistics turns them into the per-channel vectors its consumers read (forward: scale / shift / mean / rstd + running
// statistics; backward: dgamma / dbeta + the coefficients of dy = cA*g + cB*y + cC) instead of a one-block launch per
// BatchNorm sitting on the critical path between two big kernels (98 such launches per EfficientNet-B0 step).
// Reference: nn.BatchNorm2d train-mode semantics (biased variance to normalise, unbiased into running_var, momentum 0.1,
// eps 1e-5 unless overridden: dfd/timm/models/efficientnet_blocks.py:22-30) and its autograd backward.
#pragma once
#include "common.cuh"

struct BnFinDesc {            // forward, one per BatchNorm layer of a plan (device memory, built by the host once)
    const double* dsum;       // [DFD_STAT_SLOTS][C]
    const double* dsq;
    const float* gamma;
    const float* beta;
    float* running_mean;
    float* running_var;
    long long* nbt;
    float* scale;
    float* shift;
    float* mean;
    float* rstd;
    int* ticket;              // zero at rest
    double inv_count;
    double unbias;            // count / (count - 1)
    float momentum;
    float eps;
    int C;
    int _pad;
};

struct BnBwdFinDesc {         // backward
    const double* s1;         // sum g
    const double* s2;         // sum g * xhat
    const float* gamma;
    const float* mean;
    const float* rstd;
    float* dgamma;
    float* dbeta;
    float* cA;
    float* cB;
    float* cC;
    int* ticket;
    double inv_count;
    int C;
    int _pad;
};

__device__ __forceinline__ double stat_total_cg(const double* base, int C, int c) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < DFD_STAT_SLOTS; i++) s += __ldcg(base + (size_t)i * C + c);
    return s;
}

// one channel of the train-mode forward finalisation
__device__ __forceinline__ void bn_finalize_channel(const BnFinDesc& d, int c) {
    // reciprocals come from the host: fp64 divisions are long dependent sequences on this part
    const double m = stat_total_cg(d.dsum, d.C, c) * d.inv_count;
    double v = stat_total_cg(d.dsq, d.C, c) * d.inv_count - m * m;
    if (v < 0) v = 0;
    const float mean = (float)m, var = (float)v;
    d.running_mean[c] = (1.f - d.momentum) * d.running_mean[c] + d.momentum * mean;
    d.running_var[c] = (1.f - d.momentum) * d.running_var[c] + d.momentum * (float)(v * d.unbias);
    float rstd = rsqrtf(var + d.eps);
    rstd = rstd * (1.5f - 0.5f * (var + d.eps) * rstd * rstd);      // one Newton step: matches 1/sqrt to fp32 round-off
    const float sc = d.gamma[c] * rstd;
    d.scale[c] = sc;
    d.shift[c] = d.beta[c] - mean * sc;
    d.mean[c] = mean;
    d.rstd[c] = rstd;
}

__device__ __forceinline__ void bn_bwd_finalize_channel(const BnBwdFinDesc& d, int c) {
    const double sum_g = stat_total_cg(d.s1, d.C, c), sum_gx = stat_total_cg(d.s2, d.C, c);
    d.dgamma[c] += (float)sum_gx;
    d.dbeta[c] += (float)sum_g;
    const float m1 = (float)(sum_g * d.inv_count), m2 = (float)(sum_gx * d.inv_count);
    const float A = d.gamma[c] * d.rstd[c];
    const float B = -A * d.rstd[c] * m2;
    d.cA[c] = A;
    d.cB[c] = B;
    d.cC[c] = -A * m1 - B * d.mean[c];
}

// Every thread of every CTA calls this once, after the CTA's last statistics atomic. True (in every thread of that CTA) for the
// last CTA of the grid to arrive; the counter is back at zero for the next launch.
__device__ __forceinline__ bool grid_last_cta(int* ticket, int tid) {
    __shared__ int s_last_cta;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const int total = (int)(gridDim.x * gridDim.y * gridDim.z);
        const int t = atomicAdd(ticket, 1);
        s_last_cta = (t == total - 1);
        if (s_last_cta) *ticket = 0;
    }
    __syncthreads();
    const bool last = s_last_cta != 0;
    if (last) __threadfence();
    return last;
}

__device__ __forceinline__ void bn_finalize_tail(const BnFinDesc* fin, int tid, int nt) {
    if (!fin) return;                                   // uniform: a kernel argument
    if (!grid_last_cta(fin->ticket, tid)) return;
    const BnFinDesc d = *fin;
    for (int c = tid; c < d.C; c += nt) bn_finalize_channel(d, c);
    if (tid == 0 && d.nbt) *d.nbt += 1;
}

__device__ __forceinline__ void bn_bwd_finalize_tail(const BnBwdFinDesc* fin, int tid, int nt) {
    if (!fin) return;
    if (!grid_last_cta(fin->ticket, tid)) return;
    const BnBwdFinDesc d = *fin;
    for (int c = tid; c < d.C; c += nt) bn_bwd_finalize_channel(d, c);
}
