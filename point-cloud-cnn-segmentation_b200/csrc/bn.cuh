// Train-mode BatchNorm finalize shared by the CUDA-core kernels (pointwise.cuh) and the transform stage of the tcgen05 GEMM
// (gemm.cuh).
#pragma once
#include "ptx.cuh"

namespace pcseg {

// ---------------------------------------------------------------------------------------------
// Train-mode BatchNorm finalize, folded into the kernels that consume the normalisation (no launch of its own).
//   stats = {sum y, sum y^2} over n rows (conv bias excluded) ->
//   {scale = gamma*invstd, shift = beta - mean*scale, invstd, -mean*invstd}
// Every consumer thread evaluates `bn_from_stats` for the channels it needs; block 0 of the consumer also stores the
// result (later kernels and backward read it) and updates running_mean / running_var (momentum, unbiased variance,
// conv bias re-added to the mean).
// ---------------------------------------------------------------------------------------------
struct BnFinalizeArgs {
    const double* stats;      // [2][C]
    const float* gamma;
    const float* beta;
    const float* conv_bias;
    float* rmean;             // running statistics (updated by block 0), may be null
    float* rvar;
    float4* bnp;              // [C] output copy for later kernels
    double n;
    double inv_n;             // 1 / n, computed on the host: CUDA-core fp64 (a division above all) is slow on this part and every
                              // consumer thread evaluates bn_from_stats
    float eps, momentum;
    int C;
};
__device__ __forceinline__ float4 bn_from_stats(const BnFinalizeArgs& f, int c) {
    // only the cancellation-prone part (E[y^2] - mean^2) is done in fp64: this runs in every consumer thread
    const double inv_n = f.inv_n;
    const double mean = f.stats[c] * inv_n;
    double var = fma(-mean, mean, f.stats[f.C + c] * inv_n);
    if (var < 0.0) var = 0.0;
    const float invstd = rsqrtf(static_cast<float>(var) + f.eps);
    const float meanf = static_cast<float>(mean);
    const float sc = f.gamma[c] * invstd;
    return make_float4(sc, fmaf(-meanf, sc, f.beta[c]), invstd, -meanf * invstd);
}
// (tid, nth): the threads that share the work -- a whole block (bn_publish) or a group of warps of one
__device__ __forceinline__ void bn_publish_by(const BnFinalizeArgs& f, int tid, int nth) {
    for (int c = tid; c < f.C; c += nth) {
        f.bnp[c] = bn_from_stats(f, c);
        if (f.rmean != nullptr) {
            const double mean = f.stats[c] / f.n;
            double var = f.stats[f.C + c] / f.n - mean * mean;
            if (var < 0.0) var = 0.0;
            const double unb = f.n > 1.0 ? var * f.n / (f.n - 1.0) : var;
            f.rmean[c] = static_cast<float>((1.0 - f.momentum) * f.rmean[c] + f.momentum * (mean + f.conv_bias[c]));
            f.rvar[c] = static_cast<float>((1.0 - f.momentum) * f.rvar[c] + f.momentum * unb);
        }
    }
}
__device__ __forceinline__ void bn_publish(const BnFinalizeArgs& f, bool first_block) {
    if (first_block) bn_publish_by(f, threadIdx.x, blockDim.x);
}

}  // namespace pcseg
