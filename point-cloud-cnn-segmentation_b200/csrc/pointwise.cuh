// CUDA-core kernels around the tcgen05 GEMMs: ingest (4->64 layer), BatchNorm finalize /
// apply / backward, global max-pool, per-cloud bias, logits + cross-entropy head, Adam.
// All activation tensors are point-major [P][C] bf16 with C contiguous.
#pragma once
#include "ptx.cuh"
#include "bn.cuh"

namespace pcseg {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------
// Weight preparation
// ---------------------------------------------------------------------------------------------
// dst[r][c] (bf16, pitch ld_dst) = alpha[r] * src[r][c] (fp32, pitch ld_src); alpha may be null.
// dst_lo (optional, same pitch): the bf16 remainder v - float(bf16(v)) (split-bf16 inference operands [hi | lo]).
__global__ void k_convert_rows(const float* __restrict__ src, int ld_src, __nv_bfloat16* __restrict__ dst, int ld_dst,
                               int rows, int cols, const float* __restrict__ alpha, __nv_bfloat16* __restrict__ dst_lo) {
    pdl_launch_dependents();
    pdl_wait();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * cols) return;
    const int r = idx / cols, c = idx % cols;
    float v = src[static_cast<size_t>(r) * ld_src + c];
    if (alpha) v *= alpha[r];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    dst[static_cast<size_t>(r) * ld_dst + c] = hi;
    if (dst_lo) dst_lo[static_cast<size_t>(r) * ld_dst + c] = __float2bfloat16_rn(v - __bfloat162float(hi));
}
// dst[c][r] (bf16, pitch ld_dst) = src[r][c]: transposed copy used as the dgrad B operand.
__global__ void k_convert_transpose(const float* __restrict__ src, int ld_src, __nv_bfloat16* __restrict__ dst, int ld_dst,
                                    int rows, int cols) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < rows && c < cols) ? src[static_cast<size_t>(r) * ld_src + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < rows && c < cols) dst[static_cast<size_t>(c) * ld_dst + r] = __float2bfloat16_rn(tile[threadIdx.x][i]);
    }
}

// All fp32 -> bf16 weight conversions of one training step in a single launch (blockIdx.y = job).
struct ConvertJob {
    const float* src;
    __nv_bfloat16* dst;
    int ld_src, ld_dst, rows, cols, transpose;
};
constexpr int MAX_CONVERT_JOBS = 20;
// Buffers that a step needs zeroed (or filled with a 32-bit pattern) ride along as extra blockIdx.y slices: one graph node
// instead of one memset node per buffer on the critical path.
struct FillJob {
    void* dst;
    unsigned long long bytes;   // multiple of 4
    unsigned int value;
};
constexpr int MAX_FILL_JOBS = 10;
struct ConvertJobs {
    ConvertJob job[MAX_CONVERT_JOBS];
    int count;
    FillJob fill[MAX_FILL_JOBS];
    int fill_count;
};
__device__ __forceinline__ void fill_slice(const FillJob f) {
    const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x, nth = static_cast<size_t>(gridDim.x) * blockDim.x;
    unsigned int* w = static_cast<unsigned int*>(f.dst);
    const size_t nw = f.bytes >> 2;
    size_t head = ((16u - (reinterpret_cast<uintptr_t>(w) & 15u)) & 15u) >> 2;     // words up to 16-byte alignment
    if (head > nw) head = nw;
    const size_t n16 = (nw - head) >> 2;
    uint4* v = reinterpret_cast<uint4*>(w + head);
    const uint4 val = make_uint4(f.value, f.value, f.value, f.value);
    for (size_t i = tid; i < n16; i += nth) v[i] = val;
    for (size_t i = tid; i < head; i += nth) w[i] = f.value;
    for (size_t i = head + (n16 << 2) + tid; i < nw; i += nth) w[i] = f.value;
}
__global__ void __launch_bounds__(256) k_convert_multi(const ConvertJobs jobs) {
    pdl_launch_dependents();
    pdl_wait();
    if (static_cast<int>(blockIdx.y) >= jobs.count) {
        fill_slice(jobs.fill[blockIdx.y - jobs.count]);
        return;
    }
    // blockIdx.y = job; 32x32 tiles through shared memory so that both the fp32 reads and the (possibly transposed)
    // bf16 writes are coalesced
    __shared__ float tile[32][33];
    const ConvertJob j = jobs.job[blockIdx.y];
    const int tiles_c = (j.cols + 31) >> 5, tiles_r = (j.rows + 31) >> 5;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    for (int t = blockIdx.x; t < tiles_c * tiles_r; t += gridDim.x) {
        const int r0 = (t / tiles_c) << 5, c0 = (t % tiles_c) << 5;
#pragma unroll
        for (int i = ty; i < 32; i += 8) {
            const int r = r0 + i, c = c0 + tx;
            tile[i][tx] = (r < j.rows && c < j.cols) ? j.src[static_cast<size_t>(r) * j.ld_src + c] : 0.f;
        }
        __syncthreads();
        if (j.transpose) {
#pragma unroll
            for (int i = ty; i < 32; i += 8) {
                const int c = c0 + i, r = r0 + tx;
                if (r < j.rows && c < j.cols) j.dst[static_cast<size_t>(c) * j.ld_dst + r] = __float2bfloat16_rn(tile[tx][i]);
            }
        } else {
#pragma unroll
            for (int i = ty; i < 32; i += 8) {
                const int r = r0 + i, c = c0 + tx;
                if (r < j.rows && c < j.cols) j.dst[static_cast<size_t>(r) * j.ld_dst + c] = __float2bfloat16_rn(tile[i][tx]);
            }
        }
        __syncthreads();
    }
}
// the fills alone (backward: gradient arena, BN-backward sums, per-cloud sums, side buffers)
struct FillJobs {
    FillJob fill[MAX_FILL_JOBS];
    int count;
};
__global__ void __launch_bounds__(256) k_fill_multi(const FillJobs jobs) {
    pdl_launch_dependents();
    pdl_wait();
    fill_slice(jobs.fill[blockIdx.y]);
}

// Eval-mode BatchNorm folding: alpha[c] = gamma/sqrt(var+eps), delta[c] = (bias - mean)*alpha + beta.
__global__ void k_fold_bn(const float* __restrict__ conv_bias, const float* __restrict__ gamma, const float* __restrict__ beta,
                          const float* __restrict__ rmean, const float* __restrict__ rvar, float eps, int C,
                          float* __restrict__ alpha, float* __restrict__ delta) {
    pdl_launch_dependents();
    pdl_wait();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float a2 = gamma[c] / sqrtf(rvar[c] + eps);
    alpha[c] = a2;
    delta[c] = (conv_bias[c] - rmean[c]) * a2 + beta[c];
}

// ---------------------------------------------------------------------------------------------
// Ingest layer conv1 (4 -> 64) on CUDA cores.  8 threads per point, 8 channels each.
//   EVAL : a1 = relu(Wf x + bf) with BN folded                     (writes bf16)
//   TRAIN: y1 = W x (bias dropped: train-mode BN cancels it), column sum / sum-of-squares in fp64
// ---------------------------------------------------------------------------------------------
//   EVAL, x3 != 0 (split-bf16 inference): the fp32 result is stored as a bf16 pair, row = [hi (64) | lo (64)]
template <bool TRAIN>
__global__ void __launch_bounds__(256) k_ingest(const float4* __restrict__ x, int P, const float* __restrict__ W /*[64][4]*/,
                                                const float* __restrict__ alpha, const float* __restrict__ delta,
                                                __nv_bfloat16* __restrict__ out, double* __restrict__ stats, int x3) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float red[2][32][64];   // per point-slot partial sums (TRAIN only)
    const int cg = threadIdx.x & 7;    // channel group: channels cg*8 .. cg*8+7
    const int slot = threadIdx.x >> 3; // 0..31
    float w[8][4], bsh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = cg * 8 + j;
        const float a = TRAIN ? 1.f : alpha[c];
#pragma unroll
        for (int k = 0; k < 4; ++k) w[j][k] = a * W[c * 4 + k];
        bsh[j] = TRAIN ? 0.f : delta[c];
    }
    float s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
    for (int pnt = blockIdx.x * 32 + slot; pnt < P; pnt += gridDim.x * 32) {
        const float4 xv = __ldg(x + pnt);
        float o[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float v = fmaf(w[j][0], xv.x, fmaf(w[j][1], xv.y, fmaf(w[j][2], xv.z, fmaf(w[j][3], xv.w, bsh[j]))));
            if (!TRAIN) v = fmaxf(v, 0.f);
            const float full = v;
            v = round_bf16(v);
            o[j] = v;
            lo[j] = full - v;
            if (TRAIN) { s1[j] += v; s2[j] = fmaf(v, v, s2[j]); }
        }
        uint4 pk = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
        if (!TRAIN && x3) {
            *reinterpret_cast<uint4*>(out + static_cast<size_t>(pnt) * 128 + cg * 8) = pk;
            *reinterpret_cast<uint4*>(out + static_cast<size_t>(pnt) * 128 + 64 + cg * 8) =
                make_uint4(pack_bf16x2(lo[0], lo[1]), pack_bf16x2(lo[2], lo[3]), pack_bf16x2(lo[4], lo[5]), pack_bf16x2(lo[6], lo[7]));
            continue;
        }
        *reinterpret_cast<uint4*>(out + static_cast<size_t>(pnt) * 64 + cg * 8) = pk;
    }
    if (TRAIN) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { red[0][slot][cg * 8 + j] = s1[j]; red[1][slot][cg * 8 + j] = s2[j]; }
        __syncthreads();
        if (threadIdx.x < 128) {
            const int q = threadIdx.x >> 6, c = threadIdx.x & 63;
            double s = 0.0;
            for (int i = 0; i < 32; ++i) s += static_cast<double>(red[q][i][c]);
            atomicAdd(stats + q * 64 + c, s);
        }
    }
}

// a = relu(scale*y + shift) (* dropout keep / (1-p)).  Each thread owns 8 fixed channels (scale/shift live in
// registers) and walks down the rows of its block's strip; 4 independent 16-byte loads in flight per thread.
// colsum (optional, [C] fp64, zeroed by the caller): column sums of the STORED (bf16-rounded) activation, the `s` of the
// Gram-predicted statistics of the next layer (k_predict_bn).
__global__ void __launch_bounds__(256) k_bn_relu(const __nv_bfloat16* __restrict__ y, int ld_y, __nv_bfloat16* __restrict__ a,
                                                 int ld_a, long P, int C, const BnFinalizeArgs fin,
                                                 unsigned long long seed_arg, const unsigned long long* __restrict__ seed_ptr,
                                                 unsigned int thr16, float keep_scale, double* __restrict__ colsum,
                                                 int rows_per_cloud) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float red[256 * 8];
    float csum[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) csum[e] = 0.f;
    bn_publish(fin, blockIdx.x == 0);
    const unsigned long long seed = seed_arg + (seed_ptr != nullptr ? *seed_ptr : 0ull);
    const int tpr = C >> 3;                       // threads per row (C <= 2048)
    const int rpp = 256 / tpr;                    // rows per pass
    const int c0 = (threadIdx.x % tpr) << 3;
    const int rslot = threadIdx.x / tpr;
    float sc[8], sh[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const float4 bp = bn_from_stats(fin, c0 + e);
        sc[e] = bp.x;
        sh[e] = bp.y;
    }
    // rows_per_cloud > 0: the grid is (clouds x blocks per cloud), strips never straddle clouds and colsum is [clouds][C]
    long r_begin, r_end;
    if (rows_per_cloud > 0) {
        const int clouds = static_cast<int>(P / rows_per_cloud);
        const int bpc = gridDim.x / clouds;
        const int cloud = blockIdx.x / bpc, within = blockIdx.x - cloud * bpc;
        const long rpb = (rows_per_cloud + bpc - 1) / bpc;
        r_begin = static_cast<long>(cloud) * rows_per_cloud + within * rpb;
        r_end = min(static_cast<long>(cloud + 1) * rows_per_cloud, r_begin + rpb);
        if (cloud >= clouds) r_end = r_begin;
        if (colsum != nullptr) colsum += static_cast<size_t>(cloud < clouds ? cloud : 0) * C;
    } else {
        const long rows_per_block = (P + gridDim.x - 1) / gridDim.x;
        r_begin = blockIdx.x * rows_per_block;
        r_end = min(P, r_begin + rows_per_block);
    }
    // software pipelined: the four loads of the NEXT pass are in flight while this pass does its arithmetic (the Philox
    // rounds of the dropout variant are a long dependent chain that otherwise delays the next loads)
    uint4 ynext[4];
    auto fetch = [&](long r) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long rr = r + static_cast<long>(u) * rpp;
            if (rr < r_end) ynext[u] = *reinterpret_cast<const uint4*>(y + rr * ld_y + c0);
        }
    };
    if (r_begin + rslot < r_end) fetch(r_begin + rslot);
    for (long r = r_begin + rslot; r < r_end; r += 4L * rpp) {
        uint4 yw[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) yw[u] = ynext[u];
        if (r + 4L * rpp < r_end) fetch(r + 4L * rpp);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long rr = r + static_cast<long>(u) * rpp;
            if (rr >= r_end) break;
            const uint32_t ws[4] = {yw[u].x, yw[u].y, yw[u].z, yw[u].w};
            uint32_t keep = 0xFFu;
            if (thr16 != 0u) keep = dropout_keep8(seed, (static_cast<unsigned long long>(rr) * C + c0) >> 3, thr16);
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float yv = (e & 1) ? bf16_hi(ws[e >> 1]) : bf16_lo(ws[e >> 1]);
                const float t = fmaf(sc[e], yv, sh[e]);
                o[e] = (t > 0.f && ((keep >> e) & 1u)) ? t * keep_scale : 0.f;
            }
            const uint4 pk = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
            *reinterpret_cast<uint4*>(a + rr * ld_a + c0) = pk;
            if (colsum != nullptr) {
                const uint32_t ps[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) csum[e] += (e & 1) ? bf16_hi(ps[e >> 1]) : bf16_lo(ps[e >> 1]);
            }
        }
    }
    if (colsum != nullptr) {          // red[row slot][C]: combine the row slots, then one fp64 atomic per column per block
#pragma unroll
        for (int e = 0; e < 8; ++e) red[rslot * C + c0 + e] = csum[e];
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += 256) {
            float sum = 0.f;
            for (int sl = 0; sl < rpp; ++sl) sum += red[sl * C + c];
            atomicAdd(colsum + c, static_cast<double>(sum));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Gram-predicted train-mode BatchNorm (oracle/folded_bn_ref.py identity (1)).
// The batch statistics of y = a W^T over n rows follow from s = sum_p a and G = a^T a (an MN-major tcgen05 GEMM with
// A = B = a):   mean[c] = W[c,:] m,   var[c] = W[c,:] (G/n - m m^T) W[c,:]^T,   m = s/n.
// Knowing them BEFORE the GEMM runs lets its epilogue apply BN + ReLU directly (EPI_BN_RELU): y is never stored.
// W is the bf16 copy the GEMM multiplies with, so the statistics describe the fp32 accumulators exactly.
// One warp per output channel, fp64 throughout (K x K is small).  Also updates the running statistics (momentum,
// unbiased variance, conv bias re-added to the mean) and writes {sum y, sum y^2} for inspection.
// ---------------------------------------------------------------------------------------------
// Sum of the per-split partial tiles of a Gram GEMM (EPI_WGRAD, wg_mode 3) in a FIXED order, fp64 accumulation: the
// batch statistics predicted from it are then reproducible from run to run.  16 lanes per element.  blockIdx.y = group
// (cloud): out[g][i] = sum over the `splits` consecutive partial tiles of group g.  With colsum != nullptr the result is the
// CENTRED Gram matrix  Gc[r][c] = G[r][c] - s[r] s[c] / n_rows  (s = colsum + g * K): the only fp64 arithmetic of the
// predicted-statistics path happens here, once per matrix element (CUDA-core fp64 is slow on this part).
__global__ void __launch_bounds__(256) k_gram_reduce(const float* __restrict__ part, int splits, int n_elem, float* __restrict__ out,
                                                     const double* __restrict__ colsum, double n_rows, int K) {
    pdl_launch_dependents();
    pdl_wait();
    part += static_cast<size_t>(blockIdx.y) * splits * n_elem;
    out += static_cast<size_t>(blockIdx.y) * n_elem;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = t >> 4, q = t & 15;
    double acc = 0.0;
    if (i < n_elem) {
#pragma unroll 4
        for (int s = q; s < splits; s += 16) acc += static_cast<double>(part[static_cast<size_t>(s) * n_elem + i]);
    }
#pragma unroll
    for (int o = 1; o <= 8; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (q == 0 && i < n_elem) {
        if (colsum != nullptr) {
            const double* sg = colsum + static_cast<size_t>(blockIdx.y) * K;
            acc -= sg[i / K] * sg[i % K] / n_rows;
        }
        out[i] = static_cast<float>(acc);
    }
}

template <int KQ>      // K = 32 * KQ input channels; 8 warps = 8 output channels per block
__global__ void __launch_bounds__(256) k_predict_bn(const float* __restrict__ Gc /* centred Gram matrix */, const double* __restrict__ colsum,
                                                    const __nv_bfloat16* __restrict__ W, const BnFinalizeArgs fin,
                                                    double* __restrict__ stats_out) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int K = 32 * KQ;
    __shared__ float g_s[32][K];          // 32 rows of Gc at a time (all warps of the block share them)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 8 + warp;
    const bool live = c < fin.C;
    const __nv_bfloat16* wr = W + static_cast<size_t>(live ? c : 0) * K;
    float wf[KQ], t[KQ];
#pragma unroll
    for (int q = 0; q < KQ; ++q) {
        t[q] = 0.f;
        wf[q] = __bfloat162float(wr[lane + 32 * q]);
    }
#pragma unroll
    for (int jc = 0; jc < KQ; ++jc) {          // rows j = 32 jc .. 32 jc + 31 (unrolled: wf[jc] must stay in registers)
        __syncthreads();
        for (int i = threadIdx.x; i < 32 * K; i += blockDim.x) g_s[i / K][i % K] = Gc[static_cast<size_t>(32 * jc) * K + i];
        __syncthreads();
#pragma unroll 8
        for (int jj = 0; jj < 32; ++jj) {
            const float wj = __shfl_sync(0xffffffffu, wf[jc], jj);
#pragma unroll
            for (int q = 0; q < KQ; ++q) t[q] = fmaf(g_s[jj][lane + 32 * q], wj, t[q]);      // Gc is symmetric: row j read conflict-free
        }
    }
    double nvar = 0.0, sum_y = 0.0;          // n * var = w^T Gc w,  sum y = w . s
#pragma unroll
    for (int q = 0; q < KQ; ++q) {
        nvar += static_cast<double>(wf[q]) * static_cast<double>(t[q]);
        sum_y += static_cast<double>(wf[q]) * colsum[lane + 32 * q];
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        nvar += __shfl_xor_sync(0xffffffffu, nvar, o);
        sum_y += __shfl_xor_sync(0xffffffffu, sum_y, o);
    }
    if (lane != 0 || !live) return;
    const double mean = sum_y / fin.n;
    double var = nvar / fin.n;
    if (var < 0.0) var = 0.0;
    const float invstd = rsqrtf(static_cast<float>(var) + fin.eps);
    const float meanf = static_cast<float>(mean);
    const float sc = fin.gamma[c] * invstd;
    fin.bnp[c] = make_float4(sc, fmaf(-meanf, sc, fin.beta[c]), invstd, -meanf * invstd);
    if (fin.rmean != nullptr) {
        const double unb = fin.n > 1.0 ? var * fin.n / (fin.n - 1.0) : var;
        fin.rmean[c] = static_cast<float>((1.0 - fin.momentum) * fin.rmean[c] + fin.momentum * (mean + fin.conv_bias[c]));
        fin.rvar[c] = static_cast<float>((1.0 - fin.momentum) * fin.rvar[c] + fin.momentum * unb);
    }
    if (stats_out != nullptr) {
        stats_out[c] = sum_y;
        stats_out[fin.C + c] = (var + mean * mean) * fin.n;
    }
}

// ---------------------------------------------------------------------------------------------
// Train-mode global max-pool over points of relu(bn(y6)).  Because bn is monotone per channel
// the arg-extremum of the pre-BN value decides: max(y) if gamma >= 0 else min(y).  The scan itself is fused into the
// global_feat GEMM epilogue (EPI_STATS_POOL): packed 64-bit key = (orderable(sign*y) << 32) | ~index, reduced with
// atomicMax (ties -> lowest index); k_maxpool_finish decodes the keys.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t float_orderable(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float float_from_orderable(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}
// decode keys -> g (post BN+ReLU), ystar (pre-BN extremum), argidx (row within cloud)
__global__ void k_maxpool_finish(const unsigned long long* __restrict__ keys, int total, int C, const BnFinalizeArgs fin,
                                 float* __restrict__ g, float* __restrict__ ystar, int* __restrict__ argidx) {
    pdl_launch_dependents();
    pdl_wait();
    bn_publish(fin, blockIdx.x == 0);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = i % C;
    const unsigned long long k = keys[i];
    const float4 bp = bn_from_stats(fin, c);
    const float sg = bp.x >= 0.f ? 1.f : -1.f;
    const float yv = sg * float_from_orderable(static_cast<uint32_t>(k >> 32));
    argidx[i] = static_cast<int>(0xFFFFFFFFu - static_cast<uint32_t>(k & 0xFFFFFFFFull));
    ystar[i] = yv;
    g[i] = fmaxf(fmaf(bp.x, yv, bp.y), 0.f);
}

// ---------------------------------------------------------------------------------------------
// Per-cloud bias of seg_conv1: cb[b][n] = alpha[n] * sum_k Wg[n][k] * g[b][k] + delta[n]
// (Wg = seg_conv1.weight[:, 64:], fp32, pitch ldw).  One warp per output.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_cloud_bias(const float* __restrict__ Wg, int ldw, const float* __restrict__ g, int clouds,
                                                    int Nout, int K, const float* __restrict__ alpha,
                                                    const float* __restrict__ delta, float* __restrict__ cb) {
    pdl_launch_dependents();
    pdl_wait();
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= clouds * Nout) return;
    const int b = gw / Nout, n = gw % Nout;
    const float* wr = Wg + static_cast<size_t>(n) * ldw;
    const float* gr = g + static_cast<size_t>(b) * K;
    float s = 0.f;
    for (int k = lane; k < K; k += 32) s = fmaf(wr[k], gr[k], s);
    s = warp_sum(s);
    if (lane == 0) cb[gw] = (alpha ? alpha[n] : 1.f) * s + (delta ? delta[n] : 0.f);
}

// ---------------------------------------------------------------------------------------------
// Train-mode head: logits = W4 * relu(bn(y_s3)) + b4.  One warp per point (lane = 4 channels).
// Optionally fused weighted cross-entropy statistics: sum w*nll, sum w, argmax-correct count.
// ---------------------------------------------------------------------------------------------
struct CeAccum {
    double loss_num;      // sum_i w[y_i] * (-log p_i[y_i])
    double w_sum;         // sum_i w[y_i]
    unsigned long long correct;   // argmax == label over valid points
    unsigned long long valid;     // labels != -1
};

// Train-mode head forward, ONE THREAD PER POINT: the point's 128 pre-BN values are streamed with 16-byte loads (a
// warp touches 32 different rows per instruction, but consecutive instructions hit the same lines in L1, so DRAM traffic
// stays at the algorithmic 256 B/point), BN + ReLU + the 128 x NC mat-vec run on private registers with no shuffles,
// and the weighted-CE terms are accumulated per thread and reduced once per block.
// NC <= 8: exactly NC classes (everything unrolled on compile-time bounds).  NC = 16 / 32: the wide variants for 9..32
// classes, Crt = the actual count (padded class slots carry zero weights and are skipped by the loss / argmax / stores).
// HAS_BN = false: the input already is the post-ReLU activation (inference, BatchNorm folded into the GEMM before).
template <int NC, bool HAS_BN>
__global__ void __launch_bounds__(256) k_head_fwd(const __nv_bfloat16* __restrict__ ys3, long P, const BnFinalizeArgs fin,
                                                  const float* __restrict__ W4, const float* __restrict__ b4, int Crt,
                                                  float* __restrict__ logits, const long long* __restrict__ labels,
                                                  const float* __restrict__ class_w, CeAccum* __restrict__ ce) {
    pdl_launch_dependents();
    pdl_wait();
    const int C = (NC <= 8) ? NC : Crt;
    constexpr int NCP = (NC + 3) & ~3;                  // weights per channel padded to a multiple of 4 floats
    __shared__ __align__(16) float w_s[128 * NCP];      // [channel][class]
    __shared__ float2 bn_s[128];                        // {scale, shift}
    __shared__ float bias_s[NC], cw_s[NC];
    __shared__ double red_d[8][2];
    __shared__ unsigned long long red_u[8][2];
    if (HAS_BN) {
        bn_publish(fin, blockIdx.x == 0);
        if (threadIdx.x < 128) {
            const float4 bp = bn_from_stats(fin, threadIdx.x);
            bn_s[threadIdx.x] = make_float2(bp.x, bp.y);
        }
    }
    for (int i = threadIdx.x; i < 128 * NCP; i += blockDim.x) {
        const int c = i / NCP, k = i % NCP;
        w_s[i] = (k < C) ? W4[k * 128 + c] : 0.f;
    }
    if (threadIdx.x < NC) {
        bias_s[threadIdx.x] = (static_cast<int>(threadIdx.x) < C) ? __ldg(b4 + threadIdx.x) : 0.f;
        cw_s[threadIdx.x] = (class_w != nullptr && static_cast<int>(threadIdx.x) < C) ? __ldg(class_w + threadIdx.x) : 1.f;
    }
    __syncthreads();
    double loss_num = 0.0, w_sum = 0.0;
    unsigned long long correct = 0, nvalid = 0;
    const long stride = static_cast<long>(gridDim.x) * blockDim.x;
    for (long pnt = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; pnt < P; pnt += stride) {
        const uint4* row = reinterpret_cast<const uint4*>(ys3 + pnt * 128);
        float z[NC];
#pragma unroll
        for (int k = 0; k < NC; ++k) z[k] = bias_s[k];
#pragma unroll
        for (int j0 = 0; j0 < 16; j0 += 4) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = __ldg(row + j0 + u);      // 4 independent 16-byte loads in flight
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t ws[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int c = (j0 + u) * 8 + e;
                    const float yv = (e & 1) ? bf16_hi(ws[e >> 1]) : bf16_lo(ws[e >> 1]);
                    float a = yv;
                    if (HAS_BN) {
                        const float2 b2 = bn_s[c];
                        a = fmaxf(fmaf(b2.x, yv, b2.y), 0.f);
                    }
                    const float4* wc = reinterpret_cast<const float4*>(w_s + c * NCP);
#pragma unroll
                    for (int q = 0; q < NCP / 4; ++q) {
                        const float4 w4 = wc[q];                      // broadcast: every lane reads the same address
                        if (4 * q < NC) z[4 * q] = fmaf(a, w4.x, z[4 * q]);
                        if (4 * q + 1 < NC) z[4 * q + 1] = fmaf(a, w4.y, z[4 * q + 1]);
                        if (4 * q + 2 < NC) z[4 * q + 2] = fmaf(a, w4.z, z[4 * q + 2]);
                        if (4 * q + 3 < NC) z[4 * q + 3] = fmaf(a, w4.w, z[4 * q + 3]);
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < NC; ++k)
            if (k < C) logits[pnt * C + k] = z[k];
        if (labels != nullptr) {
            const long long lab = labels[pnt];
            if (lab >= 0 && lab < C) {
                float zmax = z[0];
                int am = 0;
#pragma unroll
                for (int k = 1; k < NC; ++k)
                    if (k < C && z[k] > zmax) { zmax = z[k]; am = k; }        // first maximum, like torch.argmax
                float se = 0.f, zl = 0.f;
                const float wl = cw_s[lab];
#pragma unroll
                for (int k = 0; k < NC; ++k) {
                    if (k < C) se += __expf(z[k] - zmax);
                    if (k == lab) zl = z[k];
                }
                loss_num += static_cast<double>(wl) * static_cast<double>(zmax + logf(se) - zl);
                w_sum += wl;
                correct += (am == lab);
                nvalid += 1;
            }
        }
    }
    if (labels != nullptr) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            loss_num += __shfl_xor_sync(0xffffffffu, loss_num, o);
            w_sum += __shfl_xor_sync(0xffffffffu, w_sum, o);
            correct += __shfl_xor_sync(0xffffffffu, correct, o);
            nvalid += __shfl_xor_sync(0xffffffffu, nvalid, o);
        }
        const int warp = threadIdx.x >> 5;
        if ((threadIdx.x & 31) == 0) { red_d[warp][0] = loss_num; red_d[warp][1] = w_sum; red_u[warp][0] = correct; red_u[warp][1] = nvalid; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double a0 = 0.0, a1 = 0.0;
            unsigned long long u0 = 0, u1 = 0;
            for (int i = 0; i < 8; ++i) { a0 += red_d[i][0]; a1 += red_d[i][1]; u0 += red_u[i][0]; u1 += red_u[i][1]; }
            atomicAdd(&ce->loss_num, a0);
            atomicAdd(&ce->w_sum, a1);
            atomicAdd(&ce->correct, u0);
            atomicAdd(&ce->valid, u1);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Head backward: dlogits -> {dW4, db4, dz_s3 (masked by relu), BN stats sum dz / sum dz*yhat}.
// dlogits either given (autograd path) or recomputed from logits/labels/class_w/inv_wsum (fused CE).
// One warp per point; lane = 4 channels.
// ---------------------------------------------------------------------------------------------
// NC = exact number of classes (compile-time, so the per-lane dW4 accumulators are NC x 16 registers)
template <int NC>
__global__ void __launch_bounds__(256) k_head_bwd(const __nv_bfloat16* __restrict__ ys3, long P, const float4* __restrict__ bnp,
                                                  const float* __restrict__ W4, const float* __restrict__ dlogits,
                                                  const float* __restrict__ logits, const long long* __restrict__ labels,
                                                  const float* __restrict__ class_w, const double* __restrict__ wsum_total,
                                                  __nv_bfloat16* __restrict__ dz_out, float* __restrict__ dW4,
                                                  float* __restrict__ db4, double* __restrict__ stats) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int MAXC = NC;
    constexpr int C = NC;
    // 8 lanes per point, 16 channels per lane (see k_head_fwd).  Per-lane accumulators: dW4[k][16 ch], sum dz[16], sum dz*yhat[16].
    __shared__ __align__(16) float w_s[MAXC * 128];
    __shared__ float red[8][MAXC * 128 + 2 * 128 + MAXC];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int sub = lane & 7;
    const int grp = lane >> 3;
    // bank-conflict-free layout [class][q][sub][4]: logical channel = sub*16 + q*4 + j (lanes of a point read 8 distinct
    // 16-byte chunks that cover all 32 banks)
    for (int i = threadIdx.x; i < C * 128; i += blockDim.x) {
        const int k = i >> 7, c = i & 127;
        w_s[k * 128 + ((c >> 2) & 3) * 32 + (c >> 4) * 4 + (c & 3)] = W4[i];
    }
    float sc[16], sh[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        const float4 b4v = __ldg(bnp + sub * 16 + e);
        sc[e] = b4v.x;
        sh[e] = b4v.y;
    }
    float dw[MAXC][16];
    float s1[16], s2[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        s1[e] = s2[e] = 0.f;
#pragma unroll
        for (int k = 0; k < MAXC; ++k) dw[k][e] = 0.f;
    }
    float dbk = 0.f;                                   // lane with sub == k accumulates db4[k]
    const bool fused = (dlogits == nullptr);
    const float inv_wsum = (wsum_total != nullptr) ? static_cast<float>(1.0 / *wsum_total) : 0.f;
    const float cw_k = (class_w != nullptr && sub < C) ? __ldg(class_w + sub) : 1.f;
    const float* zsrc = fused ? logits : dlogits;
    __syncthreads();
    const long warp_g = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = (static_cast<long>(gridDim.x) * blockDim.x) >> 5;
    // The kernel runs one block per SM (NC x 16 + 96 accumulator registers per lane) and every iteration is a chain
    // load -> softmax shuffles -> FMAs: the loads of the NEXT iteration are issued before the arithmetic of the current one,
    // otherwise each of the ~28 iterations per warp pays a full memory latency.
    uint4 y0n = make_uint4(0, 0, 0, 0), y1n = y0n;
    float zinn = 0.f;
    long long lab0n = -1;
    auto fetch = [&](long p0) {
        const long pnt = p0 + grp;
        y0n = make_uint4(0, 0, 0, 0);
        y1n = y0n;
        zinn = 0.f;
        lab0n = -1;
        if (pnt < P) {
            const uint4* src = reinterpret_cast<const uint4*>(ys3 + pnt * 128 + sub * 16);
            y0n = src[0];
            y1n = src[1];
            if (sub < C) zinn = zsrc[pnt * C + sub];
            if (fused && sub == 0) lab0n = labels[pnt];
        }
    };
    if (warp_g * 4 < P) fetch(warp_g * 4);
    for (long p0 = warp_g * 4; p0 < P; p0 += nwarps * 4) {
        const long pnt = p0 + grp;
        const bool ok = pnt < P;
        const uint4 y0 = y0n, y1 = y1n;
        const float zin = zinn;
        const long long lab0 = lab0n;
        if (p0 + nwarps * 4 < P) fetch(p0 + nwarps * 4);
        float dl_mine = zin;                               // lane with sub == k holds dlogit k of its point
        if (fused) {
            const long long lab = __shfl_sync(0xffffffffu, lab0, grp * 8);
            const float zm = (sub < C) ? zin : -INFINITY;
            float zmax = zm;
#pragma unroll
            for (int o = 4; o >= 1; o >>= 1) zmax = fmaxf(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
            const float ex = (sub < C) ? __expf(zm - zmax) : 0.f;
            float se = ex;
#pragma unroll
            for (int o = 4; o >= 1; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
            const float wl = __shfl_sync(0xffffffffu, cw_k, grp * 8 + static_cast<int>((lab >= 0 && lab < C) ? lab : 0));
            dl_mine = (lab >= 0 && lab < C) ? wl * inv_wsum * (ex / se - (sub == lab ? 1.f : 0.f)) : 0.f;
        }
        if (sub >= C || !ok) dl_mine = 0.f;
        dbk += dl_mine;
        float dl[MAXC];
#pragma unroll
        for (int k = 0; k < MAXC; ++k) dl[k] = (k < C) ? __shfl_sync(0xffffffffu, dl_mine, grp * 8 + k) : 0.f;
        const uint32_t ws[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
        float dz[16], da[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) da[e] = 0.f;
#pragma unroll
        for (int k = 0; k < MAXC; ++k) {
            const float4* wk = reinterpret_cast<const float4*>(w_s + k * 128 + sub * 4);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 w4 = wk[q * 8];
                da[4 * q] = fmaf(dl[k], w4.x, da[4 * q]);
                da[4 * q + 1] = fmaf(dl[k], w4.y, da[4 * q + 1]);
                da[4 * q + 2] = fmaf(dl[k], w4.z, da[4 * q + 2]);
                da[4 * q + 3] = fmaf(dl[k], w4.w, da[4 * q + 3]);
            }
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const float yv = (e & 1) ? bf16_hi(ws[e >> 1]) : bf16_lo(ws[e >> 1]);
            const float t = fmaf(sc[e], yv, sh[e]);
            const float a = fmaxf(t, 0.f);
#pragma unroll
            for (int k = 0; k < MAXC; ++k) dw[k][e] = fmaf(dl[k], a, dw[k][e]);
            dz[e] = (t > 0.f) ? round_bf16(da[e]) : 0.f;
            s1[e] += dz[e];
            s2[e] = fmaf(dz[e], yv, s2[e]);                 // sum dz*y; turned into sum dz*yhat below
        }
        if (ok) {
            uint4* dst = reinterpret_cast<uint4*>(dz_out + pnt * 128 + sub * 16);
            dst[0] = make_uint4(pack_bf16x2(dz[0], dz[1]), pack_bf16x2(dz[2], dz[3]), pack_bf16x2(dz[4], dz[5]), pack_bf16x2(dz[6], dz[7]));
            dst[1] = make_uint4(pack_bf16x2(dz[8], dz[9]), pack_bf16x2(dz[10], dz[11]), pack_bf16x2(dz[12], dz[13]), pack_bf16x2(dz[14], dz[15]));
        }
    }
    // fold the 4 point slots of a warp (lanes with equal `sub`), then the 8 warps through shared memory
#pragma unroll
    for (int e = 0; e < 16; ++e) {
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {
            s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], o);
            s2[e] += __shfl_xor_sync(0xffffffffu, s2[e], o);
#pragma unroll
            for (int k = 0; k < MAXC; ++k)
                if (k < C) dw[k][e] += __shfl_xor_sync(0xffffffffu, dw[k][e], o);
        }
    }
    dbk += __shfl_xor_sync(0xffffffffu, dbk, 8);
    dbk += __shfl_xor_sync(0xffffffffu, dbk, 16);
#pragma unroll
    for (int e = 0; e < 16; ++e) {                       // sum dz*yhat = invstd * sum dz*y + (-mean*invstd) * sum dz
        const float4 b4v = __ldg(bnp + sub * 16 + e);
        s2[e] = fmaf(b4v.z, s2[e], b4v.w * s1[e]);
    }
    float* r = red[warp];
    if (grp == 0) {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
#pragma unroll
            for (int k = 0; k < MAXC; ++k) r[k * 128 + sub * 16 + e] = dw[k][e];
            r[MAXC * 128 + sub * 16 + e] = s1[e];
            r[MAXC * 128 + 128 + sub * 16 + e] = s2[e];
        }
        if (sub < MAXC) r[MAXC * 128 + 256 + sub] = dbk;
    }
    __syncthreads();
    const int total = MAXC * 128 + 256 + MAXC;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int wv = 0; wv < 8; ++wv) s += red[wv][i];
        if (i < MAXC * 128) {
            if (i / 128 < C) atomicAdd(dW4 + i, s);
        } else if (i < MAXC * 128 + 256) {
            atomicAdd(stats + (i - MAXC * 128), static_cast<double>(s));
        } else if (i - MAXC * 128 - 256 < C) {
            atomicAdd(db4 + (i - MAXC * 128 - 256), s);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Head backward for 9..32 classes (k_head_bwd keeps NC x 16 gradient accumulators per lane, which stops at 8 classes).
// Two kernels: (A) one thread per point: dlogits (fused CE gradient, or the caller's) -> dl buffer, dz of seg_conv3;
//              (B) column reductions: dW4, db4 and the BN-backward sums, thread = (channel, half of the classes).
// NC = 16 / 32 class slots, C = actual count.
// ---------------------------------------------------------------------------------------------
template <int NC>
__global__ void __launch_bounds__(256) k_head_bwd_wide_points(const __nv_bfloat16* __restrict__ ys3, long P, const float4* __restrict__ bnp,
                                                              const float* __restrict__ W4, int C, const float* __restrict__ dlogits,
                                                              const float* __restrict__ logits, const long long* __restrict__ labels,
                                                              const float* __restrict__ class_w, const double* __restrict__ wsum_total,
                                                              float* __restrict__ dlbuf, __nv_bfloat16* __restrict__ dz_out) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ __align__(16) float w_s[128 * NC];      // [channel][class]
    __shared__ float2 bn_s[128];
    __shared__ float cw_s[NC];
    for (int i = threadIdx.x; i < 128 * NC; i += blockDim.x) {
        const int c = i / NC, k = i % NC;
        w_s[i] = (k < C) ? W4[k * 128 + c] : 0.f;
    }
    if (threadIdx.x < 128) {
        const float4 bp = __ldg(bnp + threadIdx.x);
        bn_s[threadIdx.x] = make_float2(bp.x, bp.y);
    }
    if (threadIdx.x < NC) cw_s[threadIdx.x] = (class_w != nullptr && static_cast<int>(threadIdx.x) < C) ? __ldg(class_w + threadIdx.x) : 1.f;
    __syncthreads();
    const bool fused = (dlogits == nullptr);
    const float inv_wsum = (wsum_total != nullptr) ? static_cast<float>(1.0 / *wsum_total) : 0.f;
    const long stride = static_cast<long>(gridDim.x) * blockDim.x;
    for (long pnt = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; pnt < P; pnt += stride) {
        float dl[NC];
        if (fused) {
            const long long lab = labels[pnt];
            float zmax = -INFINITY;
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                dl[k] = (k < C) ? logits[pnt * C + k] : -INFINITY;
                zmax = fmaxf(zmax, dl[k]);
            }
            float se = 0.f;
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                dl[k] = (k < C) ? __expf(dl[k] - zmax) : 0.f;
                se += dl[k];
            }
            const bool okl = lab >= 0 && lab < C;
            const float f = okl ? cw_s[okl ? lab : 0] * inv_wsum : 0.f;
#pragma unroll
            for (int k = 0; k < NC; ++k) dl[k] = f * (dl[k] / se - (k == lab ? 1.f : 0.f));
#pragma unroll
            for (int k = 0; k < NC; ++k)
                if (k < C) dlbuf[pnt * C + k] = dl[k];
        } else {
#pragma unroll
            for (int k = 0; k < NC; ++k) dl[k] = (k < C) ? dlogits[pnt * C + k] : 0.f;
        }
        const uint4* row = reinterpret_cast<const uint4*>(ys3 + pnt * 128);
        uint4* dst = reinterpret_cast<uint4*>(dz_out + pnt * 128);
#pragma unroll 1
        for (int j = 0; j < 16; ++j) {
            const uint4 v = __ldg(row + j);
            const uint32_t ws[4] = {v.x, v.y, v.z, v.w};
            float dz[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int c = j * 8 + e;
                const float yv = (e & 1) ? bf16_hi(ws[e >> 1]) : bf16_lo(ws[e >> 1]);
                const float2 b2 = bn_s[c];
                const float t = fmaf(b2.x, yv, b2.y);
                const float4* wc = reinterpret_cast<const float4*>(w_s + c * NC);
                float da = 0.f;
#pragma unroll
                for (int q = 0; q < NC / 4; ++q) {
                    const float4 w4 = wc[q];
                    da = fmaf(dl[4 * q], w4.x, da);
                    da = fmaf(dl[4 * q + 1], w4.y, da);
                    da = fmaf(dl[4 * q + 2], w4.z, da);
                    da = fmaf(dl[4 * q + 3], w4.w, da);
                }
                dz[e] = (t > 0.f) ? round_bf16(da) : 0.f;
            }
            dst[j] = make_uint4(pack_bf16x2(dz[0], dz[1]), pack_bf16x2(dz[2], dz[3]), pack_bf16x2(dz[4], dz[5]), pack_bf16x2(dz[6], dz[7]));
        }
    }
}

template <int NC>
__global__ void __launch_bounds__(256) k_head_bwd_wide_reduce(const __nv_bfloat16* __restrict__ ys3, const __nv_bfloat16* __restrict__ dz,
                                                              long P, const float4* __restrict__ bnp, int C,
                                                              const float* __restrict__ dl /*[P][C]*/, float* __restrict__ dW4,
                                                              float* __restrict__ db4, double* __restrict__ stats) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int HALF = NC / 2;
    constexpr int CHUNK = 64;                          // points staged per iteration
    __shared__ float dl_s[CHUNK][NC];
    const int c = threadIdx.x & 127, half = threadIdx.x >> 7;
    const float4 bp = __ldg(bnp + c);
    float dw[HALF];
#pragma unroll
    for (int k = 0; k < HALF; ++k) dw[k] = 0.f;
    float s1 = 0.f, s2 = 0.f, dbk = 0.f;
    const long per_block = (P + gridDim.x - 1) / gridDim.x;
    const long p_begin = blockIdx.x * per_block, p_end = (p_begin + per_block < P) ? p_begin + per_block : P;
    for (long p0 = p_begin; p0 < p_end; p0 += CHUNK) {
        const int n = static_cast<int>((p_end - p0 < CHUNK) ? p_end - p0 : CHUNK);
        __syncthreads();
        for (int i = threadIdx.x; i < CHUNK * NC; i += 256) {
            const int j = i / NC, k = i % NC;
            dl_s[j][k] = (j < n && k < C) ? dl[(p0 + j) * C + k] : 0.f;
        }
        __syncthreads();
        if (static_cast<int>(threadIdx.x) < NC)
            for (int j = 0; j < n; ++j) dbk += dl_s[j][threadIdx.x];
        for (int j = 0; j < n; ++j) {
            const float yv = __bfloat162float(ys3[(p0 + j) * 128 + c]);
            const float a = fmaxf(fmaf(bp.x, yv, bp.y), 0.f);
#pragma unroll
            for (int k = 0; k < HALF; ++k) dw[k] = fmaf(dl_s[j][half * HALF + k], a, dw[k]);
            if (half == 0) {
                const float dzv = __bfloat162float(dz[(p0 + j) * 128 + c]);
                s1 += dzv;
                s2 = fmaf(dzv, yv, s2);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < HALF; ++k)
        if (half * HALF + k < C) atomicAdd(dW4 + (half * HALF + k) * 128 + c, dw[k]);
    if (half == 0) {      // sum dz*yhat = invstd * sum dz*y + (-mean*invstd) * sum dz
        atomicAdd(stats + c, static_cast<double>(s1));
        atomicAdd(stats + 128 + c, static_cast<double>(bp.z) * s2 + static_cast<double>(bp.w) * s1);
    }
    if (static_cast<int>(threadIdx.x) < C) atomicAdd(db4 + threadIdx.x, dbk);
}

// ---------------------------------------------------------------------------------------------
// BatchNorm backward.
//   coefficients (evaluated inside k_bn_bwd_apply, no launch of their own):  dy = A*dz + Bc*y + Cc   with
//     A = scale, Bc = -scale*c2*invstd, Cc = -scale*(c1 + c2*(-mean*invstd)),  c1 = sum dz / n, c2 = sum dz*yhat / n
//   also writes dgamma = sum dz*yhat, dbeta = sum dz.
// ---------------------------------------------------------------------------------------------
struct BnBwdArgs {
    const double* stats;      // [2][C] {sum dz, sum dz*yhat}
    const float4* bnp;        // [C] forward normalisation
    float4* coef;             // [C] output copy {A, Bc, Cc, 0} (inspection / tests)
    float* dgamma;            // parameter gradients written by block (0, 0)
    float* dbeta;
    double n;
    double inv_n;             // 1 / n (host)
    int C;
};
__device__ __forceinline__ float4 bn_bwd_coef(const BnBwdArgs& f, int c) {
    const double inv_n = f.inv_n;
    const float c1 = static_cast<float>(f.stats[c] * inv_n), c2 = static_cast<float>(f.stats[f.C + c] * inv_n);
    const float4 bp = f.bnp[c];
    return make_float4(bp.x, -bp.x * c2 * bp.z, -bp.x * fmaf(c2, bp.w, c1), 0.f);
}
__device__ __forceinline__ void bn_bwd_publish(const BnBwdArgs& f, bool first_block) {
    if (!first_block) return;
    for (int c = threadIdx.x; c < f.C; c += blockDim.x) {
        f.coef[c] = bn_bwd_coef(f, c);
        f.dgamma[c] = static_cast<float>(f.stats[f.C + c]);
        f.dbeta[c] = static_cast<float>(f.stats[c]);
    }
}

// dy = A*dz + Bc*y + Cc, bf16 out (pitch ld_dy); column sums of dy -> dbias (fp32 atomics) and,
// optionally, per-cloud column sums -> dcb[cloud][C].
// SPARSE variant (max-pool backward): dz[p][c] = (row_in_cloud == argidx[cloud][c]) ? dzv[cloud][c] : 0.
// grid: (strips per cloud, clouds); every thread owns 8 fixed channels (coefficients in registers) and walks
// down the rows of its strip, two rows in flight.
// RAG variant (packed ragged batches, see k_pack_rows): 1-D grid over the strips planned on the host (rag_strips[blk] =
// {cloud, first row, end row, first row of the cloud} in packed rows, strips sized by the batch's real row count so that short clouds do not idle
// blocks); the affine term is weighted by the row multiplicity (gradients of the representative pad row are carried
// pre-multiplied, filler rows carry none).
template <bool SPARSE, bool RAG>
__global__ void __launch_bounds__(256) k_bn_bwd_apply(const __nv_bfloat16* __restrict__ dz, int ld_dz,
                                                      const __nv_bfloat16* __restrict__ y, int ld_y,
                                                      __nv_bfloat16* __restrict__ dy, int ld_dy, int N /*rows per cloud*/, int C,
                                                      int rows_per_strip, const BnBwdArgs bw,
                                                      float* __restrict__ dbias, float* __restrict__ dcb,
                                                      const int* __restrict__ argidx, const float* __restrict__ dzv,
                                                      const int* __restrict__ rag_strips, const float* __restrict__ rowmult) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float red[256 * 8];
    bn_bwd_publish(bw, blockIdx.x == 0 && blockIdx.y == 0);
    const int tpr = C >> 3;
    const int rpp = 256 / tpr;
    const int c0 = (threadIdx.x % tpr) << 3;
    const int rslot = threadIdx.x / tpr;
    const int cloud = RAG ? rag_strips[4 * blockIdx.x] : blockIdx.y;
    const size_t base = RAG ? 0 : static_cast<size_t>(cloud) * N;                 // RAG: strip bounds are packed rows already
    const int r0 = RAG ? rag_strips[4 * blockIdx.x + 1] : blockIdx.x * rows_per_strip;
    const int r1 = RAG ? rag_strips[4 * blockIdx.x + 2] : min(r0 + rows_per_strip, N);
    const int arg_base = RAG ? rag_strips[4 * blockIdx.x + 3] : 0;                // argidx is relative to the cloud
    float cA[8], cB[8], cC[8], acc[8];
    int arg[8];
    float dv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const float4 cf = bn_bwd_coef(bw, c0 + e);
        cA[e] = cf.x; cB[e] = cf.y; cC[e] = cf.z;
        acc[e] = 0.f;
        if (SPARSE) {
            arg[e] = argidx[static_cast<size_t>(cloud) * C + c0 + e] + arg_base;
            dv[e] = dzv[static_cast<size_t>(cloud) * C + c0 + e];
        }
    }
    for (int r = r0 + rslot; r < r1; r += 4 * rpp) {
        uint4 yw[4], zw[4];
        float mw[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int rr = r + u * rpp;
            if (rr < r1) {
                if (RAG) mw[u] = rowmult[base + rr];
                yw[u] = *reinterpret_cast<const uint4*>(y + (base + rr) * ld_y + c0);
                if (!SPARSE) zw[u] = *reinterpret_cast<const uint4*>(dz + (base + rr) * ld_dz + c0);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int rr = r + u * rpp;
            if (rr >= r1) break;
            const uint32_t ys[4] = {yw[u].x, yw[u].y, yw[u].z, yw[u].w};
            const uint32_t zs[4] = {zw[u].x, zw[u].y, zw[u].z, zw[u].w};
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float yv = (e & 1) ? bf16_hi(ys[e >> 1]) : bf16_lo(ys[e >> 1]);
                float dze;
                if (SPARSE) dze = (arg[e] == rr) ? dv[e] : 0.f;
                else dze = (e & 1) ? bf16_hi(zs[e >> 1]) : bf16_lo(zs[e >> 1]);
                const float aff = fmaf(cB[e], yv, cC[e]);
                const float v = round_bf16(fmaf(cA[e], dze, RAG ? mw[u] * aff : aff));
                o[e] = v;
                acc[e] += v;
            }
            *reinterpret_cast<uint4*>(dy + (base + rr) * ld_dy + c0) =
                make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
        }
    }
    // red[slot][C] : combine the row slots, then one atomic per column per block
#pragma unroll
    for (int e = 0; e < 8; ++e) red[rslot * C + c0 + e] = acc[e];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        float sum = 0.f;
        for (int sl = 0; sl < rpp; ++sl) sum += red[sl * C + c];
        if (dbias) atomicAdd(dbias + c, sum);
        if (dcb) atomicAdd(dcb + static_cast<size_t>(cloud) * C + c, sum);
    }
}

// ---------------------------------------------------------------------------------------------
// Folded BatchNorm backward (oracle/folded_bn_ref.py identities (2) and (3)) for a layer whose pre-BN
// output y = a_prev W^T is never stored.  Inputs: Q = dz^T a_prev (the raw weight-gradient GEMM), sum_dz, the Gram matrix
// G and column sums s of a_prev, the forward normalisation bnp.  With yc = y - mean:
//     dy = A dz + Bc yc + D,      A = scale,  Bc = -A invstd dgamma / n,  D = -A sum_dz / n,
//     dgamma = sum dz*yhat = invstd * rowdot(Q, W) + (-mean invstd) * sum_dz
//     dW = diag(A) Q + diag(Bc) W Gc + D s^T,            Gc = G - s s^T / n   (centred: no cancellation against D s^T)
//     dy W = [dz | a_prev] [diag(A) W ; S]  + const,     S = W^T diag(Bc) W,  const = (D - Bc mean)^T W
// k_fold_coef: one warp per output channel.  k_fold_bwd: block roles (dW | scaled transposed weights | S | const).
// ---------------------------------------------------------------------------------------------
struct FoldArgs {
    const float* Q;               // [Co][Ci]
    const __nv_bfloat16* W;       // [Co][Ci] bf16 forward weights
    const __nv_bfloat16* Wt;      // [Ci][Co] bf16 transposed weights
    const float* Gc;              // [Ci][Ci] centred Gram matrix of a_prev (k_gram_reduce)
    const double* s;              // [Ci]
    const double* sum_dz;         // [Co]
    const float4* bnp;            // [Co]
    float4* coef;                 // [Co] {A, Bc, D - Bc*mean, D}
    float* dgamma;                // parameter gradients
    float* dbeta;
    float* dbias;                 // conv bias ahead of a train-mode BN: zero
    float* dW;                    // [Co][ld_dw]
    int ld_dw;
    __nv_bfloat16* Bw;            // [Ci][ld_bw]: columns [0, Co) = A_c W[c][j], columns [Co, Co + Ci) = S[i][j]
    int ld_bw;
    float* cst;                   // [Ci]
    double n;
    int Co, Ci;
};

// per-channel coefficients: one warp per channel
__global__ void __launch_bounds__(256) k_fold_coef(const FoldArgs f) {
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= f.Co) return;
    const float* q = f.Q + static_cast<size_t>(c) * f.Ci;
    const __nv_bfloat16* w = f.W + static_cast<size_t>(c) * f.Ci;
    double dot = 0.0;
    for (int k = lane; k < f.Ci; k += 32) dot = fma(static_cast<double>(q[k]), static_cast<double>(__bfloat162float(w[k])), dot);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (lane != 0) return;
    const float4 bp = f.bnp[c];
    const double s1 = f.sum_dz[c];
    const double dgamma = static_cast<double>(bp.z) * dot + static_cast<double>(bp.w) * s1;
    const double A = bp.x;
    const double Bc = -A * static_cast<double>(bp.z) * dgamma / f.n;
    const double D = -A * s1 / f.n;
    const double mean = -static_cast<double>(bp.w) / static_cast<double>(bp.z);
    f.coef[c] = make_float4(static_cast<float>(A), static_cast<float>(Bc), static_cast<float>(D - Bc * mean), static_cast<float>(D));
    f.dgamma[c] = static_cast<float>(dgamma);
    f.dbeta[c] = static_cast<float>(s1);
    f.dbias[c] = 0.f;
}
__host__ __device__ inline int fold_coef_blocks(int Co, int Ci) { return (Co + 7) / 8; }

// out[j0 + lane] = sum_c scale(c) * W[c][j0 + lane] for one block of 256 threads: the 8 warps split the Co rows of W, lane = one
// of 32 consecutive columns (coalesced row reads, scale(c) broadcast); the partial sums meet in shared memory.  The result is
// returned to the lanes of warp 0 (other warps return 0).
template <typename ScaleFn>
__device__ __forceinline__ float block_weighted_colsum32(const __nv_bfloat16* __restrict__ W, int Ci, int Co, int j0, ScaleFn scale) {
    __shared__ float part_s[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per = (Co + 7) / 8;
    const int c0 = warp * per, c1 = min(c0 + per, Co);
    float acc = 0.f;
#pragma unroll 4
    for (int c = c0; c < c1; ++c) acc = fmaf(scale(c), __bfloat162float(W[static_cast<size_t>(c) * Ci + j0 + lane]), acc);
    part_s[warp][lane] = acc;
    __syncthreads();
    float r = 0.f;
    if (warp == 0) {
#pragma unroll
        for (int w = 0; w < 8; ++w) r += part_s[w][lane];
    }
    return r;
}

// Block roles of k_fold_bwd (1-D grid, see fold_bwd_blocks): dW (thread per element) | scaled transposed weights (thread per
// element) | S: one block per (i, 32 columns j) | constant row: one block per 32 columns.  No atomics.  Ci % 32 == 0.
__host__ __device__ inline int fold_bwd_blocks(int Co, int Ci) { return 2 * ((Co * Ci + 255) / 256) + Ci * (Ci / 32) + Ci / 32; }
__global__ void __launch_bounds__(256) k_fold_bwd(const FoldArgs f) {
    pdl_launch_dependents();
    pdl_wait();
    const int Co = f.Co, Ci = f.Ci;
    const int nA = (Co * Ci + 255) / 256;
    int blk = blockIdx.x;
    if (blk < nA) {                                   // dW[c][k] = A Q + Bc (W Gc) + D s
        const int idx = blk * 256 + threadIdx.x;
        if (idx >= Co * Ci) return;
        const int c = idx / Ci, k = idx - c * Ci;
        const __nv_bfloat16* w = f.W + static_cast<size_t>(c) * Ci;
        float acc = 0.f;
#pragma unroll 8
        for (int j = 0; j < Ci; ++j) acc = fmaf(__bfloat162float(w[j]), f.Gc[static_cast<size_t>(j) * Ci + k], acc);
        const float4 cf = f.coef[c];
        f.dW[static_cast<size_t>(c) * f.ld_dw + k] = fmaf(cf.x, f.Q[idx], fmaf(cf.y, acc, cf.w * static_cast<float>(f.s[k])));
        return;
    }
    blk -= nA;
    if (blk < nA) {                                   // Bw[j][c] = A_c W[c][j]
        const int idx = blk * 256 + threadIdx.x;
        if (idx >= Co * Ci) return;
        const int j = idx / Co, c = idx - j * Co;
        f.Bw[static_cast<size_t>(j) * f.ld_bw + c] = __float2bfloat16_rn(f.coef[c].x * __bfloat162float(f.Wt[idx]));
        return;
    }
    blk -= nA;
    const int lane = threadIdx.x & 31;
    const int jb = Ci / 32;
    if (blk < Ci * jb) {                              // S[i][j] = sum_c W[c][i] Bc_c W[c][j]  ->  Bw[j][Co + i]
        const int i = blk / jb, j0 = (blk - i * jb) * 32;
        const float r = block_weighted_colsum32(f.W, Ci, Co, j0, [&](int c) { return __bfloat162float(f.W[static_cast<size_t>(c) * Ci + i]) * f.coef[c].y; });
        if (threadIdx.x < 32) f.Bw[static_cast<size_t>(j0 + lane) * f.ld_bw + Co + i] = __float2bfloat16_rn(r);
    } else {                                          // const[j] = sum_c (D_c - Bc_c mean_c) W[c][j]
        const int j0 = (blk - Ci * jb) * 32;
        const float r = block_weighted_colsum32(f.W, Ci, Co, j0, [&](int c) { return f.coef[c].z; });
        if (threadIdx.x < 32) f.cst[j0 + lane] = r;
    }
}

// ---------------------------------------------------------------------------------------------
// seg_conv1 (Ci = 64 point features + the per-cloud term cb[b][c] of the pooled feature, pcs.py:117-123; Co = 512) with
// predicted statistics and folded backward.  y[p][c] = W[c,:] a[p,:] + cb[b(p)][c]; per cloud b: s_b = sum a, G_b = a^T a.
//   k_predict_bn_cloud : lin_b = W s_b, quad_b = W Gc_b W^T (Gc_b = G_b - s_b s_b^T / N, centred by k_gram_reduce);
//                        mean = sum_b (lin_b + N cb_b) / n,  n var = sum_b [ quad_b + N (lin_b / N + cb_b - mean)^2 ]   (centred:
//                        no cancellation against the large per-cloud term).
//   k_fold6_coef       : sum dz*y = rowdot(Q, W) + sum_b cb_b S1_b  (S1_b = per-cloud sum of dz from the GEMM epilogue),
//                        coefficients {A, Bc, D - Bc mean, D}, dgamma / dbeta / dbias, the per-cloud gradient of the pooled
//                        branch dcb[b][c] = sum_{p in b} dy = A S1_b + Bc (lin_b + N cb_b - N mean) + N D, and gsum = sum_b G_b,
//                        ssum = sum_b s_b.
//   k_fold6_bwd        : dW[c][k] = A Q + Bc ((W gsum)[c][k] + sum_b (lin_b[c] / N + cb_b[c] - mean_c) s_b[k]) + D ssum[k]
//                        (gsum = sum_b Gc_b, the centred per-cloud Gram matrices);
//                        data-gradient weights wcat[j][Co0 + c] = A_c W[c][j], wcat[j][Co0 + Co + i] = S[i][j] = sum_c W[c][i]
//                        Bc_c W[c][j]; per-cloud constant rows cst[b][j] = sum_c (Bc_c (cb_b[c] - mean_c) + D_c) W[c][j].
// ---------------------------------------------------------------------------------------------
struct Fold6Args {
    const float* Q;               // [Co][Ci]
    const __nv_bfloat16* W;       // [Co][Ci]
    const float* G;               // [clouds][Ci][Ci] centred per-cloud Gram matrices
    const double* s;              // [clouds][Ci]
    const double* part;           // [clouds][Co][2] {lin = W s_b, quad} of k_predict_bn_cloud
    const float* cb;              // [clouds][Co]
    const float* S1;              // [clouds][Co] per-cloud sums of dz
    const float4* bnp;            // [Co]
    float4* coef;                 // [Co]
    float* dgamma;
    float* dbeta;
    float* dbias;
    float* dcb;                   // [clouds][Co]
    float* gsum;                  // [Ci][Ci]
    double* ssum;                 // [Ci]
    float* dW;                    // [Co][ld_dw]
    int ld_dw;
    __nv_bfloat16* wcat;          // [Ci][ld_wcat]
    int ld_wcat, col0;            // first column of the scaled weights inside wcat
    float* cst;                   // [clouds][Ci]
    double n;                     // all rows
    int N;                        // rows per cloud
    int clouds, Co, Ci;           // Ci == 64
};

// grid (C / 8, clouds), 8 warps = 8 channels of one cloud per block: {lin, quad} of every (cloud, channel) into `part`
// ([clouds][C][2] fp64); the block that finishes a channel group last (ticket[blockIdx.x], zeroed by the caller) combines the
// clouds and publishes the normalisation.
__global__ void __launch_bounds__(256) k_predict_bn_cloud(const float* __restrict__ G, const double* __restrict__ s,
                                                          const __nv_bfloat16* __restrict__ W, const float* __restrict__ cb, int clouds,
                                                          int N, const BnFinalizeArgs fin, double* __restrict__ stats_out,
                                                          double* __restrict__ part, int* __restrict__ ticket) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int K = 64;
    __shared__ float g_s[K][K];
    __shared__ int last_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int c = blockIdx.x * 8 + warp;
    for (int i = threadIdx.x; i < K * K; i += blockDim.x) g_s[i / K][i % K] = G[static_cast<size_t>(b) * K * K + i];
    __syncthreads();
    if (c < fin.C) {
        const float wf0 = __bfloat162float(W[static_cast<size_t>(c) * K + lane]), wf1 = __bfloat162float(W[static_cast<size_t>(c) * K + 32 + lane]);
        float t0 = 0.f, t1 = 0.f;
#pragma unroll 8
        for (int j = 0; j < K; ++j) {
            const float wj = __shfl_sync(0xffffffffu, j < 32 ? wf0 : wf1, j & 31);
            t0 = fmaf(g_s[j][lane], wj, t0);
            t1 = fmaf(g_s[j][32 + lane], wj, t1);
        }
        const double* sb = s + static_cast<size_t>(b) * K;
        // quad = w^T Gc_b w = N * (within-cloud variance of W a);  lin = w . s_b
        double quad = static_cast<double>(wf0) * t0 + static_cast<double>(wf1) * t1, lin = wf0 * sb[lane] + wf1 * sb[32 + lane];
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            quad += __shfl_xor_sync(0xffffffffu, quad, o);
            lin += __shfl_xor_sync(0xffffffffu, lin, o);
        }
        if (lane == 0) {
            part[(static_cast<size_t>(b) * fin.C + c) * 2] = lin;
            part[(static_cast<size_t>(b) * fin.C + c) * 2 + 1] = quad;
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last_s = (atomicAdd(ticket + blockIdx.x, 1) == clouds - 1);
    __syncthreads();
    if (!last_s) return;
    __threadfence();
    if (threadIdx.x == 0) ticket[blockIdx.x] = 0;                    // ready for the next step
    if (threadIdx.x >= 8) return;
    const int cc = blockIdx.x * 8 + threadIdx.x;                     // one thread per channel of the group
    if (cc >= fin.C) return;
    double sum_y = 0.0, within = 0.0;
    for (int bb = 0; bb < clouds; ++bb) {
        const double lin = __ldcg(part + (static_cast<size_t>(bb) * fin.C + cc) * 2), quad = __ldcg(part + (static_cast<size_t>(bb) * fin.C + cc) * 2 + 1);
        sum_y += lin + N * static_cast<double>(cb[static_cast<size_t>(bb) * fin.C + cc]);
        within += quad;
    }
    const double mean = sum_y / fin.n;
    double between = 0.0;
    for (int bb = 0; bb < clouds; ++bb) {
        const double mu = __ldcg(part + (static_cast<size_t>(bb) * fin.C + cc) * 2) / N + static_cast<double>(cb[static_cast<size_t>(bb) * fin.C + cc]);
        between += N * (mu - mean) * (mu - mean);
    }
    double var = (within + between) / fin.n;
    if (var < 0.0) var = 0.0;
    const float invstd = rsqrtf(static_cast<float>(var) + fin.eps);
    const float meanf = static_cast<float>(mean);
    const float sc = fin.gamma[cc] * invstd;
    fin.bnp[cc] = make_float4(sc, fmaf(-meanf, sc, fin.beta[cc]), invstd, -meanf * invstd);
    if (fin.rmean != nullptr) {
        const double unb = fin.n > 1.0 ? var * fin.n / (fin.n - 1.0) : var;
        fin.rmean[cc] = static_cast<float>((1.0 - fin.momentum) * fin.rmean[cc] + fin.momentum * (mean + fin.conv_bias[cc]));
        fin.rvar[cc] = static_cast<float>((1.0 - fin.momentum) * fin.rvar[cc] + fin.momentum * unb);
    }
    if (stats_out != nullptr) {
        stats_out[cc] = sum_y;
        stats_out[fin.C + cc] = (var + mean * mean) * fin.n;
    }
}

// blocks [0, Co/8): one warp per channel; the remaining blocks sum G_b / s_b over the clouds
__global__ void __launch_bounds__(256) k_fold6_coef(const Fold6Args f) {
    pdl_launch_dependents();
    pdl_wait();
    const int K = f.Ci;
    const int nco = (f.Co + 7) / 8;
    if (static_cast<int>(blockIdx.x) >= nco) {
        const int idx = (blockIdx.x - nco) * 256 + threadIdx.x;
        if (idx < K * K) {
            float g = 0.f;
            for (int b = 0; b < f.clouds; ++b) g += f.G[static_cast<size_t>(b) * K * K + idx];
            f.gsum[idx] = g;
        }
        if (idx < K) {
            double sv = 0.0;
            for (int b = 0; b < f.clouds; ++b) sv += f.s[static_cast<size_t>(b) * K + idx];
            f.ssum[idx] = sv;
        }
        return;
    }
    const int lane = threadIdx.x & 31;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= f.Co) return;
    const float* q = f.Q + static_cast<size_t>(c) * K;
    const __nv_bfloat16* w = f.W + static_cast<size_t>(c) * K;
    const double w0 = __bfloat162float(w[lane]), w1 = __bfloat162float(w[32 + lane]);
    double dot = static_cast<double>(q[lane]) * w0 + static_cast<double>(q[32 + lane]) * w1;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    const float4 bp = f.bnp[c];
    const double mean = -static_cast<double>(bp.w) / static_cast<double>(bp.z);
    // per cloud (lane b handles cloud b, b + 32, ...): S1_b, lin_b
    double s1 = 0.0, dzy = 0.0;
    for (int b0 = 0; b0 < f.clouds; b0 += 32) {
        const int b = b0 + lane;
        double S1b = 0.0, cbv = 0.0;
        if (b < f.clouds) {
            S1b = f.S1[static_cast<size_t>(b) * f.Co + c];
            cbv = f.cb[static_cast<size_t>(b) * f.Co + c];
        }
        s1 += S1b;
        dzy += cbv * S1b;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        dzy += __shfl_xor_sync(0xffffffffu, dzy, o);
    }
    const double dgamma = static_cast<double>(bp.z) * (dot + dzy - mean * s1);
    const double A = bp.x;
    const double Bc = -A * static_cast<double>(bp.z) * dgamma / f.n;
    const double D = -A * s1 / f.n;
    // per-cloud gradient of the pooled branch: dcb[b][c] = A S1_b + Bc (lin_b + N cb_b - N mean) + N D
    for (int b = lane; b < f.clouds; b += 32) {
        const double lin = f.part[(static_cast<size_t>(b) * f.Co + c) * 2];                  // W[c,:] s_b (k_predict_bn_cloud)
        const double cbv = f.cb[static_cast<size_t>(b) * f.Co + c];
        f.dcb[static_cast<size_t>(b) * f.Co + c] =
            static_cast<float>(A * f.S1[static_cast<size_t>(b) * f.Co + c] + Bc * (lin + f.N * (cbv - mean)) + f.N * D);
    }
    if (lane != 0) return;
    f.coef[c] = make_float4(static_cast<float>(A), static_cast<float>(Bc), static_cast<float>(D - Bc * mean), static_cast<float>(D));
    f.dgamma[c] = static_cast<float>(dgamma);
    f.dbeta[c] = static_cast<float>(s1);
    f.dbias[c] = 0.f;
}
__host__ __device__ inline int fold6_coef_blocks(int Co, int Ci) { return (Co + 7) / 8 + (Ci * Ci + 255) / 256; }

// block roles: dW (thread per element) | scaled transposed weights (thread per element) | S: one block per (i, 32 columns) |
// per-cloud constant rows: one block per (cloud, 32 columns)
__host__ __device__ inline int fold6_bwd_blocks(int Co, int Ci, int clouds) { return 2 * ((Co * Ci + 255) / 256) + (Ci + clouds) * (Ci / 32); }
__global__ void __launch_bounds__(256) k_fold6_bwd(const Fold6Args f) {
    pdl_launch_dependents();
    pdl_wait();
    const int Co = f.Co, Ci = f.Ci;
    const int nA = (Co * Ci + 255) / 256;
    int blk = blockIdx.x;
    if (blk < nA) {                                   // dW[c][k]
        const int idx = blk * 256 + threadIdx.x;
        if (idx >= Co * Ci) return;
        const int c = idx / Ci, k = idx - c * Ci;
        const __nv_bfloat16* w = f.W + static_cast<size_t>(c) * Ci;
        float wg = 0.f;
#pragma unroll 8
        for (int j = 0; j < Ci; ++j) wg = fmaf(__bfloat162float(w[j]), f.gsum[j * Ci + k], wg);
        const float4 cf = f.coef[c];
        const float4 bp = f.bnp[c];
        const float mean = -bp.w / bp.z;
        float cbs = 0.f;                 // sum_b (mean of y over cloud b - mean) s_b[k]
        for (int b = 0; b < f.clouds; ++b) {
            const float mu = static_cast<float>(f.part[(static_cast<size_t>(b) * Co + c) * 2]) / f.N + f.cb[static_cast<size_t>(b) * Co + c] - mean;
            cbs = fmaf(mu, static_cast<float>(f.s[static_cast<size_t>(b) * Ci + k]), cbs);
        }
        const float sk = static_cast<float>(f.ssum[k]);
        const float yca = wg + cbs;                                                   // sum_p (y - mean)[p][c] a[p][k]
        f.dW[static_cast<size_t>(c) * f.ld_dw + k] = fmaf(cf.x, f.Q[idx], fmaf(cf.y, yca, cf.w * sk));
        return;
    }
    blk -= nA;
    if (blk < nA) {                                   // wcat[j][col0 + c] = A_c W[c][j]
        const int idx = blk * 256 + threadIdx.x;
        if (idx >= Co * Ci) return;
        const int j = idx / Co, c = idx - j * Co;
        f.wcat[static_cast<size_t>(j) * f.ld_wcat + f.col0 + c] = __float2bfloat16_rn(f.coef[c].x * __bfloat162float(f.W[static_cast<size_t>(c) * Ci + j]));
        return;
    }
    blk -= nA;
    const int lane = threadIdx.x & 31;
    const int jb = Ci / 32;
    if (blk < Ci * jb) {                              // wcat[j][col0 + Co + i] = S[i][j]
        const int i = blk / jb, j0 = (blk - i * jb) * 32;
        const float r = block_weighted_colsum32(f.W, Ci, Co, j0, [&](int c) { return __bfloat162float(f.W[static_cast<size_t>(c) * Ci + i]) * f.coef[c].y; });
        if (threadIdx.x < 32) f.wcat[static_cast<size_t>(j0 + lane) * f.ld_wcat + f.col0 + Co + i] = __float2bfloat16_rn(r);
    } else {                                          // cst[b][j] = sum_c (Bc_c (cb_b[c] - mean_c) + D_c) W[c][j]
        const int o = blk - Ci * jb;
        const int b = o / jb, j0 = (o - b * jb) * 32;
        const float r = block_weighted_colsum32(f.W, Ci, Co, j0, [&](int c) {
            const float4 cf = f.coef[c];
            const float4 bp = f.bnp[c];
            return fmaf(cf.y, f.cb[static_cast<size_t>(b) * Co + c] + bp.w / bp.z, cf.w);
        });
        if (threadIdx.x < 32) f.cst[static_cast<size_t>(b) * Ci + j0 + lane] = r;
    }
}

// ---------------------------------------------------------------------------------------------
// Folded BatchNorm backward of global_feat (Ci = Co = 1024), whose dz is the max-pool gradient: one non-zero per (cloud,
// channel) at the arg-max row (pcs.py:114).  sum dz and sum dz*yhat come from k_cloud_bwd_dg, so the coefficients need no Q.
//   k_fold5_prep : coefficients {A, Bc, D - Bc mean, D}, dgamma / dbeta / dbias, WB = diag(Bc) W (bf16, the A operand of the
//                  tcgen05 GEMM S = WB^T W) and const = (D - Bc mean)^T W.   8 channels per block.
//   k_pool_claim : every arg-max row gets ONE slot of the side buffer (lowest (cloud, channel) index that routes to it).
//   k_pool_rows_*: side[slot] = sum_{c -> row} A_c dzv[b][c] W[c][:]  (the rows of dz diag(A) W, added to the data-gradient
//                  accumulator by the GEMM epilogue) and Q[c][:] = sum_b dzv[b][c] a_prev[row][:]  (= dz^T a_prev).
//   k_gram_center: Gc = G - s s^T / n from the upper triangle of the Gram GEMM, bf16 (B operand of the W Gc GEMM).
// ---------------------------------------------------------------------------------------------
struct Fold5Args {
    const double* stats_b;        // [2][Co] {sum dz, sum dz*yhat}
    const float4* bnp;            // [Co]
    float4* coef;                 // [Co]
    float* dgamma;
    float* dbeta;
    float* dbias;
    const __nv_bfloat16* W;       // [Co][Ci]
    __nv_bfloat16* WB;            // [Co][Ci] = Bc_c W[c][:]
    float* cst;                   // [Ci] (zeroed by the caller)
    double n;
    int Co, Ci;                   // Ci == 1024 (4 columns per thread)
};
__global__ void __launch_bounds__(256) k_fold5_prep(const Fold5Args f) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float bc_s[8], cz_s[8];
    const int c0 = blockIdx.x * 8;
    if (threadIdx.x < 8 && c0 + static_cast<int>(threadIdx.x) < f.Co) {
        const int c = c0 + threadIdx.x;
        const float4 bp = f.bnp[c];
        const double s1 = f.stats_b[c], dgamma = f.stats_b[f.Co + c];
        const double A = bp.x;
        const double Bc = -A * static_cast<double>(bp.z) * dgamma / f.n;
        const double D = -A * s1 / f.n;
        const double mean = -static_cast<double>(bp.w) / static_cast<double>(bp.z);
        f.coef[c] = make_float4(static_cast<float>(A), static_cast<float>(Bc), static_cast<float>(D - Bc * mean), static_cast<float>(D));
        f.dgamma[c] = static_cast<float>(dgamma);
        f.dbeta[c] = static_cast<float>(s1);
        f.dbias[c] = 0.f;
        bc_s[threadIdx.x] = static_cast<float>(Bc);
        cz_s[threadIdx.x] = static_cast<float>(D - Bc * mean);
    }
    __syncthreads();
    const int j = threadIdx.x * 4;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
        const int c = c0 + ch;
        if (c >= f.Co) break;
        const uint2 w = *reinterpret_cast<const uint2*>(f.W + static_cast<size_t>(c) * f.Ci + j);
        const float w0 = bf16_lo(w.x), w1 = bf16_hi(w.x), w2 = bf16_lo(w.y), w3 = bf16_hi(w.y);
        const float bc = bc_s[ch], cz = cz_s[ch];
        *reinterpret_cast<uint2*>(f.WB + static_cast<size_t>(c) * f.Ci + j) = make_uint2(pack_bf16x2(bc * w0, bc * w1), pack_bf16x2(bc * w2, bc * w3));
        acc[0] = fmaf(cz, w0, acc[0]);
        acc[1] = fmaf(cz, w1, acc[1]);
        acc[2] = fmaf(cz, w2, acc[2]);
        acc[3] = fmaf(cz, w3, acc[3]);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) atomicAdd(f.cst + j + e, acc[e]);
}

__global__ void __launch_bounds__(256) k_pool_claim(const float* __restrict__ dzv, const int* __restrict__ argidx, int total, int C,
                                                    int N, int* __restrict__ rowslot) {
    pdl_launch_dependents();
    pdl_wait();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total || dzv[idx] == 0.f) return;
    const int b = idx / C;
    atomicMin(rowslot + static_cast<size_t>(b) * N + argidx[idx], idx);
}

// Max-pool gradient rows, two launches (no zero-fill of the side buffer, no atomics on Q):
//   k_pool_rows_own  : one block (128 threads x 8 columns) per CHANNEL c, looping over the clouds: Q[c][:] = sum_b dzv[b][c]
//                      a_prev[row(b,c)][:] (plain store), and for every (b, c) that owns the slot of its row the slot is
//                      INITIALISED with its own contribution side[slot] = A_c dzv[b][c] W[c][:] (plain store).
//   k_pool_rows_add  : one block per (cloud, channel) that routes to a row owned by another channel: vector red.add.
// A slot that nobody owns is never read (rowslot says so), so stale contents are harmless.  C == Ci == 1024.
__global__ void __launch_bounds__(128) k_pool_rows_own(const float* __restrict__ dzv, const int* __restrict__ argidx, int clouds, int C,
                                                       int N, const int* __restrict__ rowslot, const float4* __restrict__ coef,
                                                       const __nv_bfloat16* __restrict__ W, const __nv_bfloat16* __restrict__ a_prev,
                                                       float* __restrict__ side, float* __restrict__ Q) {
    pdl_launch_dependents();
    pdl_wait();
    const int c = blockIdx.x;
    const int j = threadIdx.x * 8;
    const uint4 w = *reinterpret_cast<const uint4*>(W + static_cast<size_t>(c) * C + j);
    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
    const float Ac = coef[c].x;
    float q[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) q[e] = 0.f;
    for (int b = 0; b < clouds; ++b) {
        const int idx = b * C + c;
        const float v = dzv[idx];
        if (v == 0.f) continue;                                   // (block-uniform)
        const size_t row = static_cast<size_t>(b) * N + argidx[idx];
        const uint4 a = *reinterpret_cast<const uint4*>(a_prev + row * C + j);
        const uint32_t as[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            q[2 * e] = fmaf(v, bf16_lo(as[e]), q[2 * e]);
            q[2 * e + 1] = fmaf(v, bf16_hi(as[e]), q[2 * e + 1]);
        }
        if (rowslot[row] == idx) {
            const float alpha = Ac * v;
            float4* e4 = reinterpret_cast<float4*>(side + static_cast<size_t>(idx) * C + j);
            e4[0] = make_float4(alpha * bf16_lo(ws[0]), alpha * bf16_hi(ws[0]), alpha * bf16_lo(ws[1]), alpha * bf16_hi(ws[1]));
            e4[1] = make_float4(alpha * bf16_lo(ws[2]), alpha * bf16_hi(ws[2]), alpha * bf16_lo(ws[3]), alpha * bf16_hi(ws[3]));
        }
    }
    float4* q4 = reinterpret_cast<float4*>(Q + static_cast<size_t>(c) * C + j);
    q4[0] = make_float4(q[0], q[1], q[2], q[3]);
    q4[1] = make_float4(q[4], q[5], q[6], q[7]);
}
__global__ void __launch_bounds__(128) k_pool_rows_add(const float* __restrict__ dzv, const int* __restrict__ argidx, int C, int N,
                                                       const int* __restrict__ rowslot, const float4* __restrict__ coef,
                                                       const __nv_bfloat16* __restrict__ W, float* __restrict__ side) {
    pdl_launch_dependents();
    pdl_wait();
    const int idx = blockIdx.x;
    const float v = dzv[idx];
    if (v == 0.f) return;
    const int b = idx / C, c = idx - b * C;
    const int slot = rowslot[static_cast<size_t>(b) * N + argidx[idx]];
    if (slot == idx) return;                                      // owner: written by k_pool_rows_own
    const float alpha = coef[c].x * v;
    const int j = threadIdx.x * 8;
    const uint4 w = *reinterpret_cast<const uint4*>(W + static_cast<size_t>(c) * C + j);
    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
    float* e = side + static_cast<size_t>(slot) * C + j;
#pragma unroll
    for (int h = 0; h < 2; ++h)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(e + 4 * h), "f"(alpha * bf16_lo(ws[2 * h])),
                     "f"(alpha * bf16_hi(ws[2 * h])), "f"(alpha * bf16_lo(ws[2 * h + 1])), "f"(alpha * bf16_hi(ws[2 * h + 1]))
                     : "memory");
}

__global__ void __launch_bounds__(256) k_gram_center(const float* __restrict__ G, const double* __restrict__ s, double n, int K,
                                                     __nv_bfloat16* __restrict__ Gc) {
    pdl_launch_dependents();
    pdl_wait();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= K * K) return;
    const int i = idx / K, j = idx - i * K;
    const int lo = min(i, j), hi = max(i, j);                 // only the upper triangle of G is computed
    Gc[idx] = __float2bfloat16_rn(static_cast<float>(static_cast<double>(G[static_cast<size_t>(lo) * K + hi]) - s[i] * s[j] / n));
}

// ---------------------------------------------------------------------------------------------
// Backward of the per-cloud seg_conv1 branch and of the max-pool.
//   dg[b][j]   = sum_n dcb[b][n] * Wg[n][j]                 (repeat/cat backward folded per cloud)
//   dWg[n][j] += sum_b dcb[b][n] * g[b][j]
//   dzv[b][j]  = dg[b][j] * (g[b][j] > 0)                   (relu + max routing value)
//   stats6     = {sum_b dzv, sum_b dzv * yhat(ystar)}       (BN backward sums of the sparse gradient)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_cloud_bwd_dg(const float* __restrict__ dcb, const float* __restrict__ Wg, int ldw, int clouds,
                                                      int Nn /*512*/, int J /*1024*/, const float* __restrict__ g,
                                                      const float* __restrict__ ystar, const float4* __restrict__ bnp6,
                                                      float* __restrict__ dzv, double* __restrict__ stats6) {
    pdl_launch_dependents();
    pdl_wait();
    // grid: (J/32, clouds); block 256 = 8 warps; lane -> output j, warp -> 1/8 of the Nn-long reduction.
    // stats6 must be zeroed by the caller.
    __shared__ float part[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + lane;
    const int b = blockIdx.y;
    const float* d = dcb + static_cast<size_t>(b) * Nn;
    const int n0 = warp * (Nn / 8), n1 = n0 + Nn / 8;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    for (int n = n0; n < n1; n += 4) {
        s0 = fmaf(d[n], Wg[static_cast<size_t>(n) * ldw + j], s0);
        s1 = fmaf(d[n + 1], Wg[static_cast<size_t>(n + 1) * ldw + j], s1);
        s2 = fmaf(d[n + 2], Wg[static_cast<size_t>(n + 2) * ldw + j], s2);
        s3 = fmaf(d[n + 3], Wg[static_cast<size_t>(n + 3) * ldw + j], s3);
    }
    part[warp][lane] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (warp == 0) {
        float s = 0.f;
#pragma unroll
        for (int wv = 0; wv < 8; ++wv) s += part[wv][lane];
        const float4 bp = __ldg(bnp6 + j);
        const size_t o = static_cast<size_t>(b) * J + j;
        const float v = (g[o] > 0.f) ? s : 0.f;
        dzv[o] = v;
        if (v != 0.f) {
            atomicAdd(stats6 + j, static_cast<double>(v));
            atomicAdd(stats6 + J + j, static_cast<double>(v) * static_cast<double>(fmaf(bp.z, ystar[o], bp.w)));
        }
    }
}
__global__ void __launch_bounds__(256) k_cloud_bwd_dw(const float* __restrict__ dcb, const float* __restrict__ g, int clouds, int Nn,
                                                      int J, float* __restrict__ dWg, int ldw) {
    pdl_launch_dependents();
    pdl_wait();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = blockIdx.y;
    if (j >= J) return;
    float s = 0.f;
    for (int b = 0; b < clouds; ++b) s = fmaf(dcb[static_cast<size_t>(b) * Nn + n], g[static_cast<size_t>(b) * J + j], s);
    dWg[static_cast<size_t>(n) * ldw + j] += s;
}

// dW1[c][k] = sum_p dy1[p][c] * x[p][k]   (64 x 4).  8 threads per point, 8 channels each.
__global__ void __launch_bounds__(256) k_ingest_bwd(const __nv_bfloat16* __restrict__ dy1, const float4* __restrict__ x, long P,
                                                    float* __restrict__ dW1) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float red[32][64 * 4];
    const int cg = threadIdx.x & 7, slot = threadIdx.x >> 3;
    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[j][k] = 0.f;
    for (long pnt = blockIdx.x * 32 + slot; pnt < P; pnt += static_cast<long>(gridDim.x) * 32) {
        const float4 xv = __ldg(x + pnt);
        const uint4 dw = *reinterpret_cast<const uint4*>(dy1 + pnt * 64 + cg * 8);
        const uint32_t ds[4] = {dw.x, dw.y, dw.z, dw.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float d = (j & 1) ? bf16_hi(ds[j >> 1]) : bf16_lo(ds[j >> 1]);
            acc[j][0] = fmaf(d, xv.x, acc[j][0]);
            acc[j][1] = fmaf(d, xv.y, acc[j][1]);
            acc[j][2] = fmaf(d, xv.z, acc[j][2]);
            acc[j][3] = fmaf(d, xv.w, acc[j][3]);
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) red[slot][(cg * 8 + j) * 4 + k] = acc[j][k];
    __syncthreads();
    {
        const int i = threadIdx.x;   // 256 = 64*4 outputs
        float s = 0.f;
        for (int sl = 0; sl < 32; ++sl) s += red[sl][i];
        atomicAdd(dW1 + i, s);
    }
}

// ---------------------------------------------------------------------------------------------
// Fused Adam (torch.optim.Adam semantics: L2 weight decay added to the gradient, bias correction).
// grad_scale lets data-parallel ranks fold an averaging factor in (1.0 when grads are already summed); grad_div (device,
// optional) divides the gradient as well: data-parallel ranks back-propagate the UN-normalised loss and divide by the
// all-reduced sum of class weights here, which takes that all-reduce off the critical path (backward is linear in the loss).
// ---------------------------------------------------------------------------------------------
struct StepState {                 // == pcseg_step_state
    unsigned long long seed;
    long long step;
    float lr, bias_corr1, bias_corr2_sqrt, reserved;
};
__global__ void k_step_advance(StepState* st, float b1, float b2) {
    pdl_launch_dependents();
    pdl_wait();
    st->seed += 0x9E3779B97F4A7C15ull;
    const long long t = st->step + 1;
    st->step = t;
    st->bias_corr1 = static_cast<float>(1.0 - pow(static_cast<double>(b1), static_cast<double>(t)));
    st->bias_corr2_sqrt = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(b2), static_cast<double>(t))));
}
__global__ void __launch_bounds__(256) k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                              float* __restrict__ v, long n, float lr, float b1, float b2, float eps, float wd,
                                              float bc1, float bc2_sqrt, float grad_scale, const StepState* __restrict__ st,
                                              const double* __restrict__ grad_div) {
    pdl_launch_dependents();
    pdl_wait();
    if (grad_div != nullptr) grad_scale = static_cast<float>(static_cast<double>(grad_scale) / *grad_div);
    if (st != nullptr) {
        lr = st->lr;
        bc1 = st->bias_corr1;
        bc2_sqrt = st->bias_corr2_sqrt;
    }
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
        const float pv = p[i];
        const float gv = fmaf(wd, pv, g[i] * grad_scale);
        const float mv = fmaf(1.f - b1, gv - m[i], m[i]);
        const float vv = fmaf(1.f - b2, gv * gv - v[i], v[i]);
        m[i] = mv;
        v[i] = vv;
        const float denom = sqrtf(vv) / bc2_sqrt + eps;
        p[i] = pv - (lr / bc1) * (mv / denom);
    }
}

// Validation metrics in one pass over the logits (pcs.py:292-304, 319-343): weighted-CE sums, accuracy counters and the
// C x C confusion matrix (rows = true class, columns = predicted class) from which F1 scores follow on the host.
// One thread per point; block-level reduction, then one atomic per counter per block.
template <int MAXC>
__global__ void __launch_bounds__(256) k_eval_metrics(const float* __restrict__ logits, const long long* __restrict__ labels, long P, int C,
                                                      const float* __restrict__ class_w, CeAccum* __restrict__ ce,
                                                      unsigned long long* __restrict__ confusion, long long* __restrict__ pred_out) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ unsigned int conf_s[MAXC * MAXC];
    __shared__ double red_d[8][2];
    __shared__ unsigned long long red_u[8][2];
    for (int i = threadIdx.x; i < MAXC * MAXC; i += blockDim.x) conf_s[i] = 0u;
    __syncthreads();
    double loss_num = 0.0, w_sum = 0.0;
    unsigned long long correct = 0, nvalid = 0;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < P; i += static_cast<long>(gridDim.x) * blockDim.x) {
        float z[MAXC];
        float zmax = -INFINITY;
        int am = 0;
#pragma unroll
        for (int k = 0; k < MAXC; ++k) {
            z[k] = (k < C) ? logits[i * C + k] : -INFINITY;
            if (z[k] > zmax) { zmax = z[k]; am = k; }          // first maximum, like torch.argmax
        }
        if (pred_out != nullptr) pred_out[i] = am;
        const long long lab = labels != nullptr ? labels[i] : -1;
        if (lab >= 0 && lab < C) {      // labels outside [0, C) would be an error in the reference; here they are ignored
            float se = 0.f, zl = 0.f;
#pragma unroll
            for (int k = 0; k < MAXC; ++k)
                if (k < C) { se += __expf(z[k] - zmax); if (k == lab) zl = z[k]; }
            const float wl = class_w ? class_w[lab] : 1.f;
            loss_num += static_cast<double>(wl) * static_cast<double>(zmax + logf(se) - zl);
            w_sum += wl;
            correct += (am == lab);
            nvalid += 1;
            atomicAdd(&conf_s[static_cast<int>(lab) * MAXC + am], 1u);
        }
    }
    for (int o = 16; o >= 1; o >>= 1) {
        loss_num += __shfl_xor_sync(0xffffffffu, loss_num, o);
        w_sum += __shfl_xor_sync(0xffffffffu, w_sum, o);
        correct += __shfl_xor_sync(0xffffffffu, correct, o);
        nvalid += __shfl_xor_sync(0xffffffffu, nvalid, o);
    }
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { red_d[warp][0] = loss_num; red_d[warp][1] = w_sum; red_u[warp][0] = correct; red_u[warp][1] = nvalid; }
    __syncthreads();
    if (threadIdx.x == 0 && ce != nullptr) {
        double a0 = 0.0, a1 = 0.0;
        unsigned long long u0 = 0, u1 = 0;
        for (int w = 0; w < 8; ++w) { a0 += red_d[w][0]; a1 += red_d[w][1]; u0 += red_u[w][0]; u1 += red_u[w][1]; }
        atomicAdd(&ce->loss_num, a0);
        atomicAdd(&ce->w_sum, a1);
        atomicAdd(&ce->correct, u0);
        atomicAdd(&ce->valid, u1);
    }
    if (confusion != nullptr)
        for (int i = threadIdx.x; i < C * C; i += blockDim.x) {
            const unsigned int v = conf_s[(i / C) * MAXC + (i % C)];
            if (v) atomicAdd(confusion + i, static_cast<unsigned long long>(v));
        }
}

// ---------------------------------------------------------------------------------------------
// Ragged (un-padded) execution of a zero-padded batch (reference collate_fn, pcs.py:44-63).
// The real rows of every cloud are packed into a matrix whose clouds start at multiples of 128 rows, so that a GEMM
// tile never straddles two clouds.  Cloud b (length L, padded length Nmax) owns packed rows off[b] .. off[b+1]:
//   rows 0 .. L-1        the real points                                     multiplicity 1
//   row  L   (L < Nmax)  ONE representative of the Nmax-L identical pad rows  multiplicity Nmax-L
//   rows after that      filler up to the next multiple of 128               multiplicity 0
// Filler rows are zero rows when the cloud has pad rows (they tie with the representative: max-pool unaffected, first
// index wins) and copies of the last real row otherwise (they tie with it).  Inference needs nothing else.  Training
// weights the representative row by its multiplicity in every reduction over points: k_stats_fix corrects the BN batch
// sums after the GEMM epilogues, and the backward pass carries the representative's gradient pre-multiplied (every
// backward op is linear in it; k_bn_bwd_apply<., true> scales the affine BN term).
// meta (int32, device): len[B] | off[B+1] | tile_cloud[rows/128]
// ---------------------------------------------------------------------------------------------
struct RaggedMeta {
    const int* len;
    const int* off;
    const int* tile_cloud;
    int B, Nmax, rows;
};

__global__ void __launch_bounds__(256) k_pack_rows(const float4* __restrict__ x, const long long* __restrict__ labels, const RaggedMeta m,
                                                   float4* __restrict__ xpack, long long* __restrict__ labpack,
                                                   float* __restrict__ rowmult) {
    pdl_launch_dependents();
    pdl_wait();
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < m.rows; r += gridDim.x * blockDim.x) {
        const int c = m.tile_cloud[r >> 7];
        const int i = r - m.off[c];
        const int L = m.len[c];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        long long lab = -1;
        float mult = 0.f;
        if (i < L) {
            const size_t src = static_cast<size_t>(c) * m.Nmax + i;
            v = __ldg(x + src);
            if (labels != nullptr) lab = labels[src];
            mult = 1.f;
        } else if (L < m.Nmax) {
            mult = (i == L) ? static_cast<float>(m.Nmax - L) : 0.f;
        } else {
            v = __ldg(x + static_cast<size_t>(c) * m.Nmax + (m.Nmax - 1));
        }
        xpack[r] = v;
        if (labpack != nullptr) labpack[r] = lab;
        if (rowmult != nullptr) rowmult[r] = mult;
    }
}

// packed logits -> the reference's padded (B, Nmax, C) tensor; every pad row receives its cloud's pad-row logits
__global__ void __launch_bounds__(256) k_unpack_logits(const float* __restrict__ lp, const RaggedMeta m, int C, float* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    const long total = static_cast<long>(m.B) * m.Nmax * C;
    for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long>(gridDim.x) * blockDim.x) {
        const long row = idx / C;
        const int k = static_cast<int>(idx - row * C);
        const int b = static_cast<int>(row / m.Nmax);
        const int i = static_cast<int>(row - static_cast<long>(b) * m.Nmax);
        const int src = m.off[b] + min(i, m.len[b]);
        out[idx] = lp[static_cast<size_t>(src) * C + k];
    }
}

// caller-supplied dlogits of the padded batch (B, Nmax, C) -> packed rows: real rows are copied, the representative pad
// row receives the SUM over its cloud's pad rows (gradients are carried pre-multiplied), filler rows zero.
// grid (ceil(rows / 256)) + B extra blocks: block (gridDim.x - B + b) reduces the pad rows of cloud b.
__global__ void __launch_bounds__(256) k_pack_dlogits(const float* __restrict__ dl, const RaggedMeta m, int C, float* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    const int copy_blocks = gridDim.x - m.B;
    if (static_cast<int>(blockIdx.x) < copy_blocks) {
        const int r = blockIdx.x * 256 + threadIdx.x;
        if (r >= m.rows) return;
        const int c = m.tile_cloud[r >> 7];
        const int i = r - m.off[c];
        const int L = m.len[c];
        if (i < L) {
            const float* src = dl + (static_cast<size_t>(c) * m.Nmax + i) * C;
            for (int k = 0; k < C; ++k) out[static_cast<size_t>(r) * C + k] = src[k];
        } else if (!(i == L && L < m.Nmax)) {
            for (int k = 0; k < C; ++k) out[static_cast<size_t>(r) * C + k] = 0.f;
        }
        return;
    }
    const int b = blockIdx.x - copy_blocks;
    const int L = m.len[b];
    if (L >= m.Nmax) return;
    __shared__ float red[256];
    for (int k = 0; k < C; ++k) {
        float s = 0.f;
        for (int i = L + threadIdx.x; i < m.Nmax; i += 256) s += dl[(static_cast<size_t>(b) * m.Nmax + i) * C + k];
        red[threadIdx.x] = s;
        __syncthreads();
        for (int w = 128; w > 0; w >>= 1) {
            if (static_cast<int>(threadIdx.x) < w) red[threadIdx.x] += red[threadIdx.x + w];
            __syncthreads();
        }
        if (threadIdx.x == 0) out[static_cast<size_t>(m.off[b] + L) * C + k] = red[0];
        __syncthreads();
    }
}

// BN batch sums of a packed layer: add (multiplicity - 1) * {y, y^2} of every non-real row (at most 128 per cloud).
// grid (C/32, B), block 256 = 32 columns x 8 row slots.
__global__ void __launch_bounds__(256) k_stats_fix(const __nv_bfloat16* __restrict__ y, int ld, int C, const RaggedMeta m,
                                                   const float* __restrict__ rowmult, double* __restrict__ stats) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ double red[2][8][32];
    const int lane = threadIdx.x & 31, slot = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + lane;
    const int b = blockIdx.y;
    const int r0 = m.off[b] + m.len[b], r1 = m.off[b + 1];
    if (r0 >= r1) return;                                   // (uniform over the block)
    double a1 = 0.0, a2 = 0.0;
    if (col < C) {
        for (int r = r0 + slot; r < r1; r += 8) {
            const double w = static_cast<double>(rowmult[r]) - 1.0;
            const double v = static_cast<double>(__bfloat162float(y[static_cast<size_t>(r) * ld + col]));
            a1 += w * v;
            a2 += w * v * v;
        }
    }
    red[0][slot][lane] = a1;
    red[1][slot][lane] = a2;
    __syncthreads();
    if (slot < 2 && col < C) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[slot][k][lane];
        atomicAdd(stats + slot * C + col, t);
    }
}

// argmax over classes (first maximum wins, like torch.argmax on ties)
__global__ void k_argmax(const float* __restrict__ logits, long P, int C, long long* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    const long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
    if (i >= P) return;
    const float* z = logits + i * C;
    float best = z[0];
    int bi = 0;
    for (int k = 1; k < C; ++k)
        if (z[k] > best) { best = z[k]; bi = k; }
    out[i] = bi;
}

}  // namespace pcseg
