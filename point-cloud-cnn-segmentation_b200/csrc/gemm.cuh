// Persistent, warp-specialised tcgen05 GEMM for the shared-MLP layers (sm_100a).
//
//   D[128 x BN] (fp32, TMEM, double buffered)  +=  A[128 x 64] * B[BN x 64]^T   per k-block
//
//   warp 0      : TMA producer (A and B tiles, 128B-swizzled, mbarrier complete_tx)
//   warp 1      : MMA issuer   (the warp walks the loop warp-uniformly, ONE elected lane issues tcgen05.mma / tcgen05.commit;
//                 the next stage's mbarrier is probed between the MMAs of the current one)
//   warp 2      : TMEM allocator
//   warps 4..11 : epilogue     (tcgen05.ld -> fused math -> swizzled smem -> TMA store / reductions); 4..19 for the
//                 stats + max-pool variant.  STATS / DGRAD epilogues work in two passes per 64-column sub-tile: pass 1 is
//                 row-mapped (TMEM lane = row) and only converts accumulators to bf16 into the staging tile, pass 2 is
//                 column-mapped over that tile (a warp owns 8 columns, per-column parameters in registers) and does the
//                 masking and the column reductions into per-CTA shared-memory accumulators (n_tile is fixed per CTA).
//
//   XF kernels only (forward GEMMs of the training step whose input still needs its train-mode BatchNorm):
//   warps 12..15: transform stage between TMA and MMA: the landed A tile holds PRE-BatchNorm values y of the layer below;
//                 these warps turn the batch sums of that layer into scale / shift once per CTA (CTA 0 also publishes them and
//                 the running statistics), then rewrite every A tile in place as a = relu(scale*y + shift) (* dropout), fence
//                 the async proxy and release the tile to the MMA warp (xf_bar).  The activation tensor is still needed by
//                 the backward pass, so
//   warp 3      : TMA-stores every transformed A tile to the activation tensor (tmA2) and only then lets the producer
//                 refill the stage (empty_bar counts the MMA commit AND this warp).
//                 The separate BN-apply kernel (one read of y, one write of a, one launch) disappears.
//
// Operand layouts:
//   MN == false : A is [M x K] row-major (K contiguous), B is [N x K] row-major  (forward, dgrad)
//   MN == true  : A is [K x M] row-major (M contiguous), B is [K x N] row-major  (wgrad: K = points)
//
// Fused epilogues (EPI):
//   EPI_BIAS_RELU : out = relu(acc + bias[n] (+ cloud_bias[cloud(row)][n]))          -> bf16 store
//   EPI_COLMAX    : per-cloud column max of relu(acc + bias[n])                       -> atomicMax (no store)
//   EPI_STATS     : out = bf16(acc (+ cloud_bias)); column sum / sum-of-squares       -> bf16 store + fp64 atomics
//   EPI_STATS_POOL: EPI_STATS + per-cloud arg-extremum of every column (train-mode max-pool, keys via atomicMax)
//   EPI_DGRAD     : dz = acc * relu'(bn(y)) * dropout; column sum dz, sum dz*yhat     -> bf16 store + fp64 atomics
//   EPI_WGRAD     : split-K partial tile                                              -> fp32 red.add
//   EPI_LOGITS    : logits = W4 * relu(acc + bias) + b4  (BN == 128 == all channels)  -> fp32 store
//   EPI_BN_RELU   : out = relu(scale[n] * acc + shift[n]) with {scale, shift} = bnp[n].xy (train-mode BatchNorm whose batch
//                   statistics were PREDICTED from the Gram matrix of the layer input, see k_predict_bn)   -> bf16 store
//   EPI_BN_RELU_DROP: EPI_BN_RELU with the per-cloud term added before the normalisation and Philox dropout after the ReLU
//                   (seg_conv1, pcs.py:117-124); a separate instantiation so that conv5's epilogue stays lean
//   EPI_DGRAD_ACT : EPI_DGRAD with the mask taken from the stored ACTIVATION of the layer below (a > 0 <=> ReLU on and not
//                   dropped; no BN parameters, no Philox); column sums: sum dz and sum a                 -> bf16 store + fp64 atomics
// DGRAD / DGRAD_ACT add bias[n] to the accumulator when p.bias != nullptr (constant row of the folded BatchNorm backward).
//
//   EPI_BIAS_RELU_X3: EPI_BIAS_RELU whose fp32 result is stored as a bf16 PAIR: hi = bf16(out) in columns [n, n + 64) and
//                   lo = bf16(out - hi) in columns [x3_lo_col + n, ...) of the output tensor (split-bf16 inference)
//
// Split-bf16 ("bf16x3") inference, p.x3_kb > 0: operands are stored as [hi | lo] column halves (x3_kb k-blocks each) and
// every logical k-block j is issued three times, (A_hi, B_hi), (A_hi, B_lo), (A_lo, B_hi): the fp32 accumulator then holds
// the product of the ~16-bit-significand operands (the dropped lo*lo term is 2^-16 smaller than the sum).
//
// The A operand may be the K-concatenation of two tensors: k-blocks [0, kb_switch) come from tmA, the rest from tmA2
// (folded BatchNorm backward: [dz | a_prev]).
//
// cloud(row): dense batches hold pts_per_cloud consecutive rows per cloud (row / pts_per_cloud; a tile may straddle two
// clouds); packed ragged batches (pointwise.cuh, k_pack_rows) start every cloud at a multiple of 128 rows and look the
// cloud of a tile up in tile_cloud[].  Either way the lookup happens once per tile, outside the column loop.
#pragma once
#include "ptx.cuh"
#include "bn.cuh"

namespace pcseg {

enum : int { EPI_BIAS_RELU = 0, EPI_COLMAX = 1, EPI_STATS = 2, EPI_DGRAD = 3, EPI_WGRAD = 4, EPI_LOGITS = 5, EPI_STATS_POOL = 6,
              EPI_BN_RELU = 7, EPI_DGRAD_ACT = 8, EPI_BIAS_RELU_X3 = 9, EPI_BN_RELU_DROP = 10 };

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
// Epilogue warps per kernel variant: 8 everywhere except the train-mode global_feat forward (stats + fused max-pool) and the
// seg_conv1 forward with dropout; measured per variant (tools/gpu_r2_ab.sh): EPI_DGRAD 16 warps 61.8 -> 70.0 us (seg_conv3 data
// gradient), EPI_DGRAD_ACT 16 warps 259 -> 272 us.  The global_feat forward's epilogue is the longest and measurably profits
// from 16 (282 vs 296 us); 16 everywhere was slower overall.
#ifndef PCSEG_DGRAD_ACT_WARPS
#define PCSEG_DGRAD_ACT_WARPS 8
#endif
#ifndef PCSEG_BN_RELU_DROP_WARPS
#define PCSEG_BN_RELU_DROP_WARPS 16     // seg_conv1 forward (K = 64: all epilogue, Philox dropout): 78.0 -> 74.8 us with 16
#endif
#ifndef PCSEG_DGRAD_WARPS
#define PCSEG_DGRAD_WARPS 8
#endif
constexpr int epi_warps_for(int epi) {
    return epi == 6 /* EPI_STATS_POOL */ ? 16 : epi == 8 /* EPI_DGRAD_ACT */ ? PCSEG_DGRAD_ACT_WARPS
           : epi == 10 /* EPI_BN_RELU_DROP */ ? PCSEG_BN_RELU_DROP_WARPS : epi == 3 /* EPI_DGRAD */ ? PCSEG_DGRAD_WARPS : 8;
}
constexpr int MAX_CLASSES = 8;    // compile-time cap on num_classes for the fused head kernels

struct GemmParams {
    int M, N, K;                 // logical problem (for MN/wgrad: M = Cout, N = Cin, K = points)
    int num_m_tiles, num_n_tiles;
    int num_splits, kb_per_split;   // split-K (wgrad only; otherwise 1 / K/64)
    int kb_switch;               // k-blocks [0, kb_switch) of A come from tmA, the rest from tmA2 (0 = everything from tmA)
    int store_out;               // EPI_STATS_POOL: 0 = statistics / max-pool only, the output tile is not written
    int kb_group, splits_per_group;   // EPI_WGRAD, kb_group > 0: the k-blocks (points) form groups of kb_group (= one cloud) and every
                                 // group is split on its own: split s covers group s / splits_per_group (per-cloud Gram
                                 // matrices; with wg_mode 3 the partial tiles of a group are consecutive)
    float* cloud_sums;           // EPI_DGRAD_ACT: [clouds][N] per-cloud column sums of dz (fp32 atomics; tiles must not straddle
                                 // clouds), in addition to the batch sums in stats
    int x3_kb;                   // > 0: split-bf16 k-schedule, = logical K / 64 (K in this struct then counts 3 * logical K)
    int x3_lo_col;               // EPI_BIAS_RELU_X3: first column of the lo halves in the output tensor
    int sym_tiles;               // EPI_WGRAD with A == B (Gram matrix): > 0 = number of (m, n) tiles that touch the upper
                                 // triangle; only those are computed (tile order: n-tile major, m-tiles 0 .. (n+1)*BN/128-1)
    int wg_mode;                 // EPI_WGRAD: 0 = fp32 red.add into out_f32 (split-K), 1 = bf16 store into out_bf16 (one split),
                                 // 2 = folded weight gradient out_f32 = cA*wq + cB*acc + cD*ws[col] with {cA, cB, -, cD} = wcoef[row]
                                 // 3 = every split stores its partial tile to out_f32 + split*M*ldc (summed in a fixed order
                                 //     by k_gram_reduce: deterministic, unlike the atomics of mode 0)
    __nv_bfloat16* out_bf16;     // wg_mode 1
    const float* wq;             // wg_mode 2: Q = dz^T a_prev [M][ldc]
    const float4* wcoef;         // wg_mode 2: per output row
    const double* ws;            // wg_mode 2: column sums of a_prev [N]
    const int* rowslot;          // DGRAD / DGRAD_ACT: per-row index into `side` (>= side_rows: none), or nullptr
    const float* side;           // [side_rows][N] fp32 rows added to the accumulator before masking (max-pool gradient rows)
    int side_rows;
    // epilogue operands (all optional depending on EPI)
    const float* bias;           // [N]
    const float* cloud_bias;     // [clouds][N] or nullptr
    int pts_per_cloud;           // rows per cloud (for cloud_bias / colmax)
    const int* tile_cloud;       // ragged (packed) execution: cloud of every 128-row tile, or nullptr (dense: row / pts_per_cloud)
    const int* cloud_off;        // ragged: first packed row of every cloud (tiles never straddle clouds)
    unsigned int* colmax;        // [clouds][N] float bits (values >= 0)
    double* stats;               // [2][N]
    const float4* bnp;           // [N] {scale, shift, invstd, -mean*invstd} of the layer whose output is being masked
    float* out_f32;              // wgrad destination
    int ldc;                     // wgrad destination pitch (elements)
    unsigned long long seed;     // dropout (effective seed = seed + *seed_ptr when seed_ptr != nullptr)
    const unsigned long long* seed_ptr;
    unsigned int drop_thr16;     // 0 = no dropout
    float keep_scale;            // 1/(1-p)
    const float* w4;             // [C][128] fp32
    const float* b4;             // [C]
    int num_classes;
    float* logits;               // [M][C]
    const float* gamma;          // [N] BN weight: its sign picks max / min for the fused train-mode max-pool
    unsigned long long* pool_keys;   // [clouds][N] packed (orderable value << 32 | ~row) arg-extremum keys
    // XF kernels (transform stage, see gemm_kernel): tmA is the PRE-BatchNorm tensor y of the layer below, tmA2 its activation
    // tensor (written by the kernel); xf_fin = that layer's BatchNorm (batch sums -> scale / shift, running statistics)
    BnFinalizeArgs xf_fin;
    unsigned long long xf_seed;  // dropout of the activation (effective seed = xf_seed + *seed_ptr), xf_thr16 == 0: none
    unsigned int xf_thr16;
    float xf_keep_scale;
    double* xf_colsum;           // [clouds][64] per-cloud column sums of the stored activation (K == 64 only), or nullptr
};

template <int BN, int EPI, bool MN, bool XF = false, bool C2 = false>
struct GemmCfg {
    static_assert(!C2 || (!MN && !XF && BN == 256), "CTA-pair kernels: K-major, BN = 256, no transform stage");
    static constexpr int EPI_WARPS = epi_warps_for(EPI);
    static constexpr int EPI_THREADS = EPI_WARPS * 32;
#ifndef PCSEG_XF_WARPS
#define PCSEG_XF_WARPS 4
#endif
    static constexpr int XF_WARPS = XF ? PCSEG_XF_WARPS : 0;      // transform warps behind the epilogue warps (4 or 8)
    static constexpr int XF_THREADS = XF_WARPS * 32;
    static constexpr int THREADS = 128 + EPI_THREADS + XF_THREADS;
    static_assert(!XF || (EPI == EPI_STATS && !MN && BN <= 256), "transform stage: K-major EPI_STATS kernels only");
    static constexpr int STAGE_A = GEMM_BM * GEMM_BK * 2;
    static constexpr int STAGE_B = (C2 ? BN / 2 : BN) * GEMM_BK * 2;      // CTA pair: each CTA holds half of the B tile
    static constexpr int STAGE = STAGE_A + STAGE_B;
    static constexpr bool HAS_OUT = (EPI == EPI_BIAS_RELU || EPI == EPI_STATS || EPI == EPI_STATS_POOL || EPI == EPI_DGRAD ||
                                     EPI == EPI_BN_RELU || EPI == EPI_DGRAD_ACT || EPI == EPI_BIAS_RELU_X3 || EPI == EPI_BN_RELU_DROP);
    static constexpr bool HAS_Y = (EPI == EPI_DGRAD || EPI == EPI_DGRAD_ACT);
    // EPI_STATS_POOL (K = 1024, the mainloop is bound by the bytes it can keep in flight): ONE staging tile instead of two
    // buys a 4th 48 KB pipeline stage; the price is a second barrier per 64-column sub-tile
#ifndef PCSEG_POOL_SINGLE_STAGING
#define PCSEG_POOL_SINGLE_STAGING 1
#endif
    static constexpr bool SINGLE_OUT = (EPI == EPI_STATS_POOL) && PCSEG_POOL_SINGLE_STAGING != 0;
    static constexpr int OUT_BYTES = HAS_OUT ? (SINGLE_OUT ? 16384 : 2 * 16384) : 0;
    static constexpr int Y_BYTES = HAS_Y ? 2 * 16384 : 0;
    static constexpr int COMB_BYTES = (EPI == EPI_LOGITS) ? (epi_warps_for(EPI) / 4 - 1) * 128 * MAX_CLASSES * 4 + 1024
                                      : (EPI == EPI_STATS_POOL) ? 4096 + BN * 4 + (epi_warps_for(EPI) / 8) * BN * 8   // + sign-flip word per
                                                                  // column + per-CTA arg-extremum keys [row group][column]
                                                                : 4096;   // column accumulators / colmax exchange / logits partials
    static constexpr int W4_BYTES = (EPI == EPI_LOGITS) ? (MAX_CLASSES * 128 + MAX_CLASSES) * 4 : 0;
    static constexpr int BAR_BYTES = 256;
    static constexpr int FIXED = OUT_BYTES + Y_BYTES + COMB_BYTES + W4_BYTES + BAR_BYTES;
    static constexpr int BUDGET = 232448 - 1024;
    static constexpr int STAGES_RAW = (BUDGET - FIXED) / STAGE;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE + FIXED;
    static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                                   : (2 * BN <= 256) ? 256 : 512;
    static_assert(STAGES >= 2, "not enough shared memory for a pipeline");
    static_assert(2 * BN <= 512, "accumulator double buffer exceeds TMEM");
};

// Column-wise reduction of a 32 (lanes) x CW (registers) block.  CW == 32: lane j returns op over lanes of v[j];
// CW == 16: lane j returns column j >> 1.
template <int CW, bool IS_MAX>
__device__ __forceinline__ float warp_colreduce(float (&v)[CW]) {
    const uint32_t lane = lane_id();
    int cnt = CW / 2;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        if (cnt >= 1) {
            const bool up = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < CW / 2; ++i) {
                if (i < cnt) {
                    float send = up ? v[i] : v[i + cnt];
                    float keep = up ? v[i + cnt] : v[i];
                    float got = __shfl_xor_sync(0xffffffffu, send, off);
                    v[i] = IS_MAX ? fmaxf(keep, got) : (keep + got);
                }
            }
            cnt >>= 1;
        } else {
            float got = __shfl_xor_sync(0xffffffffu, v[0], off);
            v[0] = IS_MAX ? fmaxf(v[0], got) : (v[0] + got);
        }
    }
    return v[0];
}

// DGRAD epilogue, pass 2 for one thread: RPT rows x 8 columns of the staged dA tile are masked in place with the ReLU
// (and dropout) mask of the layer below and accumulated into s1 = sum dz, s2 = sum dz*y.
template <bool DROP, int RPT>
__device__ __forceinline__ void dgrad_pass2_rows(uint32_t tile_s, uint32_t ytile, int rbase, int chunk, const float (&sc)[8],
                                                 const float (&sh)[8], float (&s1)[8], float (&s2)[8], unsigned long long seed_eff,
                                                 unsigned int thr16, long long row0, int ncols, int colbase) {
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int r = rbase + 32 * i;
        const int off = r * 128 + ((chunk ^ (r & 7)) << 4);
        const uint4 dw = lds128(tile_s + off);
        const uint4 yw = lds128(ytile + off);
        const uint32_t ds[4] = {dw.x, dw.y, dw.z, dw.w};
        const uint32_t ys[4] = {yw.x, yw.y, yw.z, yw.w};
        uint32_t keep = 0xFFu;
        if (DROP) {
            const unsigned long long e0 = static_cast<unsigned long long>(row0 + r) * ncols + colbase;
            keep = dropout_keep8(seed_eff, e0 >> 3, thr16);
        }
        float dz[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float yv = (e & 1) ? bf16_hi(ys[e >> 1]) : bf16_lo(ys[e >> 1]);
            const float da = (e & 1) ? bf16_hi(ds[e >> 1]) : bf16_lo(ds[e >> 1]);
            bool on = fmaf(sc[e], yv, sh[e]) > 0.f;
            if (DROP) on = on && ((keep >> e) & 1u);
            dz[e] = on ? da : 0.f;
            s1[e] += dz[e];
            s2[e] = fmaf(dz[e], yv, s2[e]);
        }
        sts128(tile_s + off, make_uint4(pack_bf16x2(dz[0], dz[1]), pack_bf16x2(dz[2], dz[3]), pack_bf16x2(dz[4], dz[5]), pack_bf16x2(dz[6], dz[7])));
    }
}

// Train-mode max-pool keys of one row x 8 columns of a tile that straddles two clouds (rows per cloud not a multiple of
// 128; rare): one atomicMax per element.  xw = 8 bf16 values of the row, flip_s = shared-memory address of the 8 sign-flip
// words of the columns.  (Scalar arguments only: arrays by reference would force the caller's registers to local memory.)
__device__ __noinline__ void pool_keys_straddle(uint4 xw, uint32_t flip_s, int gr, int pts_per_cloud,
                                                unsigned long long* keys_col, int ncols) {
    const int cl = gr / pts_per_cloud;
    const uint32_t rlow = 0xFFFFFFFFu - static_cast<uint32_t>(gr - cl * pts_per_cloud);
    const uint32_t xs[4] = {xw.x, xw.y, xw.z, xw.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const uint32_t bits = (e & 1) ? (xs[e >> 1] & 0xFFFF0000u) : (xs[e >> 1] << 16);
        uint32_t fl;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(fl) : "r"(flip_s + 4u * e));
        const uint32_t bb = bits ^ fl;
        const uint32_t ord = (bb & 0x80000000u) ? ~bb : (bb | 0x80000000u);
        atomicMax(keys_col + static_cast<size_t>(cl) * ncols + e, (static_cast<unsigned long long>(ord) << 32) | rlow);
    }
}

// DGRAD_ACT epilogue, pass 2: the mask is (stored activation > 0); accumulates s1 = sum dz, s2 = sum activation.
template <int RPT>
__device__ __forceinline__ void dgrad_act_pass2_rows(uint32_t tile_s, uint32_t atile, int rbase, int chunk, float (&s1)[8],
                                                     float (&s2)[8]) {
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int r = rbase + 32 * i;
        const int off = r * 128 + ((chunk ^ (r & 7)) << 4);
        const uint4 dw = lds128(tile_s + off);
        const uint4 aw = lds128(atile + off);
        const uint32_t ds[4] = {dw.x, dw.y, dw.z, dw.w};
        const uint32_t as[4] = {aw.x, aw.y, aw.z, aw.w};
        float dz[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float av = (e & 1) ? bf16_hi(as[e >> 1]) : bf16_lo(as[e >> 1]);
            const float da = (e & 1) ? bf16_hi(ds[e >> 1]) : bf16_lo(ds[e >> 1]);
            dz[e] = (av > 0.f) ? da : 0.f;
            s1[e] += dz[e];
            s2[e] += av;
        }
        sts128(tile_s + off, make_uint4(pack_bf16x2(dz[0], dz[1]), pack_bf16x2(dz[2], dz[3]), pack_bf16x2(dz[4], dz[5]), pack_bf16x2(dz[6], dz[7])));
    }
}

template <int BN, int EPI, bool MN, bool XF = false, bool C2 = false>
__global__ void __launch_bounds__(GemmCfg<BN, EPI, MN, XF, C2>::THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmY,
            const GemmParams p) {
    using Cfg = GemmCfg<BN, EPI, MN, XF, C2>;
    constexpr int EPI_WARPS = Cfg::EPI_WARPS;
    constexpr int EPI_THREADS = Cfg::EPI_THREADS;
    constexpr int GEMM_THREADS = Cfg::THREADS;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int SUBS = BN / 64;           // 64-column epilogue sub-tiles
    static_assert(BN % 64 == 0, "BN must be a multiple of 64");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stage_base = smem;
    uint8_t* out_stage = smem + STAGES * Cfg::STAGE;                 // 2 x 16 KB (1024-aligned)
    uint8_t* y_stage = out_stage + Cfg::OUT_BYTES;                   // 2 x 16 KB
    float* comb = reinterpret_cast<float*>(y_stage + Cfg::Y_BYTES);  // [2][2][4][64]
    float* w4s = comb + Cfg::COMB_BYTES / 4;
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(w4s) + Cfg::W4_BYTES);
    uint64_t* full_bar = bars;                  // [STAGES]
    uint64_t* empty_bar = bars + STAGES;        // [STAGES]
    uint64_t* tmem_full = bars + 2 * STAGES;    // [2]
    uint64_t* tmem_empty = tmem_full + 2;       // [2]
    uint64_t* y_full = tmem_empty + 2;          // [2]
    uint64_t* xf_bar = y_full + 2;              // [STAGES] (XF only): A tile transformed
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(xf_bar + (XF ? STAGES : 0));
    static_assert((2 * STAGES + 6 + (XF ? STAGES : 0)) * 8 + 4 <= Cfg::BAR_BYTES, "barrier block too small");

    // (broadcast from lane 0: the compiler then knows the role branches and everything derived from them are warp-uniform
    //  and keeps barrier addresses, shared-memory descriptors and TMEM addresses in uniform registers)
    const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
    const uint32_t lane = lane_id();

    // CTA pairs (C2): a "tile" is 256 rows x BN columns, shared by the two CTAs of a cluster (rank r owns rows 128 r ..);
    // tile_start / tile_step enumerate tiles per CTA (1-CTA kernels) or per cluster
    const uint32_t pair_rank = C2 ? cluster_ctarank() : 0u;
    const int tile_start = C2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int tile_step = C2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    const int total_tiles = C2 ? (p.num_m_tiles >> 1) * p.num_n_tiles
                               : ((EPI == EPI_WGRAD && p.sym_tiles > 0) ? p.sym_tiles : p.num_m_tiles * p.num_n_tiles) * p.num_splits;

    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        if (p.kb_switch > 0) tma_prefetch_desc(&tmA2);
        tma_prefetch_desc(&tmB);
        if (Cfg::HAS_OUT) tma_prefetch_desc(&tmOut);
        if (Cfg::HAS_Y) tma_prefetch_desc(&tmY);
    }
    if (warp_idx == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], XF ? 2 : 1);      // XF: MMA commit + the warp that stores the transformed tile
            if (XF) mbar_init(&xf_bar[i], Cfg::XF_WARPS);          // one arrival per transform warp
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], C2 ? 2 * EPI_THREADS : EPI_THREADS);     // pair: the leader's barrier counts both epilogues
            mbar_init(&y_full[i], 1);
        }
        fence_barrier_init();
    }
    if (warp_idx == 2) {
        if (C2) {
            tmem_alloc_pair(tmem_ptr_smem, Cfg::TMEM_COLS);
            tmem_relinquish_pair();
        } else {
            tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
            tmem_relinquish();
        }
    }
    pdl_wait();          // everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the previous kernel
    if (EPI == EPI_LOGITS) {
        for (int i = threadIdx.x; i < p.num_classes * 128; i += GEMM_THREADS) w4s[i] = p.w4[i];
        for (int i = threadIdx.x; i < p.num_classes; i += GEMM_THREADS) w4s[MAX_CLASSES * 128 + i] = p.b4[i];
    }
    tc_fence_before();
    if (C2) cluster_sync_all();      // the peer's barriers are initialised before anything signals them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);

    auto tile_coords = [&](int tile, int& m_tile, int& n_tile, int& split) {
        split = tile % p.num_splits;
        int t = tile / p.num_splits;
        if (EPI == EPI_WGRAD && p.sym_tiles > 0) {
            n_tile = 0;
            for (;;) {
                const int cnt = min((n_tile + 1) * (BN / GEMM_BM), p.num_m_tiles);
                if (t < cnt) break;
                t -= cnt;
                ++n_tile;
            }
            m_tile = t;
            return;
        }
        n_tile = t % p.num_n_tiles;
        m_tile = t / p.num_n_tiles;
        if (C2) m_tile = 2 * m_tile + static_cast<int>(pair_rank);
    };

    auto kb_range = [&](int split, int& kb0, int& kb1) {
        const int total_kb = (p.K + GEMM_BK - 1) / GEMM_BK;
        if (EPI == EPI_WGRAD && p.kb_group > 0) {
            const int g = split / p.splits_per_group, l = split - g * p.splits_per_group;
            kb0 = g * p.kb_group + l * p.kb_per_split;
            kb1 = min(min(kb0 + p.kb_per_split, (g + 1) * p.kb_group), total_kb);
        } else {
            kb0 = split * p.kb_per_split;
            kb1 = min(kb0 + p.kb_per_split, total_kb);
        }
    };

    if (warp_idx == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = tile_start; tile < total_tiles; tile += tile_step) {
                int m_tile, n_tile, split;
                tile_coords(tile, m_tile, n_tile, split);
                int kb0, kb1;
                kb_range(split, kb0, kb1);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = stage_base + stage * Cfg::STAGE;
                    uint8_t* sb = sa + Cfg::STAGE_A;
                    if constexpr (C2) {
                        // both producers load their share; all bytes are counted on the leader's barrier, where the MMA waits
                        if (pair_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE);
                        const uint32_t lbar = mapa_shared(smem_u32(&full_bar[stage]), 0);
                        tma_load_2d_pair(sa, &tmA, lbar, kb * GEMM_BK, m_tile * GEMM_BM);
                        tma_load_2d_pair(sb, &tmB, lbar, kb * GEMM_BK, n_tile * BN + static_cast<int>(pair_rank) * (BN / 2));
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE);
                    if (!MN) {
                        int ka = kb, kbb = kb;                  // k-block columns of A and B
                        if (p.x3_kb > 0) {
                            const int j = kb / 3, t = kb - 3 * j;
                            ka = j + (t == 2 ? p.x3_kb : 0);
                            kbb = j + (t == 1 ? p.x3_kb : 0);
                        }
                        if (p.kb_switch > 0 && kb >= p.kb_switch)
                            tma_load_2d(sa, &tmA2, &full_bar[stage], (kb - p.kb_switch) * GEMM_BK, m_tile * GEMM_BM);
                        else
                            tma_load_2d(sa, &tmA, &full_bar[stage], ka * GEMM_BK, m_tile * GEMM_BM);
                        tma_load_2d(sb, &tmB, &full_bar[stage], kbb * GEMM_BK, n_tile * BN);
                    } else {
#pragma unroll
                        for (int j = 0; j < GEMM_BM / 64; ++j)
                            tma_load_2d(sa + j * 8192, &tmA, &full_bar[stage], m_tile * GEMM_BM + j * 64, kb * GEMM_BK);
#pragma unroll
                        for (int j = 0; j < BN / 64; ++j)
                            tma_load_2d(sb + j * 8192, &tmB, &full_bar[stage], n_tile * BN + j * 64, kb * GEMM_BK);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp_idx == 1) {
        // ------------------------------------------------------------ MMA issuer
        // Measured (tools/gpu_r2_prof.sh): the operands of a k-block were ALWAYS in shared memory already when this thread asked,
        // yet every mbarrier.try_wait cost ~90-100 cycles of latency during which the tensor pipe ran dry (tcgen05.mma issue
        // blocks until the pipe accepts the instruction, so nothing is queued behind it): 15-25 % of the loop.  The barrier of
        // the NEXT stage is therefore probed (non-blocking test_wait) between the MMAs of the current one, where its latency
        // hides behind their execution; the blocking wait remains only for stages that really are not there yet.
        constexpr uint32_t idesc = make_idesc_bf16(C2 ? 2 * GEMM_BM : GEMM_BM, BN, MN ? 1 : 0, MN ? 1 : 0);
        constexpr int KSTEPS = GEMM_BK / 16;
#ifndef PCSEG_MMA_PROBE_AHEAD
#define PCSEG_MMA_PROBE_AHEAD 1
#endif
        {
            // the whole warp walks the loop (warp-uniform control flow); ONE elected lane issues the MMAs and commits
#ifdef PCSEG_PROF_WAIT
            // diagnostic build: where does this warp wait -- for operands (full_bar) or for a free accumulator (tmem_empty)?
            long long prof_wait_acc = 0, prof_wait_full = 0, prof_issue = 0;
            int prof_ready = 0, prof_kb = 0;
            const long long prof_t0 = clock64();
#endif
            uint64_t* const op_bar = XF ? xf_bar : full_bar;
            int stage = 0;
            uint32_t phase = 0;
            int iter = 0;
            bool ready = false;            // op_bar[stage] is already known to have completed `phase`
            for (int tile = (C2 && pair_rank != 0) ? total_tiles : tile_start; tile < total_tiles; tile += tile_step, ++iter) {   // pair: leader only
                int m_tile, n_tile, split;
                tile_coords(tile, m_tile, n_tile, split);
                int kb0, kb1;
                kb_range(split, kb0, kb1);
                const int acc = iter & 1;
                const uint32_t acc_phase = (iter >> 1) & 1;
#ifdef PCSEG_PROF_WAIT
                const long long w0 = clock64();
#endif
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
#ifdef PCSEG_PROF_WAIT
                prof_wait_acc += clock64() - w0;
#endif
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = kb0; kb < kb1; ++kb) {
#ifdef PCSEG_PROF_WAIT
                    const long long w1 = clock64();
                    if (ready) ++prof_ready;
#endif
                    if (!ready) mbar_wait(&op_bar[stage], phase);
#ifdef PCSEG_PROF_WAIT
                    const long long w2 = clock64();
                    prof_wait_full += w2 - w1;
                    ++prof_kb;
#endif
                    tc_fence_after();
                    const int nstage = (stage + 1 == STAGES) ? 0 : stage + 1;
                    const uint32_t nphase = (stage + 1 == STAGES) ? (phase ^ 1) : phase;
                    const uint32_t sa = smem_u32(stage_base + stage * Cfg::STAGE);
                    const uint32_t sb = sa + Cfg::STAGE_A;
                    uint32_t probe = 0;
                    if (elect_one_sync()) {
#pragma unroll
                        for (int k = 0; k < KSTEPS; ++k) {
                            uint64_t da, db;
                            if (!MN) {
                                da = make_smem_desc_sw128(sa + k * 32, 0, 1024);
                                db = make_smem_desc_sw128(sb + k * 32, 0, 1024);
                            } else {
                                da = make_smem_desc_sw128(sa + k * 2048, 8192, 1024);
                                db = make_smem_desc_sw128(sb + k * 2048, 8192, 1024);
                            }
                            if (C2) umma_bf16_pair(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                            else umma_bf16(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                            if (PCSEG_MMA_PROBE_AHEAD && k == KSTEPS - 2) probe = mbar_test_wait(&op_bar[nstage], nphase) ? 1u : 0u;
                        }
                        if (C2) {
                            umma_commit_pair(&empty_bar[stage]);
                            if (kb == kb1 - 1) umma_commit_pair(&tmem_full[acc]);
                        } else {
                            umma_commit(&empty_bar[stage]);
                            if (kb == kb1 - 1) umma_commit(&tmem_full[acc]);
                        }
                    }
                    ready = __any_sync(0xffffffffu, probe != 0u);      // (the elected lane's probe, known to the whole warp)
#ifdef PCSEG_PROF_WAIT
                    prof_issue += clock64() - w2;
#endif
                    stage = nstage;
                    phase = nphase;
                }
                if (kb1 <= kb0 && elect_one_sync()) umma_commit(&tmem_full[acc]);   // empty split: still release the epilogue
                __syncwarp();
            }
#ifdef PCSEG_PROF_WAIT
            if (lane == 0 && (blockIdx.x == 0 || blockIdx.x == 77) && p.K >= 1024)
                printf("PROF gemm<%d,%d,%d> cta %d tiles %d: total %lld cyc, wait accumulator %lld, wait operands %lld, issue %lld, k-blocks %d ready %d\n",
                       BN, EPI, (int)MN, blockIdx.x, iter, clock64() - prof_t0, prof_wait_acc, prof_wait_full, prof_issue, prof_kb, prof_ready);
#endif
        }
        __syncwarp();
    } else if (XF && warp_idx == 3) {
        // ------------------------------------------------------------ activation store (XF): transformed A tiles -> tmA2
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = tile_start; tile < total_tiles; tile += tile_step) {
                int m_tile, n_tile, split;
                tile_coords(tile, m_tile, n_tile, split);
                int kb0, kb1;
                kb_range(split, kb0, kb1);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&xf_bar[stage], phase);
                    if (n_tile == 0) {      // (every n-tile of a row block transforms the same values; one of them stores)
                        tma_store_2d(&tmA2, stage_base + stage * Cfg::STAGE, kb * GEMM_BK, m_tile * GEMM_BM);
                        tma_store_commit();
                        tma_store_wait_read<0>();
                    }
                    mbar_arrive(&empty_bar[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
            tma_store_wait_all<0>();
        }
    } else if (XF && warp_idx >= 4 + EPI_WARPS) {
        // ------------------------------------------------------------ transform stage (XF): y tile -> relu(bn(y)) in place
        // thread = (16-byte chunk = 8 channels, row slot): rows rsub + 16 i.  A warp touches 4 rows x 128 B per access:
        // conflict-free under the 128-byte swizzle.
        constexpr int XT = Cfg::XF_THREADS > 0 ? Cfg::XF_THREADS : 128;
        constexpr int RSTEP = XT / 8, RPX = 128 / RSTEP;      // row step / rows per thread and k-block
        const int xt = threadIdx.x - (128 + Cfg::EPI_THREADS);      // 0..XF_THREADS-1
        const uint32_t tab_s = smem_u32(comb) + 2048;               // floats [0,256) scale, [256,512) shift
        float* tabf = comb + 512;
        for (int c = xt; c < p.K; c += XT) {
            const float4 bp = bn_from_stats(p.xf_fin, c);
            tabf[c] = bp.x;
            tabf[256 + c] = bp.y;
        }
        if (p.xf_colsum != nullptr && xt < 64) tabf[128 + xt] = 0.f;
        named_bar_sync(4, XT);
        const int chunk = xt & 7, rsub = xt >> 3;          // rows rsub + RSTEP * i
        const unsigned long long xseed = p.xf_seed + (p.seed_ptr != nullptr ? *p.seed_ptr : 0ull);
        float csum[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) csum[e] = 0.f;
        int cur_cl = -1;
        auto flush_colsum = [&](int cl) {     // per-cloud column sums of the stored activation (k_predict_bn_cloud's `s`)
            // lanes sharing a chunk -> shared-memory accumulators (K == 64: floats [128,192) of the table are free) -> one fp64
            // atomic per column and CTA.  (Measured: fp64 atomics straight from the lanes cost 20 us more per launch.)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                csum[e] += __shfl_xor_sync(0xffffffffu, csum[e], 8);
                csum[e] += __shfl_xor_sync(0xffffffffu, csum[e], 16);
            }
            if (lane < 8) {
#pragma unroll
                for (int e = 0; e < 8; ++e) atomicAdd(tabf + 128 + chunk * 8 + e, csum[e]);
            }
            named_bar_sync(4, XT);
            if (xt < 64) {
                atomicAdd(p.xf_colsum + static_cast<size_t>(cl) * 64 + xt, static_cast<double>(tabf[128 + xt]));
                tabf[128 + xt] = 0.f;
            }
            named_bar_sync(4, XT);
#pragma unroll
            for (int e = 0; e < 8; ++e) csum[e] = 0.f;
        };
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = tile_start; tile < total_tiles; tile += tile_step) {
            int m_tile, n_tile, split;
            tile_coords(tile, m_tile, n_tile, split);
            int kb0, kb1;
            kb_range(split, kb0, kb1);
            const int m0 = m_tile * GEMM_BM;
            if (p.xf_colsum != nullptr) {
                const int cl = p.pts_per_cloud > 0 ? m0 / p.pts_per_cloud : 0;     // (tiles do not straddle clouds: host check)
                if (cl != cur_cl) {
                    if (cur_cl >= 0) flush_colsum(cur_cl);
                    cur_cl = cl;
                }
            }
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                const uint32_t sa = smem_u32(stage_base + stage * Cfg::STAGE);
                const int col0 = kb * GEMM_BK + chunk * 8;
                const uint4 sc0 = lds128(tab_s + 4u * col0), sc1 = lds128(tab_s + 4u * col0 + 16);
                const uint4 sh0 = lds128(tab_s + 1024u + 4u * col0), sh1 = lds128(tab_s + 1024u + 4u * col0 + 16);
                const float sc[8] = {__uint_as_float(sc0.x), __uint_as_float(sc0.y), __uint_as_float(sc0.z), __uint_as_float(sc0.w),
                                     __uint_as_float(sc1.x), __uint_as_float(sc1.y), __uint_as_float(sc1.z), __uint_as_float(sc1.w)};
                const float sh[8] = {__uint_as_float(sh0.x), __uint_as_float(sh0.y), __uint_as_float(sh0.z), __uint_as_float(sh0.w),
                                     __uint_as_float(sh1.x), __uint_as_float(sh1.y), __uint_as_float(sh1.z), __uint_as_float(sh1.w)};
                uint4 yw[RPX];
#pragma unroll
                for (int i = 0; i < RPX; ++i) {
                    const int r = rsub + RSTEP * i;
                    yw[i] = lds128(sa + r * 128 + ((chunk ^ (r & 7)) << 4));
                }
#pragma unroll
                for (int i = 0; i < RPX; ++i) {
                    const int r = rsub + RSTEP * i;
                    const uint32_t ws[4] = {yw[i].x, yw[i].y, yw[i].z, yw[i].w};
                    uint32_t keep = 0xFFu;
                    if (p.xf_thr16 != 0u)
                        keep = dropout_keep8(xseed, (static_cast<unsigned long long>(m0 + r) * p.K + col0) >> 3, p.xf_thr16);
                    if (m0 + r >= p.M) keep = 0u;      // rows beyond M must stay exactly zero (statistics of the epilogue)
                    float o[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float yv = (e & 1) ? bf16_hi(ws[e >> 1]) : bf16_lo(ws[e >> 1]);
                        const float t = fmaf(sc[e], yv, sh[e]);
                        o[e] = (t > 0.f && ((keep >> e) & 1u)) ? t * p.xf_keep_scale : 0.f;
                    }
                    const uint4 pk = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
                    sts128(sa + r * 128 + ((chunk ^ (r & 7)) << 4), pk);
                    if (p.xf_colsum != nullptr) {
                        const uint32_t ps[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
                        for (int e = 0; e < 8; ++e) csum[e] += (e & 1) ? bf16_hi(ps[e >> 1]) : bf16_lo(ps[e >> 1]);
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&xf_bar[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
        if (p.xf_colsum != nullptr && cur_cl >= 0) flush_colsum(cur_cl);
        // later kernels (and the backward pass) read the published {scale, shift, invstd, -mean*invstd}; CTA 0 writes them
        // and the running statistics AFTER its tiles so that its fp64 divisions are off the kernel's critical path
        if (blockIdx.x == 0) bn_publish_by(p.xf_fin, xt, XT);
    } else if (warp_idx >= 4 && warp_idx < 4 + EPI_WARPS) {
        // ------------------------------------------------------------ epilogue (EPI_WARPS warps)
        // Pass 1 is row-mapped: warp w may only read TMEM lanes 32*(w%4)..+31, so the NQ warps that share a lane
        // quadrant split every 64-column sub-tile into NQ column groups of CW columns.
        // Pass 2 (STATS / DGRAD) is column-mapped over the bf16 staging tile and is free of that restriction.
        constexpr int NQ = EPI_WARPS / 4;             // column groups per sub-tile in pass 1
        constexpr int CW = 64 / NQ;                   // columns per thread in pass 1 (32 or 16)
        constexpr int NH = EPI_WARPS / 8;             // row groups in pass 2
        constexpr int RPT = 4 / NH;                   // rows per thread in pass 2
        const int ew = warp_idx & 3;                  // TMEM lane quadrant == row group
        const int cq = (warp_idx - 4) >> 2;           // column group of this warp
        const int row = ew * 32 + lane;               // row inside the tile
        const int et = threadIdx.x - 128;             // 0..EPI_THREADS-1
        const uint32_t lane_sel = static_cast<uint32_t>(ew * 32) << 16;
        const bool elected = (et == 0);
        const unsigned long long seed_eff = p.seed + (p.seed_ptr != nullptr ? *p.seed_ptr : 0ull);
        (void)seed_eff;
        const uint32_t out_s = smem_u32(out_stage);   // 32-bit shared-space addresses (explicit LDS/STS below)
        const uint32_t y_s = smem_u32(y_stage);
        const uint32_t comb_s = smem_u32(comb);
        int iter = 0;
        uint32_t sub_it = 0;                          // global 64-column sub-tile counter (buffer parity)
        // this thread has read everything it needs from accumulator buffer `a`: tell the MMA warp (pair: the leader's)
        const uint32_t tmem_empty_leader = C2 ? mapa_shared(smem_u32(&tmem_empty[0]), 0) : 0u;
        auto release_acc = [&](int a) {
            if (C2) mbar_arrive_cluster(tmem_empty_leader + 8u * a);
            else mbar_arrive(&tmem_empty[a]);
        };

        auto issue_y_load = [&](uint32_t it) {        // elected thread only
            const int t_ord = it / SUBS;
            const int sub = it % SUBS;
            const long tile = static_cast<long>(tile_start) + static_cast<long>(t_ord) * tile_step;
            if (tile >= total_tiles) return;
            int m_tile, n_tile, split;
            tile_coords(static_cast<int>(tile), m_tile, n_tile, split);
            const int buf = it & 1;
            mbar_arrive_expect_tx(&y_full[buf], 16384);
            tma_load_2d(y_stage + buf * 16384, &tmY, &y_full[buf], n_tile * BN + sub * 64, m_tile * GEMM_BM);
        };
        if (Cfg::HAS_Y && elected) {
            issue_y_load(0);
            issue_y_load(1);
        }
        constexpr bool COLACC = (EPI == EPI_STATS || EPI == EPI_STATS_POOL || EPI == EPI_DGRAD || EPI == EPI_DGRAD_ACT);
        constexpr bool IS_DGRAD = (EPI == EPI_DGRAD || EPI == EPI_DGRAD_ACT);
#ifndef PCSEG_TMEM_PREFETCH
#define PCSEG_TMEM_PREFETCH 1
#endif
        // (EPI_DGRAD keeps the simple order: its pass 2 is the register-hungriest and would spill with CW more live registers)
        constexpr bool PREFETCH = PCSEG_TMEM_PREFETCH != 0 && SUBS > 1 &&
                                  (EPI == EPI_STATS || EPI == EPI_STATS_POOL || EPI == EPI_BIAS_RELU || EPI == EPI_DGRAD_ACT);
        // (measured on cfg2: the whole step gains ~10 us with it; conv5 / seg_conv1 forward (EPI_BN_RELU[_DROP], HBM-write
        //  bound, no second pass to hide the load behind) lose 2.5 us each, so they keep the simple order)
        if constexpr (COLACC) {
            // per-CTA column accumulators [NH][2][BN]: valid because every tile of this CTA has the same n_tile
            // (the host launches a grid that is a multiple of num_n_tiles)
            for (int i = et; i < NH * 2 * BN; i += EPI_THREADS) comb[i] = 0.f;
            if constexpr (EPI == EPI_STATS_POOL) {
                // sign of the BN weight per column of this CTA (n_tile is fixed per CTA): the train-mode max-pool takes the
                // max of gamma >= 0 columns and the min of the others, i.e. the max after flipping the sign bit
                const int n0_fixed = (tile_start % p.num_n_tiles) * BN;
                for (int i = et; i < BN; i += EPI_THREADS)
                    reinterpret_cast<uint32_t*>(comb)[NH * 2 * BN + i] =
                        (n0_fixed + i < p.N && __ldg(p.gamma + n0_fixed + i) >= 0.f) ? 0u : 0x80000000u;
                for (int i = et; i < NH * BN * 2; i += EPI_THREADS) reinterpret_cast<uint32_t*>(comb)[NH * 2 * BN + BN + i] = 0u;
            }
            named_bar_sync(1, EPI_THREADS);
        }
        if constexpr (EPI == EPI_BN_RELU || EPI == EPI_BN_RELU_DROP) {
            // {scale, shift} of this CTA's BN columns (n_tile is fixed per CTA): comb[0..BN) = scale, comb[BN..2BN) = shift
            const int n0_fixed = (tile_start % p.num_n_tiles) * BN;
            for (int i = et; i < BN; i += EPI_THREADS) {
                const float4 bp = __ldg(p.bnp + n0_fixed + i);
                comb[i] = bp.x;
                comb[BN + i] = bp.y;
            }
            named_bar_sync(1, EPI_THREADS);
        }

        // EPI_DGRAD_ACT with per-cloud sums: the per-CTA accumulators of sum dz are flushed whenever the CTA's tile sequence
        // (increasing rows) enters another cloud
        int cur_cl = -1;
        auto flush_cloud = [&](int cl) {
            named_bar_sync(1, EPI_THREADS);
            const int n0f = (tile_start % p.num_n_tiles) * BN;
            for (int c = et; c < BN; c += EPI_THREADS) {
                float a0 = 0.f;
#pragma unroll
                for (int h = 0; h < NH; ++h) {
                    a0 += comb[(h * 2) * BN + c];
                    comb[(h * 2) * BN + c] = 0.f;
                }
                atomicAdd(p.cloud_sums + static_cast<size_t>(cl) * p.N + n0f + c, a0);
                atomicAdd(p.stats + n0f + c, static_cast<double>(a0));
            }
            named_bar_sync(1, EPI_THREADS);
        };
        (void)cur_cl;
        (void)flush_cloud;
        // EPI_STATS_POOL: the arg-extremum keys of the CTA's tiles are merged in shared memory (pool_s: [NH][BN] 64-bit keys,
        // 0 = empty) and go to global memory with ONE atomicMax per column whenever the CTA's tile sequence enters another
        // cloud -- instead of one per column and sub-tile, whose outstanding atomics every later proxy fence had to wait for
        const uint32_t pool_s = comb_s + 4u * (NH * 2 * BN + BN);
        auto flush_pool = [&](int cl) {
            named_bar_sync(1, EPI_THREADS);
            const int n0f = (tile_start % p.num_n_tiles) * BN;
            for (int c = et; c < BN; c += EPI_THREADS) {
                unsigned long long best = 0ull;
#pragma unroll
                for (int h = 0; h < NH; ++h) {
                    const unsigned long long k = lds_u64(pool_s + 8u * (h * BN + c));
                    best = k > best ? k : best;
                    sts_u64(pool_s + 8u * (h * BN + c), 0ull);
                }
                if (best != 0ull && n0f + c < p.N) atomicMax(p.pool_keys + static_cast<size_t>(cl) * p.N + n0f + c, best);
            }
            named_bar_sync(1, EPI_THREADS);
        };
        (void)pool_s;
        (void)flush_pool;

        for (int tile = tile_start; tile < total_tiles; tile += tile_step, ++iter) {
            int m_tile, n_tile, split;
            tile_coords(tile, m_tile, n_tile, split);
            const int acc = iter & 1;
            const uint32_t acc_phase = (iter >> 1) & 1;
            const int m0 = m_tile * GEMM_BM;
            const int n0 = n_tile * BN;
            const int grow = m0 + row;
            const bool valid = grow < p.M;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_acc = tmem_base + acc * BN + lane_sel;

            if constexpr (EPI == EPI_WGRAD) {
                // fp32 split-K partials straight from registers: row = output channel, CW consecutive input channels
                float* dst_row = p.out_f32 + static_cast<size_t>(grow) * p.ldc + n0;
                int kb0, kb1;
                kb_range(split, kb0, kb1);
                const bool nonempty = kb0 < kb1;
#pragma unroll 1
                for (int c = cq; c < BN / CW; c += NQ) {
                    uint32_t v[CW];
                    tmem_ld_cols<CW>(t_acc + c * CW, v);
                    tmem_ld_wait();
                    if (p.wg_mode == 1) {
                        // single split: the tile is final -> bf16 (the B operand of a later GEMM)
                        if (valid) {
                            __nv_bfloat16* drow = p.out_bf16 + static_cast<size_t>(grow) * p.ldc + n0 + c * CW;
#pragma unroll
                            for (int j = 0; j < CW / 8; ++j)
                                *reinterpret_cast<uint4*>(drow + j * 8) =
                                    make_uint4(pack_bf16x2(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1])),
                                               pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])),
                                               pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])),
                                               pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])));
                        }
                    } else if (p.wg_mode == 2) {
                        // folded BatchNorm backward: dW = A Q + Bc (W Gc) + D s^T, acc = (W Gc)[row][col]
                        if (valid) {
                            const float4 cf = __ldg(p.wcoef + grow);
                            const float* qrow = p.wq + static_cast<size_t>(grow) * p.ldc + n0 + c * CW;
#pragma unroll
                            for (int j = 0; j < CW / 4; ++j) {
                                const float4 q4 = *reinterpret_cast<const float4*>(qrow + j * 4);
                                const int col = n0 + c * CW + j * 4;
                                float4 o;
                                o.x = fmaf(cf.x, q4.x, fmaf(cf.y, __uint_as_float(v[4 * j]), cf.w * static_cast<float>(p.ws[col])));
                                o.y = fmaf(cf.x, q4.y, fmaf(cf.y, __uint_as_float(v[4 * j + 1]), cf.w * static_cast<float>(p.ws[col + 1])));
                                o.z = fmaf(cf.x, q4.z, fmaf(cf.y, __uint_as_float(v[4 * j + 2]), cf.w * static_cast<float>(p.ws[col + 2])));
                                o.w = fmaf(cf.x, q4.w, fmaf(cf.y, __uint_as_float(v[4 * j + 3]), cf.w * static_cast<float>(p.ws[col + 3])));
                                *reinterpret_cast<float4*>(dst_row + c * CW + j * 4) = o;
                            }
                        }
                    } else if (p.wg_mode == 3) {
                        if (valid) {
                            float* prow = dst_row + static_cast<size_t>(split) * p.M * p.ldc + c * CW;
#pragma unroll
                            for (int j = 0; j < CW / 4; ++j)
                                *reinterpret_cast<float4*>(prow + j * 4) =
                                    nonempty ? make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                           __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]))
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    } else if (valid && nonempty) {
#pragma unroll
                        for (int j = 0; j < CW / 4; ++j) {
                            const int col = n0 + c * CW + j * 4;
                            if (col + 3 < p.N) {
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst_row + c * CW + j * 4),
                                             "f"(__uint_as_float(v[4 * j])), "f"(__uint_as_float(v[4 * j + 1])),
                                             "f"(__uint_as_float(v[4 * j + 2])), "f"(__uint_as_float(v[4 * j + 3]))
                                             : "memory");
                            } else {
                                for (int e = 0; e < 4; ++e)
                                    if (col + e < p.N) atomicAdd(dst_row + c * CW + j * 4 + e, __uint_as_float(v[4 * j + e]));
                            }
                        }
                    }
                }
                tc_fence_before();
                release_acc(acc);
            } else if constexpr (EPI == EPI_LOGITS) {
                static_assert(EPI != EPI_LOGITS || BN == 128, "logits epilogue needs all 128 channels in one tile");
                float lg[MAX_CLASSES];
#pragma unroll
                for (int k = 0; k < MAX_CLASSES; ++k) lg[k] = 0.f;
#pragma unroll 1
                for (int c = cq; c < BN / CW; c += NQ) {
                    uint32_t v[CW];
                    tmem_ld_cols<CW>(t_acc + c * CW, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < CW; ++i) {
                        const int col = c * CW + i;
                        const float a = fmaxf(__uint_as_float(v[i]) + __ldg(p.bias + col), 0.f);
#pragma unroll
                        for (int k = 0; k < MAX_CLASSES; ++k)
                            if (k < p.num_classes) lg[k] = fmaf(a, w4s[k * 128 + col], lg[k]);
                    }
                }
                tc_fence_before();
                release_acc(acc);
                // combine the NQ column groups through shared memory: comb[cq-1][row][class]
                named_bar_sync(1, EPI_THREADS);       // previous tile's readers are done with comb
                if (cq > 0) {
#pragma unroll
                    for (int k = 0; k < MAX_CLASSES; ++k) comb[((cq - 1) * 128 + row) * MAX_CLASSES + k] = lg[k];
                }
                named_bar_sync(1, EPI_THREADS);
                if (cq == 0 && valid) {
                    float* dst = p.logits + static_cast<size_t>(grow) * p.num_classes;
#pragma unroll
                    for (int k = 0; k < MAX_CLASSES; ++k) {
                        if (k < p.num_classes) {
                            float z = lg[k] + w4s[MAX_CLASSES * 128 + k];
#pragma unroll
                            for (int q = 1; q < NQ; ++q) z += comb[((q - 1) * 128 + row) * MAX_CLASSES + k];
                            dst[k] = z;
                        }
                    }
                }
            } else {
                // cloud of this tile: dense batches divide by the rows per cloud, packed ragged batches look it up (tiles
                // never straddle clouds there); everything per tile, nothing of it inside the column loop below
                int tile_cl = 0, row0_in_cloud = 0;
                bool uniform_cloud = true;
                if (p.tile_cloud != nullptr) {
                    tile_cl = __ldg(p.tile_cloud + (m0 >> 7));
                    if constexpr (EPI == EPI_STATS_POOL) row0_in_cloud = m0 - __ldg(p.cloud_off + tile_cl);
                } else if (p.pts_per_cloud > 0) {
                    tile_cl = m0 / p.pts_per_cloud;
                    row0_in_cloud = m0 - tile_cl * p.pts_per_cloud;
                    uniform_cloud = row0_in_cloud + (min(m0 + GEMM_BM, p.M) - m0) <= p.pts_per_cloud;
                }
                int side_slot = 0x7fffffff;
                if constexpr (IS_DGRAD) {
                    if (p.rowslot != nullptr && valid) side_slot = __ldg(p.rowslot + grow);
                }
                const int cloud = uniform_cloud ? tile_cl : (valid ? grow / p.pts_per_cloud : 0);
                const float* cb_row = (p.cloud_bias != nullptr) ? p.cloud_bias + static_cast<size_t>(cloud) * p.N : nullptr;
                const bool pool_uniform = uniform_cloud;
                if constexpr (EPI == EPI_DGRAD_ACT) {
                    if (p.cloud_sums != nullptr && tile_cl != cur_cl) {
                        if (cur_cl >= 0) flush_cloud(cur_cl);
                        cur_cl = tile_cl;
                    }
                }
                if constexpr (EPI == EPI_STATS_POOL) {
                    if (uniform_cloud && tile_cl != cur_cl) {
                        if (cur_cl >= 0) flush_pool(cur_cl);
                        cur_cl = tile_cl;
                    }
                }
                // Software pipelining of the TMEM reads (PREFETCH): the accumulator columns of sub-tile s + 1 are requested as
                // soon as pass 1 of sub-tile s has consumed its registers, so that the tcgen05.ld latency hides behind the
                // barrier and pass 2 instead of heading every sub-tile.
                uint32_t v[CW];
#pragma unroll 1
                for (int sub = 0; sub < SUBS; ++sub, ++sub_it) {
                    const int buf = Cfg::SINGLE_OUT ? 0 : (sub_it & 1);
                    const int c0 = n0 + sub * 64 + cq * CW;       // first global column handled by this thread in pass 1
                    float* comb_b = comb + buf * (4 * 64);        // COLMAX only: [buf][row group][64]
                    if constexpr (Cfg::HAS_Y) mbar_wait(&y_full[buf], (sub_it >> 1) & 1);
                    uint32_t packed[CW / 2];
                    uint32_t packed_lo[(EPI == EPI_BIAS_RELU_X3) ? CW / 2 : 1];
                    (void)packed_lo;
                    if constexpr (EPI == EPI_BIAS_RELU_X3) {       // hi -> staging buffer 0, lo -> buffer 1: both must be free
                        if (elected) tma_store_wait_read<0>();
                        named_bar_sync(3, EPI_THREADS);
                    }
                    if (!PREFETCH || sub == 0) tmem_ld_cols<CW>(t_acc + sub * 64 + cq * CW, v);
                    tmem_ld_wait_regs<CW>(v);
                    if (sub == SUBS - 1) {       // all TMEM reads of this accumulator (by this thread) are done
                        tc_fence_before();
                        release_acc(acc);
                    }
                    if constexpr (EPI == EPI_BIAS_RELU || EPI == EPI_COLMAX || EPI == EPI_BIAS_RELU_X3) {
                        float o[CW];
#pragma unroll
                        for (int i = 0; i < CW; i += 4) {
                            const float4 b4v = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + i));
                            o[i] = __uint_as_float(v[i]) + b4v.x;
                            o[i + 1] = __uint_as_float(v[i + 1]) + b4v.y;
                            o[i + 2] = __uint_as_float(v[i + 2]) + b4v.z;
                            o[i + 3] = __uint_as_float(v[i + 3]) + b4v.w;
                        }
                        if (cb_row != nullptr) {
#pragma unroll
                            for (int i = 0; i < CW; i += 4) {
                                const float4 c4v = __ldg(reinterpret_cast<const float4*>(cb_row + c0 + i));
                                o[i] += c4v.x; o[i + 1] += c4v.y; o[i + 2] += c4v.z; o[i + 3] += c4v.w;
                            }
                        }
                        if constexpr (EPI == EPI_BIAS_RELU) {
                            // rows beyond M are clipped by the TMA store: no masking needed
#pragma unroll
                            for (int i = 0; i < CW / 2; ++i) packed[i] = pack_bf16x2(fmaxf(o[2 * i], 0.f), fmaxf(o[2 * i + 1], 0.f));
                        } else if constexpr (EPI == EPI_BIAS_RELU_X3) {
#pragma unroll
                            for (int i = 0; i < CW / 2; ++i) {
                                const float a0 = fmaxf(o[2 * i], 0.f), a1 = fmaxf(o[2 * i + 1], 0.f);
                                packed[i] = pack_bf16x2(a0, a1);
                                packed_lo[i] = pack_bf16x2(a0 - bf16_lo(packed[i]), a1 - bf16_hi(packed[i]));
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < CW; ++i) o[i] = valid ? fmaxf(o[i], 0.f) : 0.f;
                            if (uniform_cloud) {
                                const float r = warp_colreduce<CW, true>(o);
                                if (CW == 32 || (lane & 1) == 0) comb_b[ew * 64 + cq * CW + (CW == 32 ? lane : (lane >> 1))] = r;
                            } else if (valid) {
#pragma unroll
                                for (int i = 0; i < CW; ++i)
                                    atomicMax(p.colmax + static_cast<size_t>(cloud) * p.N + c0 + i, __float_as_uint(o[i]));
                            }
                        }
                    } else if constexpr (EPI == EPI_BN_RELU) {
                        // rows beyond M are clipped by the TMA store
                        const float* scs = comb + sub * 64 + cq * CW;
#pragma unroll
                        for (int i = 0; i < CW / 2; ++i)
                            packed[i] = pack_bf16x2(fmaxf(fmaf(scs[2 * i], __uint_as_float(v[2 * i]), scs[BN + 2 * i]), 0.f),
                                                    fmaxf(fmaf(scs[2 * i + 1], __uint_as_float(v[2 * i + 1]), scs[BN + 2 * i + 1]), 0.f));
                    } else if constexpr (EPI == EPI_BN_RELU_DROP) {
                        const float* scs = comb + sub * 64 + cq * CW;
                        if (cb_row != nullptr) {       // per-cloud term of seg_conv1 (pcs.py:117-123), added before the normalisation
#pragma unroll
                            for (int i = 0; i < CW; i += 4) {
                                const float4 c4v = __ldg(reinterpret_cast<const float4*>(cb_row + c0 + i));
                                v[i] = __float_as_uint(__uint_as_float(v[i]) + c4v.x);
                                v[i + 1] = __float_as_uint(__uint_as_float(v[i + 1]) + c4v.y);
                                v[i + 2] = __float_as_uint(__uint_as_float(v[i + 2]) + c4v.z);
                                v[i + 3] = __float_as_uint(__uint_as_float(v[i + 3]) + c4v.w);
                            }
                        }
                        if (p.drop_thr16 != 0u) {      // dropout (pcs.py:124): Philox keep mask per 8 consecutive channels of the row
#pragma unroll
                            for (int j = 0; j < CW / 8; ++j) {
                                const unsigned long long e0 = static_cast<unsigned long long>(grow) * p.N + c0 + 8 * j;
                                const uint4 rb = dropout_bits8_fast(seed_eff, e0 >> 3);
                                const uint32_t rs[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
                                for (int e = 0; e < 8; ++e) {
                                    const int i = 8 * j + e;
                                    const uint32_t u16 = (e & 1) ? (rs[e >> 1] >> 16) : (rs[e >> 1] & 0xFFFFu);
                                    const float t = fmaxf(fmaf(scs[i], __uint_as_float(v[i]), scs[BN + i]), 0.f) * p.keep_scale;
                                    v[i] = __float_as_uint(u16 >= p.drop_thr16 ? t : 0.f);
                                }
                            }
#pragma unroll
                            for (int i = 0; i < CW / 2; ++i) packed[i] = pack_bf16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
                        } else {
#pragma unroll
                            for (int i = 0; i < CW / 2; ++i)
                                packed[i] = pack_bf16x2(fmaxf(fmaf(scs[2 * i], __uint_as_float(v[2 * i]), scs[BN + 2 * i]), 0.f),
                                                        fmaxf(fmaf(scs[2 * i + 1], __uint_as_float(v[2 * i + 1]), scs[BN + 2 * i + 1]), 0.f));
                        }
                    } else if constexpr (EPI == EPI_STATS || EPI == EPI_STATS_POOL) {
                        // pass 1 (row-mapped): accumulator (+ per-cloud term) -> bf16 -> staging tile.
                        // Rows beyond M have exactly-zero accumulators (TMA zero fill), so they add nothing to the sums.
                        if (cb_row != nullptr) {
                            if (valid) {
#pragma unroll
                                for (int i = 0; i < CW; i += 4) {
                                    const float4 c4v = __ldg(reinterpret_cast<const float4*>(cb_row + c0 + i));
                                    v[i] = __float_as_uint(__uint_as_float(v[i]) + c4v.x);
                                    v[i + 1] = __float_as_uint(__uint_as_float(v[i + 1]) + c4v.y);
                                    v[i + 2] = __float_as_uint(__uint_as_float(v[i + 2]) + c4v.z);
                                    v[i + 3] = __float_as_uint(__uint_as_float(v[i + 3]) + c4v.w);
                                }
                            }
                        }
#pragma unroll
                        for (int i = 0; i < CW / 2; ++i) packed[i] = pack_bf16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
                    } else if constexpr (IS_DGRAD) {
                        // pass 1 (row-mapped): (dA + constant row) * 1/(1-p) -> bf16 -> staging tile (masking happens
                        // column-mapped in pass 2)
                        if (p.bias != nullptr && valid) {      // (rows beyond M keep their exactly-zero accumulators)
#pragma unroll
                            for (int i = 0; i < CW; i += 4) {
                                const float4 b4v = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + i));
                                v[i] = __float_as_uint(__uint_as_float(v[i]) + b4v.x);
                                v[i + 1] = __float_as_uint(__uint_as_float(v[i + 1]) + b4v.y);
                                v[i + 2] = __float_as_uint(__uint_as_float(v[i + 2]) + b4v.z);
                                v[i + 3] = __float_as_uint(__uint_as_float(v[i + 3]) + b4v.w);
                            }
                        }
                        if (cb_row != nullptr && valid) {    // per-cloud constant row of the folded seg_conv1 backward
#pragma unroll
                            for (int i = 0; i < CW; i += 4) {
                                const float4 c4v = __ldg(reinterpret_cast<const float4*>(cb_row + c0 + i));
                                v[i] = __float_as_uint(__uint_as_float(v[i]) + c4v.x);
                                v[i + 1] = __float_as_uint(__uint_as_float(v[i + 1]) + c4v.y);
                                v[i + 2] = __float_as_uint(__uint_as_float(v[i + 2]) + c4v.z);
                                v[i + 3] = __float_as_uint(__uint_as_float(v[i + 3]) + c4v.w);
                            }
                        }
                        if (side_slot < p.side_rows) {       // rare: this row receives max-pool gradient rows
                            const float* e = p.side + static_cast<size_t>(side_slot) * p.N + c0;
#pragma unroll
                            for (int i = 0; i < CW; i += 4) {
                                const float4 e4 = *reinterpret_cast<const float4*>(e + i);
                                v[i] = __float_as_uint(__uint_as_float(v[i]) + e4.x);
                                v[i + 1] = __float_as_uint(__uint_as_float(v[i + 1]) + e4.y);
                                v[i + 2] = __float_as_uint(__uint_as_float(v[i + 2]) + e4.z);
                                v[i + 3] = __float_as_uint(__uint_as_float(v[i + 3]) + e4.w);
                            }
                        }
                        if (p.drop_thr16 != 0u) {
#pragma unroll
                            for (int i = 0; i < CW / 2; ++i)
                                packed[i] = pack_bf16x2(__uint_as_float(v[2 * i]) * p.keep_scale, __uint_as_float(v[2 * i + 1]) * p.keep_scale);
                        } else {
#pragma unroll
                            for (int i = 0; i < CW / 2; ++i) packed[i] = pack_bf16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
                        }
                    }
                    if constexpr (EPI == EPI_BIAS_RELU_X3) {
                        const uint32_t orow = out_s + row * 128;
#pragma unroll
                        for (int j = 0; j < CW / 8; ++j) {
                            const uint32_t off = ((cq * (CW / 8) + j) ^ (row & 7)) << 4;
                            sts128(orow + off, make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]));
                            sts128(orow + 16384 + off, make_uint4(packed_lo[4 * j], packed_lo[4 * j + 1], packed_lo[4 * j + 2], packed_lo[4 * j + 3]));
                        }
                        fence_proxy_async_smem();
                    } else if constexpr (Cfg::HAS_OUT) {
                        const uint32_t orow = out_s + buf * 16384 + row * 128;
#pragma unroll
                        for (int j = 0; j < CW / 8; ++j)
                            sts128(orow + (((cq * (CW / 8) + j) ^ (row & 7)) << 4),
                                   make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]));
                        if (EPI != EPI_STATS_POOL || p.store_out) fence_proxy_async_smem();   // (generic-proxy readers only otherwise)
                        if constexpr (!Cfg::SINGLE_OUT) {
                            if (elected) tma_store_wait_read<0>();     // stores issued before this iteration have drained
                        }
                    }
                    if constexpr (PREFETCH) {
                        if (sub + 1 < SUBS) tmem_ld_cols<CW>(t_acc + (sub + 1) * 64 + cq * CW, v);
                    }
                    // per-column parameters of pass 2 are fetched BEFORE the barrier so that their latency hides behind it
                    float p2a[8], p2b[8];
                    if constexpr (EPI == EPI_DGRAD) {
                        const int colbase_pre = n0 + sub * 64 + ((warp_idx - 4) & 7) * 8;
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float4 bp = __ldg(p.bnp + colbase_pre + e);
                            p2a[e] = bp.x;
                            p2b[e] = bp.y;
                        }
                    }
                    named_bar_sync(1, EPI_THREADS);
                    if constexpr (Cfg::SINGLE_OUT) {
                        // single staging tile: pass 2 only reads it, so the store can go out now; it must have drained (and
                        // every reader must be done) before pass 1 of the next sub-tile overwrites the tile
                        if (elected && p.store_out) {
                            tma_store_2d(&tmOut, out_stage, n0 + sub * 64, m0);
                            tma_store_commit();
                        }
                    }
                    if constexpr (COLACC) {
                        // pass 2 (column-mapped): warp pw owns 16-byte chunk (pw & 7) = 8 columns of row group (pw >> 3);
                        // lane l handles rows rg*(128/NH) + l + 32 i; per-column parameters live in registers.
                        const int pw = warp_idx - 4;
                        const int chunk = pw & 7;
                        const int rg = pw >> 3;
                        const int colbase = n0 + sub * 64 + chunk * 8;
                        const int rbase = rg * (128 / NH) + lane;
                        const uint32_t tile_s = out_s + buf * 16384;
                        float s1[8], s2[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) s1[e] = s2[e] = 0.f;
                        if constexpr (EPI == EPI_DGRAD_ACT) {
                            dgrad_act_pass2_rows<RPT>(tile_s, y_s + buf * 16384, rbase, chunk, s1, s2);
                        } else if constexpr (EPI == EPI_DGRAD) {
                            const uint32_t ytile = y_s + buf * 16384;
                            if (p.drop_thr16 != 0u)
                                dgrad_pass2_rows<true, RPT>(tile_s, ytile, rbase, chunk, p2a, p2b, s1, s2, seed_eff, p.drop_thr16, m0, p.N, colbase);
                            else
                                dgrad_pass2_rows<false, RPT>(tile_s, ytile, rbase, chunk, p2a, p2b, s1, s2, 0ull, 0u, m0, p.N, colbase);
                        } else if constexpr (EPI == EPI_STATS_POOL) {
                            // statistics as packed fp32 pairs (FADD2 / FFMA2); arg-extremum as ONE signed 32-bit key per
                            // column: (order-preserving image of +-x, whose low 16 bits are free because x is a bf16) |
                            // (0xFFFF - row in tile), so that an integer max keeps the first row among equal values
                            const uint32_t fl_s = comb_s + 4u * (NH * 2 * BN + sub * 64 + chunk * 8);
                            const uint4 f0 = lds128(fl_s), f1 = lds128(fl_s + 16);
                            const uint32_t flip[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
                            unsigned long long s1p[4], s2p[4];
                            int kb[8];
#pragma unroll
                            for (int j = 0; j < 4; ++j) s1p[j] = s2p[j] = 0ull;
#pragma unroll
                            for (int e = 0; e < 8; ++e) kb[e] = static_cast<int>(0x80000000u);
#pragma unroll
                            for (int i = 0; i < RPT; ++i) {
                                const int r = rbase + 32 * i;
                                const uint4 xw = lds128(tile_s + r * 128 + ((chunk ^ (r & 7)) << 4));
                                const uint32_t xs[4] = {xw.x, xw.y, xw.z, xw.w};
                                const bool row_ok = (m0 + r) < p.M;
                                const uint32_t rc = 0xFFFFu - static_cast<uint32_t>(r);
                                uint32_t bits[8];
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    bits[2 * j] = xs[j] << 16;
                                    bits[2 * j + 1] = xs[j] & 0xFFFF0000u;
                                    const unsigned long long xp = pack_f32x2(__uint_as_float(bits[2 * j]), __uint_as_float(bits[2 * j + 1]));
                                    s1p[j] = add_f32x2(s1p[j], xp);
                                    s2p[j] = fma_f32x2(xp, xp, s2p[j]);
                                }
                                if (pool_uniform) {
                                    if (row_ok) {
#pragma unroll
                                        for (int e = 0; e < 8; ++e) {
                                            const uint32_t f = bits[e] ^ flip[e];
                                            const uint32_t k = (f ^ (static_cast<uint32_t>(static_cast<int>(f) >> 31) & 0x7FFF0000u)) | rc;
                                            kb[e] = max(kb[e], static_cast<int>(k));
                                        }
                                    }
                                } else if (row_ok) {
                                    pool_keys_straddle(xw, fl_s, m0 + r, p.pts_per_cloud, p.pool_keys + colbase, p.N);
                                }
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                unpack_f32x2(s1p[j], s1[2 * j], s1[2 * j + 1]);
                                unpack_f32x2(s2p[j], s2[2 * j], s2[2 * j + 1]);
                            }
                            if (pool_uniform) {
                                // 8 keys x 32 lanes -> lanes with (lane & 3) == 0 hold the warp-wide max of one column
#pragma unroll
                                for (int offk = 16, cnt = 4; offk >= 4; offk >>= 1, cnt >>= 1) {
                                    const bool up = (lane & offk) != 0;
#pragma unroll
                                    for (int e = 0; e < cnt; ++e) {
                                        const int send = up ? kb[e] : kb[e + cnt];
                                        const int keepk = up ? kb[e + cnt] : kb[e];
                                        kb[e] = max(keepk, __shfl_xor_sync(0xffffffffu, send, offk));
                                    }
                                }
                                kb[0] = max(kb[0], __shfl_xor_sync(0xffffffffu, kb[0], 2));
                                kb[0] = max(kb[0], __shfl_xor_sync(0xffffffffu, kb[0], 1));
                                if ((lane & 3) == 0 && kb[0] != static_cast<int>(0x80000000u)) {
                                    const int colk = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                                    const uint32_t kk = static_cast<uint32_t>(kb[0]);
                                    const uint32_t sord = kk & 0xFFFF0000u;
                                    const uint32_t rwin = 0xFFFFu - (kk & 0xFFFFu);
                                    // back to the unsigned orderable form k_maxpool_finish decodes (float_orderable)
                                    const uint32_t ord = (sord ^ 0x80000000u) | ((sord & 0x80000000u) ? 0xFFFFu : 0u);
                                    const unsigned long long key = (static_cast<unsigned long long>(ord) << 32) |
                                                                   (0xFFFFFFFFu - static_cast<uint32_t>(row0_in_cloud) - rwin);
                                    // per-CTA running best of (row group, column): this lane is its only writer
                                    const uint32_t ka = pool_s + 8u * (rg * BN + sub * 64 + chunk * 8 + colk);
                                    if (key > lds_u64(ka)) sts_u64(ka, key);
                                }
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < RPT; ++i) {
                                const int r = rbase + 32 * i;
                                const uint4 xw = lds128(tile_s + r * 128 + ((chunk ^ (r & 7)) << 4));
                                const uint32_t xs[4] = {xw.x, xw.y, xw.z, xw.w};
#pragma unroll
                                for (int e = 0; e < 8; ++e) {
                                    const float x = (e & 1) ? bf16_hi(xs[e >> 1]) : bf16_lo(xs[e >> 1]);
                                    s1[e] += x;
                                    s2[e] = fmaf(x, x, s2[e]);
                                }
                            }
                        }
                        // warp reduction of the 16 partial sums: lane (q<<4 | c2<<3 | c1<<2 | c0<<1 | x) ends with quantity q, column c
                        {
                            const bool up16 = (lane & 16) != 0;
                            float t[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const float send = up16 ? s1[e] : s2[e];
                                const float keepv = up16 ? s2[e] : s1[e];
                                t[e] = keepv + __shfl_xor_sync(0xffffffffu, send, 16);
                            }
#pragma unroll
                            for (int offk = 8, cnt = 4; offk >= 2; offk >>= 1, cnt >>= 1) {
                                const bool up = (lane & offk) != 0;
#pragma unroll
                                for (int e = 0; e < cnt; ++e) {
                                    const float send = up ? t[e] : t[e + cnt];
                                    const float keepv = up ? t[e + cnt] : t[e];
                                    t[e] = keepv + __shfl_xor_sync(0xffffffffu, send, offk);
                                }
                            }
                            t[0] += __shfl_xor_sync(0xffffffffu, t[0], 1);
                            if ((lane & 1) == 0) {
                                const int qn = lane >> 4;
                                const int colk = ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                                const uint32_t ca = comb_s + 4u * ((rg * 2 + qn) * BN + sub * 64 + chunk * 8 + colk);
                                sts_f32(ca, lds_f32(ca) + t[0]);      // unique owner: no race
                            }
                        }
                        if constexpr (IS_DGRAD) {
                            fence_proxy_async_smem();
                            named_bar_sync(2, EPI_THREADS);     // staging tile was modified in place
                        }
                        if constexpr (Cfg::SINGLE_OUT) {
                            if (elected && p.store_out) tma_store_wait_read<0>();
                            named_bar_sync(2, EPI_THREADS);
                        }
                    }
                    if constexpr (EPI == EPI_BIAS_RELU_X3) {
                        if (elected) {
                            tma_store_2d(&tmOut, out_stage, n0 + sub * 64, m0);
                            tma_store_2d(&tmOut, out_stage + 16384, p.x3_lo_col + n0 + sub * 64, m0);
                            tma_store_commit();
                        }
                    } else if constexpr (Cfg::HAS_OUT && !Cfg::SINGLE_OUT) {
                        if (elected) {
                            if (EPI != EPI_STATS_POOL || p.store_out) {
                                tma_store_2d(&tmOut, out_stage + buf * 16384, n0 + sub * 64, m0);
                                tma_store_commit();
                            }
                            if constexpr (Cfg::HAS_Y) issue_y_load(sub_it + 2);
                        }
                    }
                    if constexpr (EPI == EPI_COLMAX) {
                        if (uniform_cloud && et < 64) {
                            const float s = fmaxf(fmaxf(comb_b[et], comb_b[64 + et]), fmaxf(comb_b[128 + et], comb_b[192 + et]));
                            const int cl = tile_cl;
                            const int col = n0 + sub * 64 + et;
                            if (col < p.N && s > 0.f) atomicMax(p.colmax + static_cast<size_t>(cl) * p.N + col, __float_as_uint(s));
                        }
                    }
                }
            }
        }
        if constexpr (EPI == EPI_DGRAD_ACT) {
            if (p.cloud_sums != nullptr && cur_cl >= 0) flush_cloud(cur_cl);
        }
        if constexpr (EPI == EPI_STATS_POOL) {
            if (cur_cl >= 0) flush_pool(cur_cl);
        }
        if constexpr (COLACC) {
            named_bar_sync(1, EPI_THREADS);
            const int n0_fixed = (tile_start % p.num_n_tiles) * BN;
            if (tile_start < total_tiles) {
                for (int c = et; c < BN; c += EPI_THREADS) {
                    const int col = n0_fixed + c;
                    if (col >= p.N) continue;
                    float a0 = 0.f, a1 = 0.f;
#pragma unroll
                    for (int h = 0; h < NH; ++h) { a0 += comb[(h * 2) * BN + c]; a1 += comb[(h * 2 + 1) * BN + c]; }
                    if constexpr (EPI == EPI_DGRAD) {
                        const float4 bp = __ldg(p.bnp + col);      // sum dz*yhat = invstd * sum(dz*y) + (-mean*invstd) * sum dz
                        atomicAdd(p.stats + col, static_cast<double>(a0));
                        atomicAdd(p.stats + p.N + col, static_cast<double>(bp.z) * a1 + static_cast<double>(bp.w) * a0);
                    } else {
                        atomicAdd(p.stats + col, static_cast<double>(a0));
                        atomicAdd(p.stats + p.N + col, static_cast<double>(a1));
                    }
                }
            }
        }
        if constexpr (Cfg::HAS_OUT) {
            if (elected) tma_store_wait_all<0>();
        }
    }

    // this CTA's work is done: let the next kernel of the stream start launching (it still waits for the whole grid in
    // its own griddepcontrol.wait).  Triggering at kernel start instead was measured slower in training: the dependent
    // kernel's CTAs then sit resident on the SMs for the whole GEMM.
    pdl_launch_dependents();
    tc_fence_before();
    if (C2) cluster_sync_all();      // neither CTA frees tensor memory (or exits) while its peer may still signal or read it
    else __syncthreads();
    if (warp_idx == 2) {
        tc_fence_after();
        if (C2) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
        else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

}  // namespace pcseg
