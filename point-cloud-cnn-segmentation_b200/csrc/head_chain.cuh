// Inference segmentation head as ONE kernel: seg_conv1 -> seg_conv2 -> seg_conv3 -> seg_conv4 (pcs.py:123-131,
// eval mode: BatchNorm folded, dropout = identity) for a 128-point tile, with the 512- and 256-channel intermediates
// kept on chip:
//
//   a2 tile [128 x 64]  --TMA-->  smem  --tcgen05.mma (N = 2 x 256)-->  TMEM[0..512)
//        epilogue 1: + per-cloud term, ReLU, bf16 -> smem ACT [128 x 512] in the UMMA K-major 128B-swizzled layout
//   ACT [128 x 512]  x  W_s2' [256 x 512] (streamed from L2 by TMA)  -->  TMEM[0..256)
//        epilogue 2: + bias, ReLU, bf16 -> smem ACT [128 x 256]
//   ACT [128 x 256]  x  W_s3' [128 x 256] (streamed)  -->  TMEM[256..384)
//        epilogue 3: + bias, ReLU, seg_conv4 (128 -> C) + bias -> fp32 logits
//
// HBM traffic per point: 128 B in (point_feat) + 4C B out, instead of writing and re-reading the 1 KB + 512 B
// intermediates.  Same warp roles as gemm_kernel: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator,
// warps 4..11 epilogue.
#pragma once
#include "gemm.cuh"

namespace pcseg {

struct HeadChainParams {
    int M;                       // points
    int num_tiles;               // ceil(M / 128)
    int pts_per_cloud;
    const int* tile_cloud;       // ragged (packed) execution: cloud of every 128-row tile, or nullptr
    const float* cloud_bias;     // [clouds][512]  folded seg_conv1 bias + global-feature term
    const float* bias2;          // [256] folded seg_conv2 / bn_seg2
    const float* bias3;          // [128] folded seg_conv3 / bn_seg3
    const float* w4;             // [C][128]
    const float* b4;             // [C]
    int num_classes;
    float* logits;               // [M][C]
};

constexpr int HC_THREADS = 128 + 256;
constexpr int HC_A1_BYTES = 16384;
constexpr int HC_ACT_BYTES = 8 * 16384;
constexpr int HC_STAGES = 2;
constexpr int HC_STAGE_BYTES = 32768;
constexpr int HC_W4_BYTES = (MAX_CLASSES * 128 + MAX_CLASSES) * 4;
constexpr int HC_COMB_BYTES = 128 * MAX_CLASSES * 4;
constexpr int HC_SMEM_BYTES = 1024 + HC_A1_BYTES + HC_ACT_BYTES + HC_STAGES * HC_STAGE_BYTES + HC_W4_BYTES + HC_COMB_BYTES + 256;

__global__ void __launch_bounds__(HC_THREADS, 1)
head_chain_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                  const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmB3,
                  const HeadChainParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a1 = smem;
    uint8_t* act = a1 + HC_A1_BYTES;
    uint8_t* bst = act + HC_ACT_BYTES;
    float* w4s = reinterpret_cast<float*>(bst + HC_STAGES * HC_STAGE_BYTES);
    float* comb = w4s + HC_W4_BYTES / 4;
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(comb) + HC_COMB_BYTES);
    uint64_t* full_bar = bars;                    // [2]  B stream stage filled
    uint64_t* empty_bar = bars + 2;               // [2]  B stream stage consumed
    uint64_t* a1_full = bars + 4;                 //      point_feat tile landed
    uint64_t* a1_empty = bars + 5;                //      seg_conv1 MMAs done reading it
    uint64_t* d_full = bars + 6;                  // [3]  accumulator of layer 1/2/3 complete
    uint64_t* act_ready = bars + 9;               // [2]  epilogue 1/2 finished writing ACT
    uint64_t* tmem_free = bars + 11;              //      epilogue 3 finished reading TMEM
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp_idx = threadIdx.x >> 5;
    const uint32_t lane = lane_id();

    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmA1);
        tma_prefetch_desc(&tmB1);
        tma_prefetch_desc(&tmB2);
        tma_prefetch_desc(&tmB3);
    }
    if (warp_idx == 1 && lane == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
            mbar_init(&act_ready[i], 256);
        }
        mbar_init(a1_full, 1);
        mbar_init(a1_empty, 1);
        for (int i = 0; i < 3; ++i) mbar_init(&d_full[i], 1);
        mbar_init(tmem_free, 256);
        fence_barrier_init();
    }
    if (warp_idx == 2) {
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    }
    pdl_wait();
    for (int i = threadIdx.x; i < p.num_classes * 128; i += HC_THREADS) w4s[i] = p.w4[i];
    for (int i = threadIdx.x; i < p.num_classes; i += HC_THREADS) w4s[MAX_CLASSES * 128 + i] = p.b4[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp_idx == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            auto push = [&](const CUtensorMap* m, int c0, int c1, uint32_t bytes) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                mbar_arrive_expect_tx(&full_bar[stage], bytes);
                tma_load_2d(bst + stage * HC_STAGE_BYTES, m, &full_bar[stage], c0, c1);
                if (++stage == HC_STAGES) { stage = 0; phase ^= 1; }
            };
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
                mbar_wait(a1_empty, (it & 1) ^ 1);
                mbar_arrive_expect_tx(a1_full, HC_A1_BYTES);
                tma_load_2d(a1, &tmA1, a1_full, 0, tile * 128);
                push(&tmB1, 0, 0, 32768);                                    // seg_conv1 rows   0..255
                push(&tmB1, 0, 256, 32768);                                  // seg_conv1 rows 256..511
                for (int kb = 0; kb < 8; ++kb) push(&tmB2, kb * 64, 0, 32768);   // seg_conv2 [256 x 64] k-blocks
                for (int kb = 0; kb < 4; ++kb) push(&tmB3, kb * 64, 0, 16384);   // seg_conv3 [128 x 64] k-blocks
            }
        }
    } else if (warp_idx == 1) {
        // ------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc256 = make_idesc_bf16(128, 256, 0, 0);
        constexpr uint32_t idesc128 = make_idesc_bf16(128, 128, 0, 0);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        const uint32_t a1_s = smem_u32(a1);
        const uint32_t act_s = smem_u32(act);
        const uint32_t bst_s = smem_u32(bst);
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const uint32_t par = it & 1;
            // ---- seg_conv1: D[0..512) = a2 tile x Wpf'^T
            mbar_wait(tmem_free, par ^ 1);
            mbar_wait(a1_full, par);
            tc_fence_after();
            for (int h = 0; h < 2; ++h) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (lane == 0) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(tmem_base + h * 256, make_smem_desc_sw128(a1_s + k * 32, 0, 1024),
                                  make_smem_desc_sw128(bst_s + stage * HC_STAGE_BYTES + k * 32, 0, 1024), idesc256, k > 0 ? 1u : 0u);
                    umma_commit(&empty_bar[stage]);
                    if (h == 1) { umma_commit(a1_empty); umma_commit(&d_full[0]); }
                }
                __syncwarp();
                if (++stage == HC_STAGES) { stage = 0; phase ^= 1; }
            }
            // ---- seg_conv2: D[0..256) = ACT[128 x 512] x Ws2'^T
            mbar_wait(&act_ready[0], par);
            tc_fence_after();
            for (int kb = 0; kb < 8; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (lane == 0) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(tmem_base, make_smem_desc_sw128(act_s + kb * 16384 + k * 32, 0, 1024),
                                  make_smem_desc_sw128(bst_s + stage * HC_STAGE_BYTES + k * 32, 0, 1024), idesc256,
                                  (kb > 0 || k > 0) ? 1u : 0u);
                    umma_commit(&empty_bar[stage]);
                    if (kb == 7) umma_commit(&d_full[1]);
                }
                __syncwarp();
                if (++stage == HC_STAGES) { stage = 0; phase ^= 1; }
            }
            // ---- seg_conv3: D[256..384) = ACT[128 x 256] x Ws3'^T
            mbar_wait(&act_ready[1], par);
            tc_fence_after();
            for (int kb = 0; kb < 4; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (lane == 0) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(tmem_base + 256, make_smem_desc_sw128(act_s + kb * 16384 + k * 32, 0, 1024),
                                  make_smem_desc_sw128(bst_s + stage * HC_STAGE_BYTES + k * 32, 0, 1024), idesc128,
                                  (kb > 0 || k > 0) ? 1u : 0u);
                    umma_commit(&empty_bar[stage]);
                    if (kb == 3) umma_commit(&d_full[2]);
                }
                __syncwarp();
                if (++stage == HC_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp_idx >= 4) {
        // ------------------------------------------------------------ epilogue (8 warps)
        const int ew = warp_idx & 3;                  // TMEM lane quadrant
        const int cq = (warp_idx - 4) >> 2;           // 0 / 1: left / right 32 columns of every 64-column sub-tile
        const int row = ew * 32 + lane;
        const uint32_t lane_sel = static_cast<uint32_t>(ew * 32) << 16;
        const uint32_t act_s = smem_u32(act);
        int it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const uint32_t par = it & 1;
            const int grow = tile * 128 + row;
            const bool valid = grow < p.M;
            const int cloud = p.tile_cloud ? __ldg(p.tile_cloud + tile) : (valid ? grow / p.pts_per_cloud : 0);
            const float* cb_row = p.cloud_bias + static_cast<size_t>(cloud) * 512;
            // ---- epilogue 1: relu(D1 + cb) -> ACT [128 x 512]
            mbar_wait(&d_full[0], par);
            tc_fence_after();
#pragma unroll 1
            for (int sub = 0; sub < 8; ++sub) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + lane_sel + sub * 64 + cq * 32, v);
                tmem_ld_wait();
                const int c0 = sub * 64 + cq * 32;
                uint32_t packed[16];
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 c4 = __ldg(reinterpret_cast<const float4*>(cb_row + c0 + i));
                    packed[i / 2] = pack_bf16x2(fmaxf(__uint_as_float(v[i]) + c4.x, 0.f), fmaxf(__uint_as_float(v[i + 1]) + c4.y, 0.f));
                    packed[i / 2 + 1] = pack_bf16x2(fmaxf(__uint_as_float(v[i + 2]) + c4.z, 0.f), fmaxf(__uint_as_float(v[i + 3]) + c4.w, 0.f));
                }
                const uint32_t orow = act_s + sub * 16384 + row * 128;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    sts128(orow + (((cq * 4 + j) ^ (row & 7)) << 4), make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]));
            }
            tc_fence_before();
            fence_proxy_async_smem();
            mbar_arrive(&act_ready[0]);
            // ---- epilogue 2: relu(D2 + bias2) -> ACT [128 x 256]
            mbar_wait(&d_full[1], par);
            tc_fence_after();
#pragma unroll 1
            for (int sub = 0; sub < 4; ++sub) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + lane_sel + sub * 64 + cq * 32, v);
                tmem_ld_wait();
                const int c0 = sub * 64 + cq * 32;
                uint32_t packed[16];
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 c4 = __ldg(reinterpret_cast<const float4*>(p.bias2 + c0 + i));
                    packed[i / 2] = pack_bf16x2(fmaxf(__uint_as_float(v[i]) + c4.x, 0.f), fmaxf(__uint_as_float(v[i + 1]) + c4.y, 0.f));
                    packed[i / 2 + 1] = pack_bf16x2(fmaxf(__uint_as_float(v[i + 2]) + c4.z, 0.f), fmaxf(__uint_as_float(v[i + 3]) + c4.w, 0.f));
                }
                const uint32_t orow = act_s + sub * 16384 + row * 128;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    sts128(orow + (((cq * 4 + j) ^ (row & 7)) << 4), make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]));
            }
            tc_fence_before();
            fence_proxy_async_smem();
            mbar_arrive(&act_ready[1]);
            // ---- epilogue 3: logits = W4 relu(D3 + bias3) + b4
            mbar_wait(&d_full[2], par);
            tc_fence_after();
            float lg[MAX_CLASSES];
#pragma unroll
            for (int k = 0; k < MAX_CLASSES; ++k) lg[k] = 0.f;
#pragma unroll 1
            for (int c = cq; c < 4; c += 2) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + lane_sel + 256 + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int col = c * 32 + i;
                    const float a = fmaxf(__uint_as_float(v[i]) + __ldg(p.bias3 + col), 0.f);
#pragma unroll
                    for (int k = 0; k < MAX_CLASSES; ++k)
                        if (k < p.num_classes) lg[k] = fmaf(a, w4s[k * 128 + col], lg[k]);
                }
            }
            tc_fence_before();
            mbar_arrive(tmem_free);
            named_bar_sync(1, 256);                   // previous tile's readers are done with comb
            if (cq == 1) {
#pragma unroll
                for (int k = 0; k < MAX_CLASSES; ++k) comb[row * MAX_CLASSES + k] = lg[k];
            }
            named_bar_sync(1, 256);
            if (cq == 0 && valid) {
                float* dst = p.logits + static_cast<size_t>(grow) * p.num_classes;
#pragma unroll
                for (int k = 0; k < MAX_CLASSES; ++k)
                    if (k < p.num_classes) dst[k] = lg[k] + comb[row * MAX_CLASSES + k] + w4s[MAX_CLASSES * 128 + k];
            }
        }
    }

    pdl_launch_dependents();
    tc_fence_before();
    __syncthreads();
    if (warp_idx == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace pcseg
