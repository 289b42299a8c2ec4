// Gradient all-reduce over NVLink peer memory (one 8 x B200 box, one process per GPU): replaces DataParallel's reduce-add of
// replica gradients (pcs.py:209-211, through loss.backward() at pcs.py:254) and the NCCL all-reduce of round 1.
//
// Every rank maps the gradient arenas and signal blocks of all ranks (CUDA IPC).  One kernel per step and rank, launched on
// the compute stream (capturable in the step's CUDA graph: no host involvement, no NCCL launch / stream hand-over latency):
//   barrier A   every rank's backward has finished (its arena is complete)
//   phase 1     two-shot reduce-scatter: rank r sums slice r of ALL arenas (peer loads over NVLink) into its own arena AND
//               into its publish buffer
//   barrier B
//   phase 2     all-gather: rank r copies the reduced slices p != r from their owners' publish buffers
// No third barrier: a publish buffer is rewritten only in the next step's phase 1, i.e. after the next barrier A, which a
// rank reaches only after it has finished this kernel (its phase-2 reads included); the arenas themselves are read by
// peers in phase 1 only, and barrier B separates that from anything that follows.
// The {loss numerator, sum of class weights} pair of the deferred loss normalisation travels in the signal block and is summed
// after barrier A.  Barriers are flag exchanges in peer memory: rank r writes its epoch counter into slot r of every peer's
// signal block (st.release.sys) and spins until all slots of its own block carry that epoch (ld.acquire.sys).
// NVSwitch gives every GPU full bandwidth to every peer, so the slices are read from all peers concurrently.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pcseg {

constexpr int AR_MAX_RANKS = 8;
constexpr int AR_SLOT_U32 = 32;          // one 128-byte line per (rank, barrier) slot

struct PeerSignals {                     // lives in every rank's IPC-shared signal block
    uint32_t flag[3][AR_MAX_RANKS][AR_SLOT_U32];     // [barrier][writer rank][padding]
    double lw[AR_MAX_RANKS][2];                      // written by the owner only: its {loss numerator, sum w}
};

struct PeerArArgs {
    float* arena[AR_MAX_RANKS];          // gradient arenas of all ranks (arena[rank] is local)
    float* pub[AR_MAX_RANKS];            // publish buffers (one slice each) of all ranks
    PeerSignals* sig[AR_MAX_RANKS];      // signal blocks of all ranks
    long long n;                         // floats in the arena
    int rank, world;
    uint32_t* epoch;                     // local device counter, incremented by the kernel (graph replays see fresh epochs)
    const double* lw_in;                 // local {loss numerator, sum w} (nullable)
    double* lw_out;                      // global sums (nullable)
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_cg_f4(const float4* p) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

// Grid-wide + cross-rank barrier.  Block 0 does the flag exchange; the other blocks wait on a local generation counter.
// `local` = {arrive counter, generation} in local device memory.
__device__ __forceinline__ void peer_barrier(const PeerArArgs& a, int which, uint32_t epoch, uint32_t* local) {
    __syncthreads();
    if (threadIdx.x == 0) {
        // one system-scope fence per block: fences are cumulative, the block's writes were ordered before it by bar.sync
        __threadfence_system();
        const uint32_t target = (epoch - 1) * 2 + which + 1; // barriers passed so far (the grid size never changes)
        const uint32_t arrived = atomicAdd(&local[0], 1u) + 1;
        if (arrived == gridDim.x * target) {                 // last block of this rank to arrive: exchange flags with the peers
            for (int p = 0; p < a.world; ++p) st_release_sys(&a.sig[p]->flag[which][a.rank][0], epoch);
            for (int p = 0; p < a.world; ++p) {
                const uint32_t* f = &a.sig[a.rank]->flag[which][p][0];
                const long long t0 = clock64();
                while (ld_acquire_sys(f) != epoch) {
                    if (clock64() - t0 > 20000000000LL) {    // ~10 s of SM clocks: a peer died or never launched; abort instead of hanging
                        printf("pcseg peer all-reduce: rank %d timed out waiting for rank %d (barrier %d, epoch %u)\n", a.rank, p, which, epoch);
                        __trap();
                    }
                }
            }
            __threadfence_system();
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(&local[1]), "r"(target) : "memory");
        } else {
            uint32_t g;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(g) : "l"(&local[1]) : "memory");
            } while (g < target);
        }
        __threadfence_system();
    }
    __syncthreads();
}

template <int WORLD>
__global__ void __launch_bounds__(512) k_peer_allreduce(const PeerArArgs a, uint32_t* local /* {arrive counter, generation} */) {
    // epoch of this launch: read once (every block reads the same value: the increment happens after barrier C)
    const uint32_t epoch = a.epoch[0] + 1;
    if (a.lw_in != nullptr && blockIdx.x == 0 && threadIdx.x < 2) a.sig[a.rank]->lw[a.rank][threadIdx.x] = a.lw_in[threadIdx.x];
    peer_barrier(a, 0, epoch, local);

    if (a.lw_out != nullptr && blockIdx.x == 0 && threadIdx.x < 2) {
        double s = 0.0;
        for (int p = 0; p < a.world; ++p) s += *reinterpret_cast<volatile double*>(&a.sig[p]->lw[p][threadIdx.x]);
        a.lw_out[threadIdx.x] = s;
    }
    // slices in units of float4; slice r = [r * per, min((r + 1) * per, n4))
    const long long n4 = a.n / 4;                            // (the arena length is padded to a multiple of 4 floats by the host)
    const long long per = (n4 + a.world - 1) / a.world;
    const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    {   // phase 1: reduce my slice
        const long long lo = a.rank * per, hi = min(lo + per, n4);
        float4* mine = reinterpret_cast<float4*>(a.arena[a.rank]);
        float4* pub = reinterpret_cast<float4*>(a.pub[a.rank]) - lo;
        const float4* peer[WORLD];
#pragma unroll
        for (int q = 1; q < WORLD; ++q) peer[q] = reinterpret_cast<const float4*>(a.arena[(a.rank + q) % WORLD]);   // staggered start
        constexpr int U = (WORLD <= 2) ? 4 : (WORLD <= 4 ? 2 : 1);          // elements per thread in flight: U * (WORLD - 1) peer loads
        for (long long i0 = lo + tid; i0 < hi; i0 += stride * U) {
            float4 v[U][WORLD];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long i = i0 + u * stride;
                if (i < hi) {
                    v[u][0] = mine[i];
#pragma unroll
                    for (int q = 1; q < WORLD; ++q) v[u][q] = ld_cg_f4(peer[q] + i);      // all peer loads in flight before the first add
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long i = i0 + u * stride;
                if (i < hi) {
#pragma unroll
                    for (int q = 1; q < WORLD; ++q) { v[u][0].x += v[u][q].x; v[u][0].y += v[u][q].y; v[u][0].z += v[u][q].z; v[u][0].w += v[u][q].w; }
                    mine[i] = v[u][0];
                    pub[i] = v[u][0];
                }
            }
        }
    }
    peer_barrier(a, 1, epoch, local);
    {   // phase 2: gather the other slices
        float4* mine = reinterpret_cast<float4*>(a.arena[a.rank]);
#pragma unroll
        for (int q = 1; q < WORLD; ++q) {
            const int p = (a.rank + q) % WORLD;
            const long long lo = p * per, hi = min(lo + per, n4);
            const float4* src = reinterpret_cast<const float4*>(a.pub[p]) - lo;
            for (long long i0 = lo + tid; i0 < hi; i0 += stride * 4) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (i0 + u * stride < hi) v[u] = ld_cg_f4(src + i0 + u * stride);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (i0 + u * stride < hi) mine[i0 + u * stride] = v[u];
            }
        }
    }
    // the epoch counter may only advance once every block has read it: count the blocks that are done
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(&local[3], 1u) + 1 == gridDim.x * epoch) a.epoch[0] = epoch;
    }
}

}  // namespace pcseg
