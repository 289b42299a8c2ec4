// C-ABI implementation: model layout, workspace planning, TMA descriptors and the launch
// sequences for eval forward, train forward and backward.  See include/pcseg_b200.h.
#include "../../include/pcseg_b200.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "gemm.cuh"
#include "head_chain.cuh"
#include "peer_allreduce.cuh"
#include "pointwise.cuh"

using namespace pcseg;

// MAX_CLASSES (8, gemm.cuh) is the cap of the fused head kernels (head_chain_kernel, EPI_LOGITS, k_head_bwd); models with
// 9 .. PCSEG_MAX_CLASSES classes run seg_conv3 / seg_conv4 through the wide kernels (k_head_fwd<16|32>, k_head_bwd_wide_*).
constexpr int API_MAX_CLASSES = PCSEG_MAX_CLASSES;
static_assert(API_MAX_CLASSES == 32 && MAX_CLASSES == 8, "class caps: fused kernels 8, wide kernels 32");
static_assert(sizeof(pcseg_ce_accum) == sizeof(CeAccum), "CE accumulator layout mismatch");
static_assert(sizeof(pcseg_step_state) == sizeof(StepState) && sizeof(StepState) == 32, "step state layout mismatch");

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static long long g_launches = 0;

static int fail(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}
#define CUDA_OK(expr)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)
#define LAUNCH_OK(name)                                                                        \
    do {                                                                                       \
        ++g_launches;                                                                          \
        cudaError_t e__ = cudaGetLastError();                                                  \
        if (e__ != cudaSuccess) return fail("launch of %s failed: %s", name, cudaGetErrorString(e__)); \
    } while (0)
#define TRY(expr)                   \
    do {                            \
        int r__ = (expr);           \
        if (r__ != 0) return r__;   \
    } while (0)

extern "C" const char* pcseg_last_error(void) { return g_err.c_str(); }
extern "C" const char* pcseg_version(void) { return "pcseg_b200 0.1 (sm_100a, tcgen05/TMA)"; }
extern "C" long long pcseg_launch_count(void) { return g_launches; }


// ------------------------------------------------------------------------------------------------
// kernel launch with optional programmatic dependent launch (PDL): the next kernel of the stream may start its prologue
// while this one drains; every kernel calls griddepcontrol.wait before touching data (ptx.cuh).
// Default per entry point (g_pdl_call, set by the API functions): ON for every eagerly launched call (measured: -3 % ..
// -9 % for inference and for training steps of <= 64k points, neutral for larger ones), OFF while the stream is being
// captured: the graph-replayed training step was measured 4-5 % slower with the attribute.  PCSEG_PDL=0 / 1 forces it.
// ------------------------------------------------------------------------------------------------
static thread_local bool g_pdl_call = false;
static bool stream_is_capturing(void* stream) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(static_cast<cudaStream_t>(stream), &st) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return st != cudaStreamCaptureStatusNone;
}
static bool pdl_enabled() {
    static int v = -2;
    if (v == -2) {
        const char* e = getenv("PCSEG_PDL");
        v = (e && e[0] == '1') ? 1 : ((e && e[0] == '0') ? 0 : -1);
    }
    return v >= 0 ? v == 1 : g_pdl_call;
}
template <typename... KArgs, typename... Args>
static void pdl_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);     // errors surface through cudaGetLastError() in LAUNCH_OK
}
// the same for kernels that run as clusters of two CTAs (cta_group::2 GEMMs)
template <typename... KArgs, typename... Args>
static void pdl_launch_pair(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = 2;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ------------------------------------------------------------------------------------------------
// model layout (reference pcs.py:70-94)
// ------------------------------------------------------------------------------------------------
namespace {

constexpr int NUM_CONV = 10;
constexpr int NUM_BN = 9;
constexpr float BN_EPS = 1e-5f;
constexpr float BN_MOMENTUM = 0.1f;

struct ConvDef { int cin, cout; };
inline void conv_defs(int C, ConvDef* d) {
    const ConvDef base[NUM_CONV] = {{4, 64}, {64, 64}, {64, 64}, {64, 128}, {128, 1024}, {1024, 1024},
                                    {1088, 512}, {512, 256}, {256, 128}, {128, C}};
    for (int i = 0; i < NUM_CONV; ++i) d[i] = base[i];
}
// parameter tensors in state_dict order: conv i -> (2i weight, 2i+1 bias); bn j -> (20+2j weight, 21+2j bias)
struct Layout {
    long long off[PCSEG_NUM_PARAM_TENSORS];
    long long numel[PCSEG_NUM_PARAM_TENSORS];
    long long total;
    long long bn_off[NUM_BN][2];
    long long bn_total;
    ConvDef conv[NUM_CONV];
};
inline Layout make_layout(int C) {
    Layout L;
    conv_defs(C, L.conv);
    long long o = 0;
    for (int i = 0; i < NUM_CONV; ++i) {
        L.off[2 * i] = o; L.numel[2 * i] = 1LL * L.conv[i].cin * L.conv[i].cout; o += L.numel[2 * i];
        L.off[2 * i + 1] = o; L.numel[2 * i + 1] = L.conv[i].cout; o += L.conv[i].cout;
    }
    long long b = 0;
    for (int j = 0; j < NUM_BN; ++j) {
        const int c = L.conv[j].cout;
        L.off[20 + 2 * j] = o; L.numel[20 + 2 * j] = c; o += c;
        L.off[21 + 2 * j] = o; L.numel[21 + 2 * j] = c; o += c;
        L.bn_off[j][0] = b; b += c;
        L.bn_off[j][1] = b; b += c;
    }
    L.total = o;
    L.bn_total = b;
    return L;
}

// ------------------------------------------------------------------------------------------------
// TMA descriptors
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

int load_encode() {
    if (g_encode) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || fn == nullptr || q != cudaDriverEntryPointSuccess)
        return fail("cuTensorMapEncodeTiled unavailable (%s): a CUDA 12 driver and a GPU are required", cudaGetErrorString(e));
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    return 0;
}

// 2-D bf16 tensor [outer][inner] with row pitch ld (elements); box = box_inner x box_outer; 128B swizzle.
int make_tmap(CUtensorMap* m, const void* ptr, long long inner, long long outer, long long ld, int box_inner, int box_outer) {
    TRY(load_encode());
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld * 2) % 16 != 0)
        return fail("TMA operand misaligned (ptr %p, pitch %lld elements)", ptr, ld);
    if (box_inner * 2 != 128) return fail("128B swizzle needs a 64-element inner box");
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_outer)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with %d (inner %lld outer %lld ld %lld)", (int)r, inner, outer, ld);
    return 0;
}

int g_num_sms[64] = {};
int g_sm_limit = 0;          // pcseg_set_sm_limit: persistent grids leave SMs free for concurrent collectives
int num_sms() {
    int dev = 0;
    cudaGetDevice(&dev);
    int& n = g_num_sms[dev & 63];
    if (n == 0) {
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return (g_sm_limit > 0 && g_sm_limit < n) ? g_sm_limit : n;
}

// ------------------------------------------------------------------------------------------------
// GEMM launcher
// ------------------------------------------------------------------------------------------------
struct GemmOp {
    CUtensorMap tmA, tmA2, tmB, tmOut, tmY;
    GemmParams p;
    int bn, epi;
    bool mn;
    bool xf = false;       // transform-stage variant (tmA = pre-BN input, tmA2 = its activation tensor, written by the kernel)
    bool c2 = false;       // CTA-pair variant (cta_group::2): tmB boxes hold half a tile (bn / 2 rows)
    bool ready = false;
};

// CTA-pair GEMM: clusters of two CTAs, every cluster keeps one n_tile (the cluster count is a multiple of num_n_tiles)
template <int EPI>
int launch_gemm_pair_t(const GemmOp& op, cudaStream_t s) {
    using Cfg = GemmCfg<256, EPI, false, false, true>;
    static unsigned long long attr_set_mask = 0;
    auto kern = gemm_kernel<256, EPI, false, false, true>;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(attr_set_mask & (1ull << (dev & 63)))) {
        CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_set_mask |= 1ull << (dev & 63);
    }
    const int tiles = (op.p.num_m_tiles / 2) * op.p.num_n_tiles;
    int clusters = num_sms() / 2;
    if (clusters > tiles) clusters = tiles;
    clusters = (clusters / op.p.num_n_tiles) * op.p.num_n_tiles;
    if (clusters < 1) return fail("internal: CTA-pair GEMM with %d tiles", tiles);
    pdl_launch_pair(kern, 2 * clusters, Cfg::THREADS, Cfg::SMEM_BYTES, s, op.tmA, op.tmA2, op.tmB, op.tmOut, op.tmY, op.p);
    LAUNCH_OK("gemm_kernel<pair>");
    return 0;
}

template <int BN, int EPI, bool MN, bool XF = false>
int launch_gemm_t(const GemmOp& op, cudaStream_t s) {
    using Cfg = GemmCfg<BN, EPI, MN, XF>;
    static unsigned long long attr_set_mask = 0;       // per device: the attribute is a per-device function property
    auto kern = gemm_kernel<BN, EPI, MN, XF>;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(attr_set_mask & (1ull << (dev & 63)))) {
        CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_set_mask |= 1ull << (dev & 63);
    }
    const int tiles = op.p.num_m_tiles * op.p.num_n_tiles * op.p.num_splits;
    int grid = tiles < num_sms() ? tiles : num_sms();
    if (!MN) grid = (grid / op.p.num_n_tiles) * op.p.num_n_tiles;   // every CTA keeps one n_tile (per-CTA column accumulators)
    pdl_launch(kern, grid, Cfg::THREADS, Cfg::SMEM_BYTES, s, op.tmA, op.tmA2, op.tmB, op.tmOut, op.tmY, op.p);
    LAUNCH_OK("gemm_kernel");
    return 0;
}

int launch_gemm(const GemmOp& op, cudaStream_t s) {
    if (!op.ready) return fail("internal: GEMM op not initialised");
    if (op.c2) {
        if (op.bn == 256 && op.epi == EPI_STATS_POOL && !op.mn) return launch_gemm_pair_t<EPI_STATS_POOL>(op, s);
        if (op.bn == 256 && op.epi == EPI_DGRAD_ACT && !op.mn) return launch_gemm_pair_t<EPI_DGRAD_ACT>(op, s);
        if (op.bn == 256 && op.epi == EPI_COLMAX && !op.mn) return launch_gemm_pair_t<EPI_COLMAX>(op, s);
        return fail("internal: no CTA-pair GEMM instantiation for BN=%d EPI=%d", op.bn, op.epi);
    }
    if (op.xf) {
        if (op.bn == 64 && op.epi == EPI_STATS && !op.mn) return launch_gemm_t<64, EPI_STATS, false, true>(op, s);
        if (op.bn == 128 && op.epi == EPI_STATS && !op.mn) return launch_gemm_t<128, EPI_STATS, false, true>(op, s);
        return fail("internal: no transform-stage GEMM instantiation for BN=%d EPI=%d", op.bn, op.epi);
    }
#define CASE(BN_, EPI_, MN_) \
    if (op.bn == BN_ && op.epi == EPI_ && op.mn == MN_) return launch_gemm_t<BN_, EPI_, MN_>(op, s);
    CASE(64, EPI_BIAS_RELU, false) CASE(128, EPI_BIAS_RELU, false) CASE(256, EPI_BIAS_RELU, false)
    CASE(256, EPI_COLMAX, false)
    CASE(128, EPI_LOGITS, false)
    CASE(64, EPI_STATS, false) CASE(128, EPI_STATS, false) CASE(256, EPI_STATS, false)
    CASE(256, EPI_STATS_POOL, false)
    CASE(64, EPI_DGRAD, false) CASE(128, EPI_DGRAD, false) CASE(256, EPI_DGRAD, false)
    CASE(256, EPI_BN_RELU, false) CASE(256, EPI_DGRAD_ACT, false) CASE(256, EPI_BN_RELU_DROP, false)
    CASE(64, EPI_BIAS_RELU_X3, false) CASE(128, EPI_BIAS_RELU_X3, false) CASE(256, EPI_BIAS_RELU_X3, false)
    CASE(64, EPI_WGRAD, true) CASE(128, EPI_WGRAD, true) CASE(256, EPI_WGRAD, true)
#undef CASE
    return fail("internal: no GEMM instantiation for BN=%d EPI=%d MN=%d", op.bn, op.epi, (int)op.mn);
}

inline int pick_bn(int n) { return n >= 256 ? 256 : (n >= 128 ? 128 : 64); }

// Row-major GEMM  D[M,N] = A[M,K] * B[N,K]^T  with a fused epilogue.
int setup_gemm_kmajor(GemmOp* op, int epi, const void* A, int lda, const void* B, int ldb, long long M, int N, int K,
                      void* out, int ldo, const void* ymask, int ldy) {
    if (K % 64 != 0) return fail("GEMM K=%d must be a multiple of 64", K);
    if (N % 64 != 0) return fail("GEMM N=%d must be a multiple of 64", N);
    memset(&op->p, 0, sizeof(op->p));
    op->bn = (epi == EPI_LOGITS) ? 128 : pick_bn(N);
    if (epi == EPI_COLMAX || epi == EPI_STATS_POOL) op->bn = 256;
    if (N % op->bn != 0) return fail("GEMM N=%d not a multiple of the tile width %d", N, op->bn);
    op->epi = epi;
    op->mn = false;
    TRY(make_tmap(&op->tmA, A, K, M, lda, 64, 128));
    op->tmA2 = op->tmA;
    TRY(make_tmap(&op->tmB, B, K, N, ldb, 64, op->bn));
    if (out) TRY(make_tmap(&op->tmOut, out, N, M, ldo, 64, 128)); else op->tmOut = op->tmA;
    if (ymask) TRY(make_tmap(&op->tmY, ymask, N, M, ldy, 64, 128)); else op->tmY = op->tmA;
    op->p.M = static_cast<int>(M);
    op->p.N = N;
    op->p.K = K;
    op->p.num_m_tiles = static_cast<int>((M + 127) / 128);
    op->p.num_n_tiles = N / op->bn;
    op->p.num_splits = 1;
    op->p.kb_per_split = K / 64;
    op->p.keep_scale = 1.f;
    op->p.store_out = 1;
    op->ready = true;
    return 0;
}

// CTA-pair form of a K-major BN = 256 op (cta_group::2): every CTA of a pair loads half of the B tile.  Needs whole pairs
// of 128-row tiles.  PCSEG_PAIR=1 selects it for the three K = 1024 GEMMs (global_feat forward / data gradient / inference
// max-pool).  OFF by default -- measured on cfg2 (tools/gpu_r2_prof.sh, MMA-warp wait accounting): the pair kernels are
// bit-identical but not faster.  With 6 x 32 KB stages instead of 4 x 48 KB the MMA warp still waits 18 % of the time for
// operands and issues one 128 x 256 x 16 MMA per SM every ~175 cycles either way (the rate cuBLAS reaches at this part's
// clocks), while 72 clusters x 2 leave 4 SMs idle and round 27.7 tiles per CTA up to 29: 256 vs 242 us forward, 290 vs 259 us
// data gradient (whose epilogue, not its operands, is what the MMA warp waits for in the pair form).
int make_pair_op(GemmOp* op, const void* B, int ldb) {
    static int enabled = -1;
    if (enabled < 0) {
        const char* e = getenv("PCSEG_PAIR");
        enabled = (e && e[0] == '1');
    }
    if (!enabled || op->bn != 256 || op->mn || op->p.kb_switch != 0 || (op->p.num_m_tiles & 1) || (op->p.M % 256) != 0) return 0;
    TRY(make_tmap(&op->tmB, B, op->p.K, op->p.N, ldb, 64, 128));
    op->c2 = true;
    return 0;
}

// Same with the A operand given as the K-concatenation [A1 (K1 columns) | A2 (K2 columns)] of two tensors.
int setup_gemm_kmajor_cat(GemmOp* op, int epi, const void* A1, int lda1, int K1, const void* A2, int lda2, int K2, const void* B,
                          int ldb, long long M, int N, void* out, int ldo, const void* ymask, int ldy) {
    if (K1 % 64 != 0 || K2 % 64 != 0) return fail("GEMM K1=%d / K2=%d must be multiples of 64", K1, K2);
    TRY(setup_gemm_kmajor(op, epi, A1, lda1, B, ldb, M, N, K1, out, ldo, ymask, ldy));
    TRY(make_tmap(&op->tmA2, A2, K2, M, lda2, 64, 128));
    TRY(make_tmap(&op->tmB, B, K1 + K2, N, ldb, 64, op->bn));
    op->p.K = K1 + K2;
    op->p.kb_per_split = (K1 + K2) / 64;
    op->p.kb_switch = K1 / 64;
    return 0;
}

// Split-bf16 GEMM: A [M][2K] and B [N][2K] hold [hi | lo] column halves; out (optional) [M][2N] likewise.
int setup_gemm_x3(GemmOp* op, int epi, const void* A, const void* B, long long M, int N, int K, void* out) {
    TRY(setup_gemm_kmajor(op, epi, A, 2 * K, B, 2 * K, M, N, 2 * K, nullptr, 0, nullptr, 0));
    if (out) TRY(make_tmap(&op->tmOut, out, 2 * N, M, 2 * N, 64, 128));
    op->p.K = 3 * K;
    op->p.kb_per_split = 3 * K / 64;
    op->p.x3_kb = K / 64;
    op->p.x3_lo_col = N;
    return 0;
}

// Weight-gradient GEMM  D[Mc,Nc] += A[P,Mc]^T * B[P,Nc]  (both operands point-major, K = points), fp32 atomics.
int setup_gemm_wgrad(GemmOp* op, const void* A, int lda, int Mc, const void* B, int ldb, int Nc, long long P, float* out, int ldc) {
    memset(&op->p, 0, sizeof(op->p));
    if (Nc % 64 != 0) return fail("wgrad N=%d must be a multiple of 64", Nc);
    op->bn = pick_bn(Nc);
    if (Nc % op->bn != 0) return fail("wgrad N=%d not a multiple of %d", Nc, op->bn);
    op->epi = EPI_WGRAD;
    op->mn = true;
    TRY(make_tmap(&op->tmA, A, Mc, P, lda, 64, 64));
    TRY(make_tmap(&op->tmB, B, Nc, P, ldb, 64, 64));
    op->tmA2 = op->tmA;
    op->tmOut = op->tmA;
    op->tmY = op->tmA;
    op->p.M = Mc;
    op->p.N = Nc;
    op->p.K = static_cast<int>(P);
    op->p.num_m_tiles = (Mc + 127) / 128;
    op->p.num_n_tiles = Nc / op->bn;
    const int total_kb = static_cast<int>((P + 63) / 64);
    int splits = num_sms() / (op->p.num_m_tiles * op->p.num_n_tiles);
    if (splits < 1) splits = 1;
    if (splits > total_kb) splits = total_kb;
    op->p.kb_per_split = (total_kb + splits - 1) / splits;
    op->p.num_splits = (total_kb + op->p.kb_per_split - 1) / op->p.kb_per_split;
    op->p.out_f32 = out;
    op->p.ldc = ldc;
    op->p.keep_scale = 1.f;
    op->ready = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// workspace carving
// ------------------------------------------------------------------------------------------------
constexpr int RAG_MAX_STRIPS = 1024;
constexpr int GRAM_MAX_SPLITS = 160;         // >= SM count: one partial tile per CTA of the forward Gram GEMM
// folded conv5 scratch that is zeroed once per backward: Q (1024 x 128, accumulated by the split-K weight-gradient GEMM)
constexpr size_t FOLD4_ZERO_FLOATS = 1024 * 128;

struct Carver {
    uint8_t* base;
    size_t off = 0;
    explicit Carver(void* b) : base(static_cast<uint8_t*>(b)) {}
    template <typename T>
    T* take(size_t count) {
        off = (off + 1023) & ~size_t(1023);
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
};

typedef __nv_bfloat16 bf16;

}  // namespace

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct pcseg_ctx {
    int C = 0;
    Layout L;
    int B = 0, N = 0;
    long long P = 0;
    bool bound = false, train = false, eval_ready = false;
    bool x3 = false;              // inference with split-bf16 ("bf16x3") operands: fp32-grade logits (bind mode 2)

    // ---- shared small buffers
    float* zeros1024 = nullptr;
    float* gmax = nullptr;        // [B][1024] eval max-pool (float bits >= 0) / train g
    float* cb = nullptr;          // [B][512]
    // ---- eval
    float* alpha[NUM_BN] = {};    // folded BN scale per layer
    float* delta[NUM_BN] = {};    // folded bias per layer
    float* w1 = nullptr;          // conv1 weight copy [64][4]
    float* w4 = nullptr;          // seg_conv4 weight copy [C][128]
    float* b4 = nullptr;
    float* wg = nullptr;          // seg_conv1.weight[:, 64:] dense copy [512][1024] fp32
    bf16* wk[NUM_BN] = {};        // bf16 [Cout][Cin] forward weights (index = conv index; [6] = point-feature part [512][64])
    bf16* act[NUM_BN] = {};       // eval: activations a_l; train: post-BN/ReLU activations
    // GEMM operations come in two sets: [0] the dense (B, N) batch, [1] the packed ragged batch (same buffers, tensor
    // maps spanning the whole row capacity, row counts patched per call by rag_plan)
    struct OpSet {
        GemmOp ev[NUM_BN];            // eval forward GEMMs (index = conv index 1..8)
        CUtensorMap hcA1;             // fused inference head (head_chain_kernel): point_feat operand
        HeadChainParams hcp;
        GemmOp fw[NUM_BN], dg[NUM_BN], wg_op[NUM_BN];   // training
        GemmOp fwx[NUM_BN];           // training forward with the BN-apply of the layer below fused in (transform stage, XF)
    };
    OpSet ops[2];
    CUtensorMap hcB1, hcB2, hcB3;
    bool use_head_chain = true;
    // ---- ragged execution (k_pack_rows): device tables, packed input / logits / labels, row multiplicities
    long long cap_rows = 0;       // B * ceil128(N): rows every point-major buffer can hold
    int* meta = nullptr;          // len[B] | off[B+1] | tile_cloud[rows/128] | strips[rag_strips][4]
    float* xpack = nullptr;       // [cap_rows][4]
    float* lpack = nullptr;       // [cap_rows][C]
    float* dlbuf = nullptr;       // [cap_rows][C] loss gradient wrt the logits (train, more than 8 classes only)
    long long* labpack = nullptr; // [cap_rows]      (train)
    float* rowmult = nullptr;     // [cap_rows]      (train)
    std::vector<int> meta_host;
    int rag_N = 0;                // padded length of the current ragged batch (<= N of the binding)
    long long rag_rows = 0;       // packed rows of the current ragged batch
    int rag_strips = 0;           // row strips of k_bn_bwd_apply<., true> (planned per batch)
    bool rag_active = false;      // the latest training forward was ragged (backward follows it)
    // ---- train
    bf16* y[NUM_BN] = {};         // pre-BN conv outputs
    bf16* dz[NUM_BN] = {};        // gradient wrt BN output (after ReLU / dropout mask)
    bf16* dy[NUM_BN] = {};        // gradient wrt conv output
    bf16* dycat = nullptr;        // [P][576] = [dy(conv3) | dy(seg_conv1)]
    bf16* wt[NUM_BN] = {};        // bf16 transposed weights [Cin][Cout] for dgrad (index = conv index)
    bf16* wcat = nullptr;         // [64][576] = [W3^T | Wpf^T]
    double* stats_f = nullptr;    // forward  stats, per layer [2][C] at stat_off
    double* stats_b = nullptr;    // backward stats
    size_t stat_off[NUM_BN] = {};
    size_t stat_total = 0;
    float4* bnp[NUM_BN] = {};
    float4* coef[NUM_BN] = {};
    unsigned long long* keys = nullptr;
    float* ystar = nullptr;
    int* argidx = nullptr;
    float* dcb = nullptr;
    float* dzv = nullptr;
    // ---- Gram-predicted BatchNorm + folded BatchNorm backward (dense batches).  Index = conv index of the
    // layer whose y / dy are never materialised (4 = conv5); gram / colsum describe its INPUT activation.
    bool folded = true;
    float* gramf[NUM_BN] = {};    // [Ci][Ci] fp32  a_prev^T a_prev
    double* colsum[NUM_BN] = {};  // [Ci] fp64      sum_p a_prev   (directly behind gramf: one memset)
    float* qraw[NUM_BN] = {};     // [Co][Ci] fp32  dz^T a_prev
    bf16* bwf[NUM_BN] = {};       // [Ci][Co + Ci]  data-gradient weights [diag(A) W ; S]^T
    float* cstf[NUM_BN] = {};     // [Ci]           constant row of the data gradient
    float* gcf[NUM_BN] = {};      // [Ci][Ci] centred Gram matrix
    float* grampart4 = nullptr;   // [GRAM_MAX_SPLITS][128][128] per-split partial tiles of a3^T a3 (summed by k_gram_reduce)
    // folded global_feat (index 5): gramf[5] = a4^T a4 (upper triangle), qraw[5] | cstf[5] contiguous (one memset)
    bf16* wb5 = nullptr;          // [1024][1024] diag(Bc) W5
    bf16* s5b = nullptr;          // [1024][1024] S5 = W5^T diag(Bc) W5 (symmetric; B operand of the data-gradient GEMM)
    bf16* gc5b = nullptr;         // [1024][1024] centred Gram matrix of a4
    float* side5 = nullptr;       // [B*1024][1024] rows of dz5 diag(A) W5 (max-pool gradient rows)
    int* rowslot5 = nullptr;      // [cap_rows] side-buffer slot of every point (>= B*1024: none)
    bool store_y5 = false;        // PCSEG_STORE_Y5=1: keep writing global_feat's pre-BN output (tests)
    bool xf_seg3 = false;         // PCSEG_XF_SEG3=1: seg_conv3 too (measured slower: 4 transform warps per SM generate the Philox
                                  // dropout mask of 256 channels more slowly than k_bn_relu does with the whole machine)
    bool xf = false;              // PCSEG_XF=1, dense training forward: conv2 / conv3 / conv4 apply the BatchNorm + ReLU of their input
                                  // inside the GEMM (gemm_kernel XF) instead of a k_bn_relu launch.  OFF by default: measured on
                                  // cfg2, the three GEMMs grow by 9 + 17 + 4 us (4 transform warps per SM in front of the MMA of a
                                  // latency-bound 7-tiles-per-CTA kernel) while three 12-13 us launches disappear: 1.869 vs 1.867 ms
    GemmOp s5_op, t5_op;
    // folded seg_conv1 (index 6; needs N % 128 == 0 so that tiles never straddle clouds): per-cloud Gram matrices of a1
    bool fold6 = false;
    float* grampart6 = nullptr;   // [splits][64][64] partial tiles (splits of one cloud are consecutive)
    float* cloudsum6 = nullptr;   // [B][512] per-cloud sums of dz6; cloudsum6 | qraw[6] contiguous (one memset)
    float* gsum6 = nullptr;       // [64][64] sum_b G_b
    double* ssum6 = nullptr;      // [64]     sum_b s_b
    float* cst6 = nullptr;        // [B][64]  per-cloud constant rows of the data gradient into point_feat
    bf16* wcat6 = nullptr;        // [64][640] = [W3^T | diag(A) Wpf^T | S6]
    double* part6 = nullptr;      // [B][512][2] {lin, quad} of every (cloud, channel) (k_predict_bn_cloud)
    int* ticket6 = nullptr;       // [64] completion counters of the channel groups (zero between launches)
    GemmOp gram_op[NUM_BN];
    unsigned long long seed = 0;
    const unsigned long long* seed_ptr = nullptr;
    unsigned int thr16 = 0;
    float keep_scale = 1.f;
    // optional per-GEMM event timing
    bool profiling = false;
    struct Stamp { cudaEvent_t a, b; int tag; };
    std::vector<Stamp> stamps;
};

static int timed_gemm(pcseg_ctx* c, const GemmOp& op, int tag, cudaStream_t s) {
    if (!c->profiling) return launch_gemm(op, s);
    pcseg_ctx::Stamp st;
    st.tag = tag;
    CUDA_OK(cudaEventCreate(&st.a));
    CUDA_OK(cudaEventCreate(&st.b));
    CUDA_OK(cudaEventRecord(st.a, s));
    int r = launch_gemm(op, s);
    CUDA_OK(cudaEventRecord(st.b, s));
    c->stamps.push_back(st);
    return r;
}

// event stamps around arbitrary launches while profiling (tags >= 80: CUDA-core kernels, see engine.profile_read)
struct StampScope {
    pcseg_ctx* c;
    cudaStream_t s;
    pcseg_ctx::Stamp st;
    bool on;
    StampScope(pcseg_ctx* c_, int tag, cudaStream_t s_) : c(c_), s(s_), on(c_->profiling) {
        if (!on) return;
        st.tag = tag;
        cudaEventCreate(&st.a);
        cudaEventCreate(&st.b);
        cudaEventRecord(st.a, s);
    }
    ~StampScope() {
        if (!on) return;
        cudaEventRecord(st.b, s);
        c->stamps.push_back(st);
    }
};

static int carve(pcseg_ctx* c, void* ws, int B, int N, bool train, size_t* bytes_out) {
    const size_t x3f = (c->x3 && !train) ? 2 : 1;      // split-bf16 inference stores [hi | lo] halves
    // Shape-independent buffers (weights, per-channel vectors) come first so that their addresses do not depend on
    // (B, N): bindings of different batch shapes can then share one caller-owned workspace and keep prepared weights.
    Carver k(ws);
    // rows are sized for the packed ragged layout as well: every cloud rounded up to a multiple of 128 rows
    const size_t P = static_cast<size_t>(B) * ((static_cast<size_t>(N) + 127) / 128 * 128);
    c->cap_rows = static_cast<long long>(P);
    const ConvDef* cv = c->L.conv;
    c->zeros1024 = k.take<float>(1024);
    c->w1 = k.take<float>(256);
    c->w4 = k.take<float>(API_MAX_CLASSES * 128);
    c->b4 = k.take<float>(API_MAX_CLASSES);
    c->wg = k.take<float>(512 * 1024);
    for (int i = 0; i < NUM_BN; ++i) {
        c->alpha[i] = k.take<float>(cv[i].cout);
        c->delta[i] = k.take<float>(cv[i].cout);
    }
    for (int i = 1; i < NUM_BN; ++i) {
        const int cin = (i == 6) ? 64 : cv[i].cin;
        c->wk[i] = k.take<bf16>(static_cast<size_t>(cv[i].cout) * cin * x3f);
    }
    if (train) {
        size_t so = 0;
        for (int i = 0; i < NUM_BN; ++i) { c->stat_off[i] = so; so += 2 * static_cast<size_t>(cv[i].cout); }
        c->stat_total = so;
        c->stats_f = k.take<double>(so);
        c->stats_b = k.take<double>(so);
        for (int i = 0; i < NUM_BN; ++i) {
            c->bnp[i] = k.take<float4>(cv[i].cout);
            c->coef[i] = k.take<float4>(cv[i].cout);
        }
        for (int i = 1; i < NUM_BN; ++i) {
            const int cin = (i == 6) ? 64 : cv[i].cin;
            c->wt[i] = k.take<bf16>(static_cast<size_t>(cv[i].cout) * cin);
        }
        c->wcat = k.take<bf16>(64 * 576);
        {   // folded conv5: Ci = 128, Co = 1024
            c->gramf[4] = k.take<float>(128 * 128 + 2 * 128);      // + colsum (fp64) right behind
            c->colsum[4] = reinterpret_cast<double*>(c->gramf[4] ? c->gramf[4] + 128 * 128 : nullptr);
            c->qraw[4] = k.take<float>(FOLD4_ZERO_FLOATS);
            c->cstf[4] = k.take<float>(128);
            c->gcf[4] = k.take<float>(128 * 128);
            c->grampart4 = k.take<float>(static_cast<size_t>(GRAM_MAX_SPLITS) * 128 * 128);
            c->bwf[4] = k.take<bf16>(128 * (1024 + 128));
        }
        {   // folded global_feat: Ci = Co = 1024
            c->gramf[5] = k.take<float>(1024 * 1024);
            c->qraw[5] = k.take<float>(1024 * 1024 + 1024);
            c->cstf[5] = c->qraw[5] ? c->qraw[5] + 1024 * 1024 : nullptr;
            c->wb5 = k.take<bf16>(1024 * 1024);
            c->s5b = k.take<bf16>(1024 * 1024);
            c->gc5b = k.take<bf16>(1024 * 1024);
        }
    }
    // ---- shape-dependent buffers
    c->gmax = k.take<float>(static_cast<size_t>(B) * 1024);
    c->cb = k.take<float>(static_cast<size_t>(B) * 512);
    c->meta = k.take<int>(2 * static_cast<size_t>(B) + 1 + P / 128 + 4 * (static_cast<size_t>(RAG_MAX_STRIPS) + B));
    c->xpack = k.take<float>(P * 4);
    c->lpack = k.take<float>(P * c->C);
    if (train) {
        c->labpack = k.take<long long>(P);
        c->rowmult = k.take<float>(P);
        c->dlbuf = (c->C > MAX_CLASSES) ? k.take<float>(P * c->C) : nullptr;
    }
    if (!train) {
        // a1..a5, a_s1, a_s2 (global_feat output is reduced in-kernel; seg_conv3 output feeds the fused logits epilogue)
        for (int i = 0; i < NUM_BN; ++i) {
            if (i == 5 || (i == 8 && c->C <= MAX_CLASSES)) continue;      // (wide head: seg_conv3's output is materialised)
            c->act[i] = k.take<bf16>(P * cv[i].cout * x3f);
        }
    } else {
        c->keys = k.take<unsigned long long>(static_cast<size_t>(B) * 1024);
        c->ystar = k.take<float>(static_cast<size_t>(B) * 1024);
        c->argidx = k.take<int>(static_cast<size_t>(B) * 1024);
        c->dcb = k.take<float>(static_cast<size_t>(B) * 512);
        c->dzv = k.take<float>(static_cast<size_t>(B) * 1024);
        c->side5 = k.take<float>(static_cast<size_t>(B) * 1024 * 1024);
        c->rowslot5 = k.take<int>(P);
        {   // folded seg_conv1: Ci = 64, Co = 512
            c->gramf[6] = k.take<float>(static_cast<size_t>(B) * 64 * 64);
            c->colsum[6] = k.take<double>(static_cast<size_t>(B) * 64);
            c->grampart6 = k.take<float>((static_cast<size_t>(B) + GRAM_MAX_SPLITS) * 64 * 64);
            c->cloudsum6 = k.take<float>(static_cast<size_t>(B) * 512 + 512 * 64);
            c->qraw[6] = c->cloudsum6 ? c->cloudsum6 + static_cast<size_t>(B) * 512 : nullptr;
            c->gsum6 = k.take<float>(64 * 64);
            c->ssum6 = k.take<double>(64);
            c->cst6 = k.take<float>(static_cast<size_t>(B) * 64);
            c->wcat6 = k.take<bf16>(64 * 640);
            c->part6 = k.take<double>(static_cast<size_t>(B) * 512 * 2);
            c->ticket6 = k.take<int>(64);
        }
        c->dycat = k.take<bf16>(P * 576);
        for (int i = 0; i < NUM_BN; ++i) {
            c->y[i] = k.take<bf16>(P * cv[i].cout);
            if (i != 5 && i != 8) c->act[i] = k.take<bf16>(P * cv[i].cout);
            if (i != 5) c->dz[i] = k.take<bf16>(P * cv[i].cout);
            if (i != 2 && i != 6) c->dy[i] = k.take<bf16>(P * cv[i].cout);
        }
    }
    *bytes_out = (k.off + 1023) & ~size_t(1023);
    return 0;
}

extern "C" long long pcseg_param_count(int C) { return make_layout(C).total; }
extern "C" long long pcseg_param_offset(int C, int t) { return (t < 0 || t >= PCSEG_NUM_PARAM_TENSORS) ? -1 : make_layout(C).off[t]; }
extern "C" long long pcseg_param_numel(int C, int t) { return (t < 0 || t >= PCSEG_NUM_PARAM_TENSORS) ? -1 : make_layout(C).numel[t]; }
extern "C" long long pcseg_bn_buffer_count(void) { return make_layout(3).bn_total; }
extern "C" long long pcseg_bn_buffer_offset(int bn, int which) {
    return (bn < 0 || bn >= NUM_BN || which < 0 || which > 1) ? -1 : make_layout(3).bn_off[bn][which];
}
extern "C" long long pcseg_workspace_bytes(int B, int N, int C, int train) {
    if (B <= 0 || N <= 0 || C < 1 || C > API_MAX_CLASSES) return -1;
    pcseg_ctx tmp;
    tmp.C = C;
    tmp.L = make_layout(C);
    size_t bytes = 0;
    tmp.x3 = train == 2;
    carve(&tmp, nullptr, B, N, train == 1, &bytes);
    return static_cast<long long>(bytes);
}

extern "C" int pcseg_set_sm_limit(int n) {
    g_sm_limit = n > 0 ? n : 0;
    return 0;
}

extern "C" int pcseg_create(pcseg_ctx** out, int num_classes) {
    if (!out) return fail("pcseg_create: null out pointer");
    if (num_classes < 1 || num_classes > API_MAX_CLASSES) return fail("num_classes=%d unsupported (1..%d)", num_classes, API_MAX_CLASSES);
    pcseg_ctx* c = new pcseg_ctx();
    c->C = num_classes;
    c->L = make_layout(num_classes);
    {
        const char* e = getenv("PCSEG_FOLDED");       // 0: legacy training step (y / dy of every layer materialised)
        c->folded = !(e && e[0] == '0');
        const char* y5 = getenv("PCSEG_STORE_Y5");
        c->store_y5 = y5 && y5[0] == '1';
    }
    *out = c;
    return 0;
}
extern "C" int pcseg_destroy(pcseg_ctx* c) {
    delete c;
    return 0;
}

extern "C" int pcseg_bind(pcseg_ctx* c, int B, int N, void* ws, long long ws_bytes, int train) {
    if (!c) return fail("pcseg_bind: null ctx");
    if (B <= 0 || N <= 0) return fail("pcseg_bind: bad shape B=%d N=%d", B, N);
    if (static_cast<long long>(B) * ((static_cast<long long>(N) + 127) / 128 * 128) >= (1LL << 31) - 256)
        return fail("pcseg_bind: B*N too large for 32-bit row indices");
    if (!ws || (reinterpret_cast<uintptr_t>(ws) & 1023)) return fail("pcseg_bind: workspace must be non-null and 1024-byte aligned");
    size_t need = 0;
    c->B = B; c->N = N; c->P = static_cast<long long>(B) * N;
    if (train < 0 || train > 2) return fail("pcseg_bind: mode must be 0 (inference), 1 (training) or 2 (split-bf16 inference)");
    c->train = train == 1;
    c->x3 = train == 2;
    if (c->x3 && c->C > MAX_CLASSES) return fail("pcseg_bind: split-bf16 inference supports up to %d classes", MAX_CLASSES);
    carve(c, ws, B, N, c->train, &need);
    if (static_cast<long long>(need) > ws_bytes) return fail("pcseg_bind: workspace too small (%lld < %zu)", ws_bytes, need);
    const ConvDef* cv = c->L.conv;
    c->bound = false;
    c->eval_ready = false;
    c->rag_active = false;
    c->fold6 = false;
    c->xf = false;
    for (int set = 0; set < 2; ++set) {
    pcseg_ctx::OpSet& O = c->ops[set];
    const bool rag = set == 1;
    const long long P = rag ? c->cap_rows : c->P;       // rows spanned by the tensor maps (ragged: patched per call)
    const int* tile_cloud = rag ? c->meta + 2 * B + 1 : nullptr;
    const int* cloud_off = rag ? c->meta + B : nullptr;
    if (!c->train && c->x3) {
        // split-bf16 inference (dense batches only): every GEMM runs the three-product k-schedule on [hi | lo] operands
        if (rag) continue;
        for (int i = 1; i <= 4; ++i) {
            TRY(setup_gemm_x3(&O.ev[i], EPI_BIAS_RELU_X3, c->act[i - 1], c->wk[i], P, cv[i].cout, cv[i].cin, c->act[i]));
            O.ev[i].p.bias = c->delta[i];
        }
        TRY(setup_gemm_x3(&O.ev[5], EPI_COLMAX, c->act[4], c->wk[5], P, 1024, 1024, nullptr));
        O.ev[5].bn = 256;
        O.ev[5].p.bias = c->delta[5];
        O.ev[5].p.colmax = reinterpret_cast<unsigned int*>(c->gmax);
        O.ev[5].p.pts_per_cloud = N;
        TRY(setup_gemm_x3(&O.ev[6], EPI_BIAS_RELU_X3, c->act[1], c->wk[6], P, 512, 64, c->act[6]));
        O.ev[6].p.bias = c->zeros1024;
        O.ev[6].p.cloud_bias = c->cb;
        O.ev[6].p.pts_per_cloud = N;
        TRY(setup_gemm_x3(&O.ev[7], EPI_BIAS_RELU_X3, c->act[6], c->wk[7], P, 256, 512, c->act[7]));
        O.ev[7].p.bias = c->delta[7];
        TRY(setup_gemm_x3(&O.ev[8], EPI_LOGITS, c->act[7], c->wk[8], P, 128, 256, nullptr));
        O.ev[8].p.bias = c->delta[8];
        O.ev[8].p.w4 = c->w4;
        O.ev[8].p.b4 = c->b4;
        O.ev[8].p.num_classes = c->C;
        c->use_head_chain = false;
    } else if (!c->train) {
        // conv2..conv5
        for (int i = 1; i <= 4; ++i) {
            TRY(setup_gemm_kmajor(&O.ev[i], EPI_BIAS_RELU, c->act[i - 1], cv[i].cin, c->wk[i], cv[i].cin, P, cv[i].cout, cv[i].cin,
                                  c->act[i], cv[i].cout, nullptr, 0));
            O.ev[i].p.bias = c->delta[i];
        }
        // global_feat + max-pool
        TRY(setup_gemm_kmajor(&O.ev[5], EPI_COLMAX, c->act[4], 1024, c->wk[5], 1024, P, 1024, 1024, nullptr, 0, nullptr, 0));
        O.ev[5].p.bias = c->delta[5];
        O.ev[5].p.colmax = reinterpret_cast<unsigned int*>(c->gmax);
        O.ev[5].p.pts_per_cloud = N;
        O.ev[5].p.tile_cloud = tile_cloud;
        if (!rag) TRY(make_pair_op(&O.ev[5], c->wk[5], 1024));
        // seg_conv1: point-feature part + per-cloud bias
        TRY(setup_gemm_kmajor(&O.ev[6], EPI_BIAS_RELU, c->act[1], 64, c->wk[6], 64, P, 512, 64, c->act[6], 512, nullptr, 0));
        O.ev[6].p.bias = c->zeros1024;
        O.ev[6].p.cloud_bias = c->cb;
        O.ev[6].p.pts_per_cloud = N;
        O.ev[6].p.tile_cloud = tile_cloud;
        TRY(setup_gemm_kmajor(&O.ev[7], EPI_BIAS_RELU, c->act[6], 512, c->wk[7], 512, P, 256, 512, c->act[7], 256, nullptr, 0));
        O.ev[7].p.bias = c->delta[7];
        if (c->C <= MAX_CLASSES) {
            TRY(setup_gemm_kmajor(&O.ev[8], EPI_LOGITS, c->act[7], 256, c->wk[8], 256, P, 128, 256, nullptr, 0, nullptr, 0));
            O.ev[8].p.bias = c->delta[8];
            O.ev[8].p.w4 = c->w4;
            O.ev[8].p.b4 = c->b4;
            O.ev[8].p.num_classes = c->C;
        } else {        // more than 8 classes: seg_conv3 as a plain GEMM, seg_conv4 by k_head_fwd<16|32, false>
            TRY(setup_gemm_kmajor(&O.ev[8], EPI_BIAS_RELU, c->act[7], 256, c->wk[8], 256, P, 128, 256, c->act[8], 128, nullptr, 0));
            O.ev[8].p.bias = c->delta[8];
        }
        // fused head: seg_conv1..4 in one kernel (PCSEG_EVAL_CHAIN=0 selects the layer-by-layer kernels above)
        TRY(make_tmap(&O.hcA1, c->act[1], 64, P, 64, 64, 128));
        if (!rag) TRY(make_tmap(&c->hcB1, c->wk[6], 64, 512, 64, 64, 256));
        if (!rag) TRY(make_tmap(&c->hcB2, c->wk[7], 512, 256, 512, 64, 256));
        if (!rag) TRY(make_tmap(&c->hcB3, c->wk[8], 256, 128, 256, 64, 128));
        memset(&O.hcp, 0, sizeof(O.hcp));
        O.hcp.M = static_cast<int>(P);
        O.hcp.num_tiles = static_cast<int>((P + 127) / 128);
        O.hcp.pts_per_cloud = N;
        O.hcp.tile_cloud = tile_cloud;
        O.hcp.cloud_bias = c->cb;
        O.hcp.bias2 = c->delta[7];
        O.hcp.bias3 = c->delta[8];
        O.hcp.w4 = c->w4;
        O.hcp.b4 = c->b4;
        O.hcp.num_classes = c->C;
        {
            const char* e = getenv("PCSEG_EVAL_CHAIN");
            c->use_head_chain = !(e && e[0] == '0') && c->C <= MAX_CLASSES;
        }
    } else {
        // ---- forward: y_i = a_{i-1} W_i^T, statistics in the epilogue
        for (int i = 1; i <= 5; ++i) {
            TRY(setup_gemm_kmajor(&O.fw[i], i == 5 ? EPI_STATS_POOL : EPI_STATS, c->act[i - 1], cv[i].cin, c->wk[i], cv[i].cin, P,
                                  cv[i].cout, cv[i].cin, c->y[i], cv[i].cout, nullptr, 0));
            O.fw[i].p.stats = c->stats_f + c->stat_off[i];
        }
        O.fw[5].p.pool_keys = c->keys;          // train-mode max-pool fused into global_feat's epilogue
        O.fw[5].p.pts_per_cloud = N;
        O.fw[5].p.tile_cloud = tile_cloud;
        O.fw[5].p.cloud_off = cloud_off;
        if (!rag) TRY(make_pair_op(&O.fw[5], c->wk[5], 1024));
        TRY(setup_gemm_kmajor(&O.fw[6], EPI_STATS, c->act[1], 64, c->wk[6], 64, P, 512, 64, c->y[6], 512, nullptr, 0));
        O.fw[6].p.stats = c->stats_f + c->stat_off[6];
        O.fw[6].p.cloud_bias = c->cb;
        O.fw[6].p.pts_per_cloud = N;
        O.fw[6].p.tile_cloud = tile_cloud;
        for (int i = 7; i <= 8; ++i) {
            TRY(setup_gemm_kmajor(&O.fw[i], EPI_STATS, c->act[i - 1], cv[i].cin, c->wk[i], cv[i].cin, P, cv[i].cout, cv[i].cin,
                                  c->y[i], cv[i].cout, nullptr, 0));
            O.fw[i].p.stats = c->stats_f + c->stat_off[i];
        }
        // ---- the same forward GEMMs reading the PRE-BatchNorm output of the layer below (transform stage): conv2, conv3, conv4,
        //      seg_conv3.  (conv5 / seg_conv1 need their input activation earlier, for the Gram matrices; seg_conv2's input is
        //      written by seg_conv1's epilogue.)  Tiles must not straddle clouds (per-cloud column sums of point_feat).
        if (!rag) {
            c->xf = (N % 128 == 0) && getenv("PCSEG_XF") && getenv("PCSEG_XF")[0] == '1';
            c->xf_seg3 = getenv("PCSEG_XF_SEG3") && getenv("PCSEG_XF_SEG3")[0] == '1';
            const int xl[4] = {1, 2, 3, 8};
            for (int j = 0; j < 4; ++j) {
                const int i = xl[j];
                TRY(setup_gemm_kmajor(&O.fwx[i], EPI_STATS, c->y[i - 1], cv[i].cin, c->wk[i], cv[i].cin, P, cv[i].cout, cv[i].cin,
                                      c->y[i], cv[i].cout, nullptr, 0));
                TRY(make_tmap(&O.fwx[i].tmA2, c->act[i - 1], cv[i].cin, P, cv[i].cin, 64, 128));
                O.fwx[i].p.stats = c->stats_f + c->stat_off[i];
                O.fwx[i].p.pts_per_cloud = N;
                O.fwx[i].p.xf_keep_scale = 1.f;
                O.fwx[i].xf = true;
                if (cv[i].cin > 256) return fail("internal: transform stage supports K <= 256");
            }
        }
        // ---- backward data gradients: dz_{i-1} = (dy_i W_i) masked by layer i-1
        auto dgrad = [&](int i, const bf16* dyA, int lda, int K, const bf16* Bw, int prev) -> int {
            TRY(setup_gemm_kmajor(&O.dg[i], EPI_DGRAD, dyA, lda, Bw, K, P, cv[prev].cout, K, c->dz[prev], cv[prev].cout,
                                  c->y[prev], cv[prev].cout));
            O.dg[i].p.stats = c->stats_b + c->stat_off[prev];
            O.dg[i].p.bnp = c->bnp[prev];
            return 0;
        };
        TRY(dgrad(8, c->dy[8], 128, 128, c->wt[8], 7));
        TRY(dgrad(7, c->dy[7], 256, 256, c->wt[7], 6));
        TRY(dgrad(5, c->dy[5], 1024, 1024, c->wt[5], 4));
        TRY(dgrad(4, c->dy[4], 1024, 1024, c->wt[4], 3));
        TRY(dgrad(3, c->dy[3], 128, 128, c->wt[3], 2));
        TRY(dgrad(2, c->dycat, 576, 576, c->wcat, 1));     // skip join: [dy3 | dy_seg1] * [W3 ; Wpf]
        TRY(dgrad(1, c->dy[1], 64, 64, c->wt[1], 0));
        // ---- backward weight gradients (destinations patched with the grads arena at call time)
        auto wgrad = [&](int i, const bf16* dyA, int lda, const bf16* aB, int ldb, int cin) -> int {
            return setup_gemm_wgrad(&O.wg_op[i], dyA, lda, cv[i].cout, aB, ldb, cin, P, nullptr, cv[i].cin);
        };
        TRY(wgrad(8, c->dy[8], 128, c->act[7], 256, 256));
        TRY(wgrad(7, c->dy[7], 256, c->act[6], 512, 512));
        TRY(wgrad(6, c->dycat + 64, 576, c->act[1], 64, 64));
        TRY(wgrad(5, c->dy[5], 1024, c->act[4], 1024, 1024));
        TRY(wgrad(4, c->dy[4], 1024, c->act[3], 128, 128));
        TRY(wgrad(3, c->dy[3], 128, c->act[2], 64, 64));
        TRY(wgrad(2, c->dycat, 576, c->act[1], 64, 64));
        TRY(wgrad(1, c->dy[1], 64, c->act[0], 64, 64));
        if (c->folded && !rag) {
            // conv5 with Gram-predicted statistics: G = a3^T a3, BN + ReLU in the GEMM epilogue, y4 never stored
            TRY(setup_gemm_wgrad(&c->gram_op[4], c->act[3], 128, 128, c->act[3], 128, 128, P, c->grampart4, 128));
            c->gram_op[4].p.wg_mode = 3;        // per-split partial tiles, summed in a fixed order: reproducible statistics
            if (c->gram_op[4].p.num_splits > GRAM_MAX_SPLITS) return fail("internal: %d Gram splits", c->gram_op[4].p.num_splits);
            TRY(setup_gemm_kmajor(&O.fw[4], EPI_BN_RELU, c->act[3], 128, c->wk[4], 128, P, 1024, 128, c->act[4], 1024, nullptr, 0));
            O.fw[4].p.bnp = c->bnp[4];
            // global_feat: statistics + max-pool only (its pre-BN output is not needed by the folded backward)
            O.fw[5].p.store_out = c->store_y5 ? 1 : 0;
            // folded BN backward of global_feat.  S5 = (diag(Bc) W5)^T W5 on the tensor cores, bf16 result in one split
            TRY(setup_gemm_wgrad(&c->s5_op, c->wb5, 1024, 1024, c->wk[5], 1024, 1024, 1024, nullptr, 1024));
            c->s5_op.p.num_splits = 1;
            c->s5_op.p.kb_per_split = 1024 / 64;
            c->s5_op.p.wg_mode = 1;
            c->s5_op.p.out_bf16 = c->s5b;
            // data gradient: dz4 = [a4 > 0] . (a4 S5 + const + max-pool gradient rows), column sums: sum dz4, sum a4
            TRY(setup_gemm_kmajor(&O.dg[5], EPI_DGRAD_ACT, c->act[4], 1024, c->s5b, 1024, P, 1024, 1024, c->dz[4], 1024, c->act[4], 1024));
            O.dg[5].p.stats = c->stats_b + c->stat_off[4];
            O.dg[5].p.bias = c->cstf[5];
            O.dg[5].p.rowslot = c->rowslot5;
            O.dg[5].p.side = c->side5;
            O.dg[5].p.side_rows = B * 1024;
            TRY(make_pair_op(&O.dg[5], c->s5b, 1024));
            // weight gradient: G4 = a4^T a4 (upper-triangle tiles only), then dW5 = A Q5 + Bc (W5 Gc4) + D s4^T in the
            // epilogue of the W5 Gc4 GEMM
            TRY(setup_gemm_wgrad(&c->gram_op[5], c->act[4], 1024, 1024, c->act[4], 1024, 1024, P, c->gramf[5], 1024));
            {
                GemmParams& gp = c->gram_op[5].p;
                int tiles = 0;
                for (int nt = 0; nt < gp.num_n_tiles; ++nt) {
                    const int cnt = (nt + 1) * (c->gram_op[5].bn / 128);
                    tiles += cnt < gp.num_m_tiles ? cnt : gp.num_m_tiles;
                }
                gp.sym_tiles = tiles;
                const int total_kb = static_cast<int>((P + 63) / 64);
                int splits = num_sms() / tiles;
                if (splits < 1) splits = 1;
                if (splits > total_kb) splits = total_kb;
                gp.kb_per_split = (total_kb + splits - 1) / splits;
                gp.num_splits = (total_kb + gp.kb_per_split - 1) / gp.kb_per_split;
            }
            c->fold6 = (N % 128 == 0) && !(getenv("PCSEG_FOLD6") && getenv("PCSEG_FOLD6")[0] == '0');
            if (c->fold6) {
                CUDA_OK(cudaMemset(c->ticket6, 0, 64 * sizeof(int)));
                // seg_conv1: per-cloud Gram matrices of a1 (point_feat), every cloud split on its own
                TRY(setup_gemm_wgrad(&c->gram_op[6], c->act[1], 64, 64, c->act[1], 64, 64, P, c->grampart6, 64));
                GemmParams& gp = c->gram_op[6].p;
                gp.kb_group = N / 64;
                int spg = num_sms() / B;
                if (spg < 1) spg = 1;
                if (spg > gp.kb_group) spg = gp.kb_group;
                gp.kb_per_split = (gp.kb_group + spg - 1) / spg;
                gp.splits_per_group = (gp.kb_group + gp.kb_per_split - 1) / gp.kb_per_split;
                gp.num_splits = B * gp.splits_per_group;
                gp.wg_mode = 3;
                if (gp.num_splits > B + GRAM_MAX_SPLITS) return fail("internal: %d per-cloud Gram splits", gp.num_splits);
                // forward: BN + ReLU + dropout in the epilogue, per-cloud term added before the normalisation
                TRY(setup_gemm_kmajor(&O.fw[6], EPI_BN_RELU_DROP, c->act[1], 64, c->wk[6], 64, P, 512, 64, c->act[6], 512, nullptr, 0));
                O.fw[6].p.bnp = c->bnp[6];
                O.fw[6].p.cloud_bias = c->cb;
                O.fw[6].p.pts_per_cloud = N;
                // seg_conv2 data gradient: dz6 (straight into dycat[:, 64:576]) masked by the stored activation a6, per-cloud sums
                TRY(setup_gemm_kmajor(&O.dg[7], EPI_DGRAD_ACT, c->dy[7], 256, c->wt[7], 256, P, 512, 256, c->dycat + 64, 576, c->act[6], 512));
                O.dg[7].p.stats = c->stats_b + c->stat_off[6];
                O.dg[7].p.cloud_sums = c->cloudsum6;
                O.dg[7].p.pts_per_cloud = N;
                // Q6 = dz6^T a1 (raw), then the skip join: dz1 = mask1 . ([dy2 | dz6 | a1] [W3 ; diag(A) Wpf ; S6] + const_b)
                TRY(setup_gemm_wgrad(&O.wg_op[6], c->dycat + 64, 576, 512, c->act[1], 64, 64, P, c->qraw[6], 64));
                TRY(setup_gemm_kmajor_cat(&O.dg[2], EPI_DGRAD, c->dycat, 576, 576, c->act[1], 64, 64, c->wcat6, 640, P, 64, c->dz[1], 64, c->y[1], 64));
                O.dg[2].p.stats = c->stats_b + c->stat_off[1];
                O.dg[2].p.bnp = c->bnp[1];
                O.dg[2].p.cloud_bias = c->cst6;
                O.dg[2].p.pts_per_cloud = N;
            }
            TRY(setup_gemm_wgrad(&c->t5_op, c->wt[5], 1024, 1024, c->gc5b, 1024, 1024, 1024, nullptr, 1024));
            c->t5_op.p.num_splits = 1;
            c->t5_op.p.kb_per_split = 1024 / 64;
            c->t5_op.p.wg_mode = 2;
            c->t5_op.p.wq = c->qraw[5];
            c->t5_op.p.wcoef = c->coef[5];
            c->t5_op.p.ws = c->stats_b + c->stat_off[4] + 1024;
            // folded BN backward of conv5: Q = dz4^T a3 (raw), then dz3 = mask3 . ([dz4 | a3] [diag(A) W ; S] + const)
            TRY(setup_gemm_wgrad(&O.wg_op[4], c->dz[4], 1024, 1024, c->act[3], 128, 128, P, c->qraw[4], 128));
            TRY(setup_gemm_kmajor_cat(&O.dg[4], EPI_DGRAD, c->dz[4], 1024, 1024, c->act[3], 128, 128, c->bwf[4], 1024 + 128, P, 128, c->dz[3], 128,
                                      c->y[3], 128));
            O.dg[4].p.stats = c->stats_b + c->stat_off[3];
            O.dg[4].p.bnp = c->bnp[3];
            O.dg[4].p.bias = c->cstf[4];
        }
    }
    }   // op sets
    c->bound = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// weight preparation
// ------------------------------------------------------------------------------------------------
static int convert_rows(const float* src, int ld_src, bf16* dst, int ld_dst, int rows, int cols, const float* alpha, cudaStream_t s,
                        bf16* dst_lo = nullptr) {
    const int n = rows * cols;
    pdl_launch(k_convert_rows, (n + 255) / 256, 256, 0, s, src, ld_src, dst, ld_dst, rows, cols, alpha, dst_lo);
    LAUNCH_OK("k_convert_rows");
    return 0;
}
static int convert_transpose(const float* src, int ld_src, bf16* dst, int ld_dst, int rows, int cols, cudaStream_t s) {
    dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
    pdl_launch(k_convert_transpose, grid, block, 0, s, src, ld_src, dst, ld_dst, rows, cols);
    LAUNCH_OK("k_convert_transpose");
    return 0;
}

extern "C" int pcseg_prepare_eval(pcseg_ctx* c, const float* params, const float* bnbuf, void* stream) {
    g_pdl_call = true;
    if (!c || !c->bound || c->train) return fail("pcseg_prepare_eval: context not bound in eval mode");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const Layout& L = c->L;
    CUDA_OK(cudaMemsetAsync(c->zeros1024, 0, 1024 * sizeof(float), s));
    for (int i = 0; i < NUM_BN; ++i) {
        const int co = L.conv[i].cout;
        pdl_launch(k_fold_bn, (co + 127) / 128, 128, 0, s, params + L.off[2 * i + 1], params + L.off[20 + 2 * i], params + L.off[21 + 2 * i],
                                                   bnbuf + L.bn_off[i][0], bnbuf + L.bn_off[i][1], BN_EPS, co, c->alpha[i], c->delta[i]);
        LAUNCH_OK("k_fold_bn");
    }
    CUDA_OK(cudaMemcpyAsync(c->w1, params + L.off[0], 256 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    CUDA_OK(cudaMemcpyAsync(c->w4, params + L.off[18], static_cast<size_t>(c->C) * 128 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    CUDA_OK(cudaMemcpyAsync(c->b4, params + L.off[19], static_cast<size_t>(c->C) * sizeof(float), cudaMemcpyDeviceToDevice, s));
    for (int i = 1; i < NUM_BN; ++i) {
        const int xf = c->x3 ? 2 : 1;          // split-bf16: rows of [hi | lo] halves
        if (i == 6) {
            TRY(convert_rows(params + L.off[12], 1088, c->wk[6], 64 * xf, 512, 64, c->alpha[6], s, c->x3 ? c->wk[6] + 64 : nullptr));
            CUDA_OK(cudaMemcpy2DAsync(c->wg, 1024 * sizeof(float), params + L.off[12] + 64, 1088 * sizeof(float), 1024 * sizeof(float), 512,
                                      cudaMemcpyDeviceToDevice, s));
        } else {
            const int ci = L.conv[i].cin;
            TRY(convert_rows(params + L.off[2 * i], ci, c->wk[i], ci * xf, L.conv[i].cout, ci, c->alpha[i], s, c->x3 ? c->wk[i] + ci : nullptr));
        }
    }
    c->eval_ready = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// ragged batches: host-side plan of the packed layout (see k_pack_rows) + upload of the device tables
// ------------------------------------------------------------------------------------------------
static RaggedMeta rag_meta(const pcseg_ctx* c) {
    RaggedMeta m;
    m.len = c->meta;
    m.off = c->meta + c->B;
    m.tile_cloud = c->meta + 2 * c->B + 1;
    m.B = c->B;
    m.Nmax = c->rag_N;
    m.rows = static_cast<int>(c->rag_rows);
    return m;
}
static void patch_rows(GemmOp& op, long long rows) {
    if (!op.ready) return;
    if (op.mn) {          // weight gradient: the points are the reduction dimension
        op.p.K = static_cast<int>(rows);
        const int total_kb = static_cast<int>(rows / 64);
        int splits = num_sms() / (op.p.num_m_tiles * op.p.num_n_tiles);
        if (splits < 1) splits = 1;
        if (splits > total_kb) splits = total_kb;
        op.p.kb_per_split = (total_kb + splits - 1) / splits;
        op.p.num_splits = (total_kb + op.p.kb_per_split - 1) / op.p.kb_per_split;
    } else {
        op.p.M = static_cast<int>(rows);
        op.p.num_m_tiles = static_cast<int>(rows / 128);
    }
}
// Pure host planner of the packed layout (no CUDA calls; exported as pcseg_ragged_plan so that it can be inspected and
// tested without a device).  h must hold 2B + 1 + sum(alloc)/128 + 4 (RAG_MAX_STRIPS + B) ints; returns the ints written.
static long long plan_ragged(int B, int N, const int* lengths, std::vector<int>& h, long long* rows_out, int* strips_out) {
    long long off = 0;
    for (int b = 0; b < B; ++b) {
        const int L = lengths[b];
        if (L < 0 || L > N) return fail("ragged batch: lengths[%d]=%d outside 0..%d", b, L, N), -1;
        const long long alloc = (static_cast<long long>(L) + (L < N ? 1 : 0) + 127) / 128 * 128;
        h[b] = L;
        h[B + b] = static_cast<int>(off);
        for (long long t = off / 128; t < (off + alloc) / 128; ++t) h[2 * B + 1 + t] = b;
        off += alloc;
    }
    h[2 * B] = static_cast<int>(off);
    // row strips for the BN-backward kernels: about two blocks per SM over the packed rows, never across clouds
    size_t w = 2 * static_cast<size_t>(B) + 1 + static_cast<size_t>(off / 128);
    int target = 2 * num_sms();
    if (target > RAG_MAX_STRIPS) target = RAG_MAX_STRIPS;
    const long long tiles = off / 128;
    const long long tps = (tiles + target - 1) / target;        // tiles per strip
    int n = 0;
    for (int b = 0; b < B; ++b) {
        const long long r_begin = h[B + b], r_end = h[B + b + 1];
        for (long long r = r_begin; r < r_end; r += tps * 128, ++n) {
            h[w + 4 * n] = b;
            h[w + 4 * n + 1] = static_cast<int>(r);
            h[w + 4 * n + 2] = static_cast<int>(r + tps * 128 < r_end ? r + tps * 128 : r_end);
            h[w + 4 * n + 3] = static_cast<int>(r_begin);
        }
    }
    *rows_out = off;
    *strips_out = n;
    return static_cast<long long>(w + 4 * static_cast<size_t>(n));
}
static size_t plan_capacity_ints(int B, int N) {
    const size_t cap_rows = static_cast<size_t>(B) * ((static_cast<size_t>(N) + 127) / 128 * 128);
    return 2 * static_cast<size_t>(B) + 1 + cap_rows / 128 + 4 * (static_cast<size_t>(RAG_MAX_STRIPS) + B);
}

extern "C" long long pcseg_ragged_plan(int B, int nmax, const int* lengths, int* meta_out, long long meta_capacity,
                                       long long* rows_out, int* strips_out) {
    if (B <= 0 || nmax <= 0 || !lengths) return fail("pcseg_ragged_plan: bad arguments"), -1;
    std::vector<int> h(plan_capacity_ints(B, nmax), 0);
    long long rows = 0;
    int strips = 0;
    const long long n = plan_ragged(B, nmax, lengths, h, &rows, &strips);
    if (n < 0) return -1;
    if (rows_out) *rows_out = rows;
    if (strips_out) *strips_out = strips;
    if (meta_out) {
        if (meta_capacity < n) return fail("pcseg_ragged_plan: meta buffer too small (%lld < %lld ints)", meta_capacity, n), -1;
        memcpy(meta_out, h.data(), static_cast<size_t>(n) * sizeof(int));
    }
    return n;
}

static int rag_plan(pcseg_ctx* c, const int* lengths, int nmax, cudaStream_t s) {
    if (nmax <= 0) nmax = c->N;
    if (nmax > c->N) return fail("ragged batch: padded length %d exceeds the bound capacity %d", nmax, c->N);
    c->rag_N = nmax;
    const int B = c->B;
    std::vector<int>& h = c->meta_host;
    h.assign(plan_capacity_ints(B, c->N), 0);
    long long off = 0;
    int nstrips = 0;
    const long long w = plan_ragged(B, nmax, lengths, h, &off, &nstrips);
    if (w < 0) return 1;
    if (off > c->cap_rows) return fail("internal: packed rows %lld exceed the capacity %lld", off, c->cap_rows);
    c->rag_rows = off;
    c->rag_strips = nstrips;
    // (pageable source: the copy is staged before the call returns, meta_host may be rewritten by the next plan)
    CUDA_OK(cudaMemcpyAsync(c->meta, h.data(), static_cast<size_t>(w) * sizeof(int), cudaMemcpyHostToDevice, s));
    pcseg_ctx::OpSet& O = c->ops[1];
    for (int i = 0; i < NUM_BN; ++i) {
        patch_rows(O.ev[i], off);
        patch_rows(O.fw[i], off);
        patch_rows(O.dg[i], off);
        patch_rows(O.wg_op[i], off);
    }
    O.hcp.M = static_cast<int>(off);
    O.hcp.num_tiles = static_cast<int>(off / 128);
    return 0;
}
static int rag_pack(pcseg_ctx* c, const float* x, const long long* labels, bool train, cudaStream_t s) {
    int grid = static_cast<int>((c->rag_rows + 255) / 256);
    if (grid > num_sms() * 8) grid = num_sms() * 8;
    pdl_launch(k_pack_rows, grid, 256, 0, s, reinterpret_cast<const float4*>(x), labels, rag_meta(c), reinterpret_cast<float4*>(c->xpack),
               train ? c->labpack : static_cast<long long*>(nullptr), train ? c->rowmult : static_cast<float*>(nullptr));
    LAUNCH_OK("k_pack_rows");
    return 0;
}
static int rag_unpack_logits(pcseg_ctx* c, float* logits, cudaStream_t s) {
    const long long total = static_cast<long long>(c->B) * c->rag_N * c->C;
    int grid = static_cast<int>((total + 255) / 256);
    if (grid > num_sms() * 8) grid = num_sms() * 8;
    pdl_launch(k_unpack_logits, grid, 256, 0, s, static_cast<const float*>(c->lpack), rag_meta(c), c->C, logits);
    LAUNCH_OK("k_unpack_logits");
    return 0;
}

// rows = points to process (dense: B*N; ragged: packed rows), x / logits in that row space
// part 1: ingest .. global_feat + max-pool (the pooled feature of every cloud is left in c->gmax)
static int forward_eval_trunk(pcseg_ctx* c, pcseg_ctx::OpSet& O, long long rows, const float* x, cudaStream_t s) {
    CUDA_OK(cudaMemsetAsync(c->gmax, 0, static_cast<size_t>(c->B) * 1024 * sizeof(float), s));
    {
        int grid = static_cast<int>((rows + 31) / 32);
        if (grid > num_sms() * 8) grid = num_sms() * 8;
        pdl_launch(k_ingest<false>, grid, 256, 0, s, reinterpret_cast<const float4*>(x), static_cast<int>(rows), c->w1, c->alpha[0], c->delta[0],
                                            c->act[0], nullptr, c->x3 ? 1 : 0);
        LAUNCH_OK("k_ingest");
    }
    for (int i = 1; i <= 5; ++i) TRY(timed_gemm(c, O.ev[i], 64 + i, s));      // (event-timed only while profiling)
    return 0;
}
// part 2: per-cloud seg_conv1 term from the pooled feature, segmentation head, logits
static int forward_eval_head(pcseg_ctx* c, pcseg_ctx::OpSet& O, float* logits, cudaStream_t s) {
    {
        const int warps = c->B * 512;
        pdl_launch(k_cloud_bias, (warps * 32 + 255) / 256, 256, 0, s, c->wg, 1024, c->gmax, c->B, 512, 1024, c->alpha[6], c->delta[6], c->cb);
        LAUNCH_OK("k_cloud_bias");
    }
    if (c->use_head_chain) {
        static unsigned long long attr_set_mask = 0;
        int dev = 0;
        cudaGetDevice(&dev);
        if (!(attr_set_mask & (1ull << (dev & 63)))) {
            CUDA_OK(cudaFuncSetAttribute(head_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HC_SMEM_BYTES));
            attr_set_mask |= 1ull << (dev & 63);
        }
        HeadChainParams hp = O.hcp;
        hp.logits = logits;
        const int grid = hp.num_tiles < num_sms() ? hp.num_tiles : num_sms();
        pdl_launch(head_chain_kernel, grid, HC_THREADS, HC_SMEM_BYTES, s, O.hcA1, c->hcB1, c->hcB2, c->hcB3, hp);
        LAUNCH_OK("head_chain_kernel");
    } else if (c->C <= MAX_CLASSES) {
        TRY(launch_gemm(O.ev[6], s));
        TRY(launch_gemm(O.ev[7], s));
        GemmOp head = O.ev[8];
        head.p.logits = logits;
        TRY(launch_gemm(head, s));
    } else {
        TRY(launch_gemm(O.ev[6], s));
        TRY(launch_gemm(O.ev[7], s));
        TRY(launch_gemm(O.ev[8], s));
        const long rows = O.ev[8].p.M;
        int grid = static_cast<int>((rows + 255) / 256);
        if (grid > num_sms() * 4) grid = num_sms() * 4;
        BnFinalizeArgs none;
        memset(&none, 0, sizeof(none));
        if (c->C <= 16)
            pdl_launch(k_head_fwd<16, false>, grid, 256, 0, s, static_cast<const bf16*>(c->act[8]), rows, none, static_cast<const float*>(c->w4),
                       static_cast<const float*>(c->b4), c->C, logits, nullptr, nullptr, nullptr);
        else
            pdl_launch(k_head_fwd<32, false>, grid, 256, 0, s, static_cast<const bf16*>(c->act[8]), rows, none, static_cast<const float*>(c->w4),
                       static_cast<const float*>(c->b4), c->C, logits, nullptr, nullptr, nullptr);
        LAUNCH_OK("k_head_fwd<wide>");
    }
    return 0;
}
static int forward_eval_rows(pcseg_ctx* c, pcseg_ctx::OpSet& O, long long rows, const float* x, float* logits, cudaStream_t s) {
    TRY(forward_eval_trunk(c, O, rows, x, s));
    return forward_eval_head(c, O, logits, s);
}
static int argmax_labels(pcseg_ctx* c, long long P, const float* logits, long long* labels_out, cudaStream_t s) {
    pdl_launch(k_argmax, static_cast<int>((P + 255) / 256), 256, 0, s, logits, P, c->C, labels_out);
    LAUNCH_OK("k_argmax");
    return 0;
}

extern "C" int pcseg_forward_eval(pcseg_ctx* c, const float* x, float* logits, long long* labels_out, void* stream) {
    g_pdl_call = true;
    if (!c || !c->bound || c->train) return fail("pcseg_forward_eval: context not bound in eval mode");
    if (!c->eval_ready) return fail("pcseg_forward_eval: call pcseg_prepare_eval first");
    if (!x || !logits) return fail("pcseg_forward_eval: null tensor");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    TRY(forward_eval_rows(c, c->ops[0], c->P, x, logits, s));
    if (labels_out) TRY(argmax_labels(c, c->P, logits, labels_out, s));
    return 0;
}

extern "C" int pcseg_forward_eval_part(pcseg_ctx* c, const float* x, float* logits, long long* labels_out, int part, void* stream) {
    g_pdl_call = true;
    if (!c || !c->bound || c->train) return fail("pcseg_forward_eval_part: context not bound in eval mode");
    if (!c->eval_ready) return fail("pcseg_forward_eval_part: call pcseg_prepare_eval first");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (part == 1) {
        if (!x) return fail("pcseg_forward_eval_part: null input");
        return forward_eval_trunk(c, c->ops[0], c->P, x, s);
    }
    if (part == 2) {
        if (!logits) return fail("pcseg_forward_eval_part: null logits");
        TRY(forward_eval_head(c, c->ops[0], logits, s));
        if (labels_out) TRY(argmax_labels(c, c->P, logits, labels_out, s));
        return 0;
    }
    return fail("pcseg_forward_eval_part: part must be 1 or 2");
}
extern "C" int pcseg_pooled_feature(pcseg_ctx* c, float** pooled) {
    if (!c || !c->bound || !pooled) return fail("pcseg_pooled_feature: context not bound");
    *pooled = c->gmax;
    return 0;
}

extern "C" int pcseg_forward_eval_ragged(pcseg_ctx* c, const float* x, const int* lengths, int nmax, float* logits,
                                         long long* labels_out, void* stream) {
    g_pdl_call = true;
    if (!c || !c->bound || c->train) return fail("pcseg_forward_eval_ragged: context not bound in eval mode");
    if (!c->eval_ready) return fail("pcseg_forward_eval_ragged: call pcseg_prepare_eval first");
    if (!x || !logits || !lengths) return fail("pcseg_forward_eval_ragged: null argument");
    if (c->x3) return fail("pcseg_forward_eval_ragged: split-bf16 inference runs dense batches only");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    TRY(rag_plan(c, lengths, nmax, s));
    TRY(rag_pack(c, x, nullptr, false, s));
    TRY(forward_eval_rows(c, c->ops[1], c->rag_rows, c->xpack, c->lpack, s));
    TRY(rag_unpack_logits(c, logits, s));
    if (labels_out) TRY(argmax_labels(c, static_cast<long long>(c->B) * c->rag_N, logits, labels_out, s));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// training forward
// ------------------------------------------------------------------------------------------------
// grid for the row-strip elementwise kernels: enough blocks to fill the GPU, each with >= 4 passes of rows
static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return (e && e[0]) ? atoi(e) : dflt;
}
static int strip_grid(long long rows, int C) {
    const int rpp = 256 / (C / 8);
    static const int passes = env_int("PCSEG_STRIP_PASSES", 4);       // loop iterations (of 4 x rpp rows) per block
    long long g = rows / (4LL * rpp * (passes > 0 ? passes : 4));
    const long long cap = static_cast<long long>(num_sms()) * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return static_cast<int>(g);
}
static int ew_grid(long long work_items) {
    long long g = (work_items + 255) / 256;
    const long long cap = static_cast<long long>(num_sms()) * 16;
    return static_cast<int>(g < cap ? (g < 1 ? 1 : g) : cap);
}

// P = rows to process (dense: B*N; ragged: packed rows); BN statistics are always normalised by the logical B*N rows
static int forward_train_rows(pcseg_ctx* c, const bool rag, const float* x, const float* params, float* bnbuf, unsigned long long seed,
                              float dropout_p, float* logits, const long long* labels, const float* class_w,
                              pcseg_ce_accum* ce, const pcseg_step_state* state, cudaStream_t s) {
    pcseg_ctx::OpSet& O = c->ops[rag ? 1 : 0];
    const Layout& L = c->L;
    const ConvDef* cv = L.conv;
    const long long P = rag ? c->rag_rows : c->P;
    c->seed = seed;
    c->seed_ptr = state ? &state->seed : nullptr;
    c->thr16 = static_cast<unsigned int>(dropout_p * 65536.0f + 0.5f);
    c->keep_scale = c->thr16 ? 1.f / (1.f - static_cast<float>(c->thr16) / 65536.f) : 1.f;

    // bf16 weights (forward [Cout][Cin], transposed [Cin][Cout] for dgrad): one launch for all 16 conversions
    {
        ConvertJobs jobs;
        int nj = 0;
        auto add = [&](const float* src, int ld_src, bf16* dst, int ld_dst, int rows, int cols, int tr) {
            jobs.job[nj++] = ConvertJob{src, dst, ld_src, ld_dst, rows, cols, tr};
        };
        for (int i = 1; i < NUM_BN; ++i) {
            if (i == 6) {
                add(params + L.off[12], 1088, c->wk[6], 64, 512, 64, 0);
                add(params + L.off[12], 1088, c->wcat + 64, 576, 512, 64, 1);
            } else {
                add(params + L.off[2 * i], cv[i].cin, c->wk[i], cv[i].cin, cv[i].cout, cv[i].cin, 0);
                if (i == 2) {
                    add(params + L.off[4], 64, c->wcat, 576, 64, 64, 1);
                    if (c->folded && !rag && c->fold6) add(params + L.off[4], 64, c->wcat6, 640, 64, 64, 1);
                }
                else add(params + L.off[2 * i], cv[i].cin, c->wt[i], cv[i].cout, cv[i].cout, cv[i].cin, 1);
            }
        }
        jobs.count = nj;
        // the accumulators of the forward pass are cleared by extra slices of the same launch
        int nf = 0;
        auto zero = [&](void* ptr, size_t bytes) { jobs.fill[nf++] = FillJob{ptr, bytes, 0u}; };
        zero(c->stats_f, c->stat_total * sizeof(double));
        zero(c->keys, static_cast<size_t>(c->B) * 1024 * sizeof(unsigned long long));
        if (ce) zero(ce, sizeof(pcseg_ce_accum));
        if (c->folded && !rag) zero(c->colsum[4], 128 * sizeof(double));
        if (c->folded && !rag && c->fold6) zero(c->colsum[6], static_cast<size_t>(c->B) * 64 * sizeof(double));
        jobs.fill_count = nf;
        StampScope ts(c, 88, s);
        pdl_launch(k_convert_multi, dim3(256, nj + nf), 256, 0, s, jobs);
        LAUNCH_OK("k_convert_multi");
    }
    const bool folded = c->folded && !rag;
    const bool fold6 = folded && c->fold6;
    const bool xf = c->xf && !rag;

    auto fin_args = [&](int i) {
        BnFinalizeArgs f;
        f.stats = c->stats_f + c->stat_off[i];
        f.gamma = params + L.off[20 + 2 * i];
        f.beta = params + L.off[21 + 2 * i];
        f.conv_bias = params + L.off[2 * i + 1];
        f.rmean = bnbuf + L.bn_off[i][0];
        f.rvar = bnbuf + L.bn_off[i][1];
        f.bnp = c->bnp[i];
        f.n = static_cast<double>(rag ? static_cast<long long>(c->B) * c->rag_N : c->P);
        f.inv_n = 1.0 / f.n;
        f.eps = BN_EPS;
        f.momentum = BN_MOMENTUM;
        f.C = cv[i].cout;
        return f;
    };
    // BN finalize of layer i is folded into this kernel (block 0 publishes bnp + running statistics)
    auto bn_relu = [&](int i, unsigned long long sd, unsigned int thr, float ks, double* colsum = nullptr, int rows_per_cloud = 0) -> int {
        const int co = cv[i].cout;
        StampScope ts(c, 89, s);
        int grid = strip_grid(P, co);
        if (rows_per_cloud > 0) {          // (clouds x blocks per cloud): strips never straddle clouds, colsum is per cloud
            const int clouds = static_cast<int>(P / rows_per_cloud);
            int bpc = grid / clouds;
            if (bpc < 1) bpc = 1;
            grid = clouds * bpc;
        }
        pdl_launch(k_bn_relu, grid, 256, 0, s, c->y[i], co, c->act[i], co, P, co, fin_args(i), sd, c->seed_ptr, thr, ks, colsum, rows_per_cloud);
        LAUNCH_OK("k_bn_relu");
        return 0;
    };

    // ragged: weight the representative pad rows by their multiplicity and drop the filler rows from the batch sums
    auto stats_fix = [&](int i) -> int {
        if (!rag) return 0;
        const int co = cv[i].cout;
        pdl_launch(k_stats_fix, dim3((co + 31) / 32, c->B), 256, 0, s, static_cast<const bf16*>(c->y[i]), co, co, rag_meta(c),
                   static_cast<const float*>(c->rowmult), c->stats_f + c->stat_off[i]);
        LAUNCH_OK("k_stats_fix");
        return 0;
    };

    {   // conv1 on CUDA cores
        int grid = static_cast<int>((P + 31) / 32);
        if (grid > num_sms() * 2) grid = num_sms() * 2;     // every block ends with 128 fp64 atomics on the same addresses
        pdl_launch(k_ingest<true>, grid, 256, 0, s, reinterpret_cast<const float4*>(x), static_cast<int>(P), params + L.off[0], nullptr, nullptr,
                                           c->y[0], c->stats_f + c->stat_off[0], 0);
        LAUNCH_OK("k_ingest");
        TRY(stats_fix(0));
        if (!xf) TRY(bn_relu(0, 0, 0, 1.f));
    }
    // transform-stage GEMM of layer i: BatchNorm + ReLU (+ dropout, column sums) of layer i - 1 happen on its A tiles
    auto xf_gemm = [&](int i, unsigned long long sd, unsigned int thr, float ks, double* colsum) -> int {
        GemmOp op = O.fwx[i];
        op.p.xf_fin = fin_args(i - 1);
        op.p.xf_seed = sd;
        op.p.seed_ptr = c->seed_ptr;
        op.p.xf_thr16 = thr;
        op.p.xf_keep_scale = ks;
        op.p.xf_colsum = colsum;
        return timed_gemm(c, op, i, s);
    };
    for (int i = 1; i <= 5; ++i) {
        if (i == 5) {
            GemmOp op = O.fw[5];
            op.p.gamma = params + L.off[20 + 2 * 5];
            TRY(timed_gemm(c, op, 5, s));
        } else if (i == 4 && folded) {
            // conv5: batch statistics predicted from the Gram matrix of a3, BN + ReLU applied in the GEMM epilogue
            TRY(timed_gemm(c, c->gram_op[4], 48 + 4, s));
            {
                StampScope ts(c, 80, s);
                pdl_launch(k_gram_reduce, 128 * 128 * 16 / 256, 256, 0, s, static_cast<const float*>(c->grampart4), c->gram_op[4].p.num_splits,
                           128 * 128, c->gramf[4], static_cast<const double*>(c->colsum[4]), static_cast<double>(c->P), 128);
                LAUNCH_OK("k_gram_reduce");
            }
            {
                StampScope ts(c, 81, s);
                pdl_launch(k_predict_bn<4>, 1024 / 8, 256, 0, s, static_cast<const float*>(c->gramf[4]), static_cast<const double*>(c->colsum[4]),
                           static_cast<const bf16*>(c->wk[4]), fin_args(4), c->stats_f + c->stat_off[4]);
                LAUNCH_OK("k_predict_bn");
            }
            TRY(timed_gemm(c, O.fw[4], 4, s));
            continue;
        } else if (xf && i <= 3) {
            // conv2 / conv3 / conv4 apply bn1 / bn2 / bn3 themselves; conv3 also takes the per-cloud column sums of point_feat
            TRY(xf_gemm(i, 0, 0, 1.f, (i == 2 && fold6) ? c->colsum[6] : nullptr));
        } else {
            TRY(timed_gemm(c, O.fw[i], i, s));
        }
        TRY(stats_fix(i));
        if (xf && i <= 2) continue;            // the next layer's transform stage applies this BatchNorm
        if (i == 1 && fold6) TRY(bn_relu(1, 0, 0, 1.f, c->colsum[6], c->N));      // per-cloud column sums of point_feat
        else if (i < 5) TRY(bn_relu(i, 0, 0, 1.f, (i == 3 && folded) ? c->colsum[4] : nullptr));
    }
    {   // global max-pool of relu(bn(y6)) with arg-index
        const int total = c->B * 1024;
        pdl_launch(k_maxpool_finish, (total + 255) / 256, 256, 0, s, c->keys, total, 1024, fin_args(5), c->gmax, c->ystar, c->argidx);
        LAUNCH_OK("k_maxpool_finish");
        const int warps = c->B * 512;
        pdl_launch(k_cloud_bias, (warps * 32 + 255) / 256, 256, 0, s, params + L.off[12] + 64, 1088, c->gmax, c->B, 512, 1024, nullptr, nullptr, c->cb);
        LAUNCH_OK("k_cloud_bias");
    }
    if (fold6) {
        // seg_conv1 with predicted statistics: per-cloud Gram matrices of point_feat, BN + ReLU + dropout in the GEMM epilogue
        TRY(timed_gemm(c, c->gram_op[6], 48 + 7, s));
        {
            StampScope ts(c, 80, s);
            pdl_launch(k_gram_reduce, dim3(64 * 64 * 16 / 256, c->B), 256, 0, s, static_cast<const float*>(c->grampart6),
                       c->gram_op[6].p.splits_per_group, 64 * 64, c->gramf[6], static_cast<const double*>(c->colsum[6]),
                       static_cast<double>(c->N), 64);
            LAUNCH_OK("k_gram_reduce");
        }
        {
            StampScope ts(c, 81, s);
            pdl_launch(k_predict_bn_cloud, dim3(512 / 8, c->B), 256, 0, s, static_cast<const float*>(c->gramf[6]),
                       static_cast<const double*>(c->colsum[6]), static_cast<const bf16*>(c->wk[6]), static_cast<const float*>(c->cb), c->B, c->N,
                       fin_args(6), c->stats_f + c->stat_off[6], c->part6, c->ticket6);
            LAUNCH_OK("k_predict_bn_cloud");
        }
        GemmOp op = O.fw[6];
        op.p.seed = seed + 1;
        op.p.seed_ptr = c->seed_ptr;
        op.p.drop_thr16 = c->thr16;
        op.p.keep_scale = c->keep_scale;
        TRY(timed_gemm(c, op, 6, s));
    } else {
        TRY(timed_gemm(c, O.fw[6], 6, s));
        TRY(stats_fix(6));
        TRY(bn_relu(6, seed + 1, c->thr16, c->keep_scale));
    }
    TRY(timed_gemm(c, O.fw[7], 7, s));
    TRY(stats_fix(7));
    if (xf && c->xf_seg3) {
        TRY(xf_gemm(8, seed + 2, c->thr16, c->keep_scale, nullptr));      // seg_conv3 applies bn_seg2 + ReLU + dropout itself
    } else {
        TRY(bn_relu(7, seed + 2, c->thr16, c->keep_scale));
        TRY(timed_gemm(c, O.fw[8], 8, s));
    }
    TRY(stats_fix(8));
    {
        int grid = static_cast<int>((P + 255) / 256);        // one thread per point
        if (grid > num_sms() * 4) grid = num_sms() * 4;
#define HEAD_FWD(NC_)                                                                                                          \
    case NC_:                                                                                                                  \
        pdl_launch(k_head_fwd<NC_, true>, grid, 256, 0, s, c->y[8], P, fin_args(8), params + L.off[18], params + L.off[19], c->C, logits, \
                   labels, class_w, reinterpret_cast<CeAccum*>(ce));                                                          \
        break;
        StampScope ts(c, 91, s);
        switch (c->C <= MAX_CLASSES ? c->C : (c->C <= 16 ? 16 : 32)) {       // class slots of the kernel
            HEAD_FWD(1) HEAD_FWD(2) HEAD_FWD(3) HEAD_FWD(4) HEAD_FWD(5) HEAD_FWD(6) HEAD_FWD(7) HEAD_FWD(8) HEAD_FWD(16) HEAD_FWD(32)
            default: return fail("pcseg_forward_train: unsupported num_classes %d", c->C);
        }
#undef HEAD_FWD
        LAUNCH_OK("k_head_fwd");
    }
    return 0;
}

extern "C" int pcseg_forward_train(pcseg_ctx* c, const float* x, const float* params, float* bnbuf, unsigned long long seed,
                                   float dropout_p, float* logits, const long long* labels, const float* class_w,
                                   pcseg_ce_accum* ce, const pcseg_step_state* state, void* stream) {
    g_pdl_call = !stream_is_capturing(stream);
    if (!c || !c->bound || !c->train) return fail("pcseg_forward_train: context not bound in train mode");
    if (!x || !params || !bnbuf || !logits) return fail("pcseg_forward_train: null tensor");
    if (dropout_p < 0.f || dropout_p >= 1.f) return fail("pcseg_forward_train: dropout_p out of range");
    if (labels && !ce) return fail("pcseg_forward_train: labels given without a CE accumulator");
    c->rag_active = false;
    return forward_train_rows(c, false, x, params, bnbuf, seed, dropout_p, logits, labels, class_w, ce, state,
                              static_cast<cudaStream_t>(stream));
}

extern "C" int pcseg_forward_train_ragged(pcseg_ctx* c, const float* x, const int* lengths, int nmax, const float* params, float* bnbuf,
                                          unsigned long long seed, float dropout_p, float* logits, const long long* labels,
                                          const float* class_w, pcseg_ce_accum* ce, const pcseg_step_state* state, void* stream) {
    g_pdl_call = true;
    if (!c || !c->bound || !c->train) return fail("pcseg_forward_train_ragged: context not bound in train mode");
    if (!x || !lengths || !params || !bnbuf || !logits) return fail("pcseg_forward_train_ragged: null argument");
    if (dropout_p < 0.f || dropout_p >= 1.f) return fail("pcseg_forward_train_ragged: dropout_p out of range");
    if (labels && !ce) return fail("pcseg_forward_train_ragged: labels given without a CE accumulator");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    TRY(rag_plan(c, lengths, nmax, s));
    TRY(rag_pack(c, x, labels, true, s));
    c->rag_active = true;
    TRY(forward_train_rows(c, true, c->xpack, params, bnbuf, seed, dropout_p, c->lpack, labels ? c->labpack : nullptr, class_w, ce, state, s));
    return rag_unpack_logits(c, logits, s);
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
// rows per strip of k_bn_bwd_apply: about 4 blocks per SM in total (each block ends with one atomic per column)
static int apply_rows_per_strip(int N, int B, int C) {
    const int rpp = 256 / (C / 8);
    const int unit = rpp * 4;                                    // rows consumed per loop iteration of a block
    static const int per_sm = env_int("PCSEG_APPLY_BLOCKS_PER_SM", 2);
    long long target_blocks = static_cast<long long>(per_sm > 0 ? per_sm : 2) * num_sms();      // == resident blocks (102 regs x 256 threads -> 2 per SM): one wave, half the atomics
    long long strips_per_cloud = (target_blocks + B - 1) / B;
    if (strips_per_cloud < 1) strips_per_cloud = 1;
    int rps = static_cast<int>((N + strips_per_cloud - 1) / strips_per_cloud);
    rps = ((rps + unit - 1) / unit) * unit;
    if (rps < unit) rps = unit;
    return rps;
}

extern "C" int pcseg_backward(pcseg_ctx* c, const float* x, const float* params, const float* dlogits, const float* logits,
                              const long long* labels, const float* class_w, const double* wsum_total, float* grads, int phase,
                              void* stream) {
    g_pdl_call = !stream_is_capturing(stream);
    if (!c || !c->bound || !c->train) return fail("pcseg_backward: context not bound in train mode");
    if (phase < 0 || phase > 2) return fail("pcseg_backward: phase must be 0, 1 or 2");
    if (!x || !params || !grads) return fail("pcseg_backward: null tensor");
    if (!dlogits && !(logits && labels && wsum_total)) return fail("pcseg_backward: need dlogits, or logits + labels + wsum_total");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const Layout& L = c->L;
    const ConvDef* cv = L.conv;
    const bool rag = c->rag_active;      // the forward this backward belongs to ran on the packed ragged batch
    pcseg_ctx::OpSet& O = c->ops[rag ? 1 : 0];
    const long long P = rag ? c->rag_rows : c->P;
    const int B = c->B;
    const int N = c->N;                   // rows per cloud (dense launches only: the ragged ones walk planned strips)
    const int* rag_strips = rag ? c->meta + 2 * B + 1 + c->rag_rows / 128 : nullptr;
    if (rag) {
        x = c->xpack;
        if (dlogits) {
            // caller's gradient of the padded logits -> packed rows (the packed logits are not needed in this mode)
            if (phase != 2) {
                const int copy_blocks = static_cast<int>((c->rag_rows + 255) / 256);
                pdl_launch(k_pack_dlogits, copy_blocks + B, 256, 0, s, dlogits, rag_meta(c), c->C, c->lpack);
                LAUNCH_OK("k_pack_dlogits");
            }
            dlogits = c->lpack;
            logits = nullptr;
            labels = nullptr;
        } else {
            logits = c->lpack;
            labels = c->labpack;
        }
    }

    if (phase != 2) {
        // every accumulator of the backward pass is cleared by ONE launch (one graph node instead of up to nine memset
        // nodes on the critical path); the buffers of the later layers are not touched by anything in between
        FillJobs fj;
        int nf = 0;
        auto zero = [&](void* ptr, size_t bytes, unsigned int value = 0u) { fj.fill[nf++] = FillJob{ptr, bytes, value}; };
        zero(grads, static_cast<size_t>(L.total) * sizeof(float));
        zero(c->stats_b, c->stat_total * sizeof(double));
        zero(c->dcb, static_cast<size_t>(B) * 512 * sizeof(float));
        if (c->folded && !rag) {
            if (c->fold6) zero(c->cloudsum6, (static_cast<size_t>(B) * 512 + 512 * 64) * sizeof(float));
            zero(c->cstf[5], 1024 * sizeof(float));       // (Q5 and the side rows are written, not accumulated)
            zero(c->rowslot5, static_cast<size_t>(c->P) * sizeof(int), 0x7f7f7f7fu);
            zero(c->gramf[5], 1024 * 1024 * sizeof(float));
            zero(c->qraw[4], FOLD4_ZERO_FLOATS * sizeof(float));
        }
        fj.count = nf;
        StampScope ts(c, 93, s);
        pdl_launch(k_fill_multi, dim3(128, nf), 256, 0, s, fj);
        LAUNCH_OK("k_fill_multi");
    }

    auto bwd_args = [&](int i) {
        BnBwdArgs b;
        b.stats = c->stats_b + c->stat_off[i];
        b.bnp = c->bnp[i];
        b.coef = c->coef[i];
        b.dgamma = grads + L.off[20 + 2 * i];
        b.dbeta = grads + L.off[21 + 2 * i];
        b.n = static_cast<double>(rag ? static_cast<long long>(c->B) * c->rag_N : c->P);
        b.inv_n = 1.0 / b.n;
        b.C = cv[i].cout;
        return b;
    };
    // (BN-backward coefficients are evaluated inside k_bn_bwd_apply: no launch of their own)
    auto apply = [&](int i, bf16* dy_out, int ld_dy, float* dcb) -> int {
        const int co = cv[i].cout;
        StampScope ts(c, 90, s);
        const int rps = apply_rows_per_strip(N, B, co);
        dim3 grid((N + rps - 1) / rps, B);
        if (rag)
            pdl_launch(k_bn_bwd_apply<false, true>, c->rag_strips, 256, 0, s, c->dz[i], co, c->y[i], co, dy_out, ld_dy, N, co, rps, bwd_args(i),
                       grads + L.off[2 * i + 1], dcb, nullptr, nullptr, rag_strips, c->rowmult);
        else
            pdl_launch(k_bn_bwd_apply<false, false>, grid, 256, 0, s, c->dz[i], co, c->y[i], co, dy_out, ld_dy, N, co, rps, bwd_args(i),
                       grads + L.off[2 * i + 1], dcb, nullptr, nullptr, nullptr, nullptr);
        LAUNCH_OK("k_bn_bwd_apply");
        return 0;
    };
    auto wgrad = [&](int i, float* dst, int ldc) -> int {
        GemmOp op = O.wg_op[i];
        op.p.out_f32 = dst;
        op.p.ldc = ldc;
        return timed_gemm(c, op, 32 + i, s);
    };
    auto dgrad = [&](int i, unsigned long long sd, unsigned int thr, float ks) -> int {
        GemmOp op = O.dg[i];
        // (storing the keep masks in forward and re-loading them here was measured slower than regenerating Philox)
        op.p.seed = sd;
        op.p.seed_ptr = c->seed_ptr;
        op.p.drop_thr16 = thr;
        op.p.keep_scale = ks;
        return timed_gemm(c, op, 16 + i, s);
    };

    if (phase != 2) {
    if (c->C > MAX_CLASSES) {   // seg_conv4 + loss gradient, 9..32 classes: per-point kernel + column reductions
        int grid = static_cast<int>((P + 255) / 256);
        if (grid > num_sms() * 4) grid = num_sms() * 4;
        const float* dl_src = dlogits ? dlogits : c->dlbuf;
        long long rgrid = (P + 63) / 64;
        if (rgrid > 2LL * num_sms()) rgrid = 2LL * num_sms();
#define HEAD_BWD_WIDE(NC_)                                                                                                   \
        pdl_launch(k_head_bwd_wide_points<NC_>, grid, 256, 0, s, c->y[8], P, c->bnp[8], params + L.off[18], c->C, dlogits, logits, labels,  \
                   class_w, wsum_total, c->dlbuf, c->dz[8]);                                                                 \
        LAUNCH_OK("k_head_bwd_wide_points");                                                                                 \
        pdl_launch(k_head_bwd_wide_reduce<NC_>, static_cast<int>(rgrid), 256, 0, s, c->y[8], c->dz[8], P, c->bnp[8], c->C, dl_src,           \
                   grads + L.off[18], grads + L.off[19], c->stats_b + c->stat_off[8]);                                      \
        LAUNCH_OK("k_head_bwd_wide_reduce");
        if (c->C <= 16) { HEAD_BWD_WIDE(16) } else { HEAD_BWD_WIDE(32) }
#undef HEAD_BWD_WIDE
    } else {   // seg_conv4 + loss gradient
        int grid = static_cast<int>((P + 31) / 32);
        if (grid > num_sms() * 2) grid = num_sms() * 2;
#define HEAD_BWD(NC_)                                                                                                         \
    case NC_:                                                                                                                 \
        pdl_launch(k_head_bwd<NC_>, grid, 256, 0, s, c->y[8], P, c->bnp[8], params + L.off[18], dlogits, logits, labels, class_w, wsum_total, \
                                             c->dz[8], grads + L.off[18], grads + L.off[19], c->stats_b + c->stat_off[8]);  \
        break;
        StampScope ts(c, 92, s);
        switch (c->C) {
            HEAD_BWD(1) HEAD_BWD(2) HEAD_BWD(3) HEAD_BWD(4) HEAD_BWD(5) HEAD_BWD(6) HEAD_BWD(7) HEAD_BWD(8)
            default: return fail("pcseg_backward: unsupported num_classes %d", c->C);
        }
#undef HEAD_BWD
        LAUNCH_OK("k_head_bwd");
    }
    // seg_conv3
    TRY(apply(8, c->dy[8], 128, nullptr));
    TRY(wgrad(8, grads + L.off[16], 256));
    TRY(dgrad(8, c->seed + 2, c->thr16, c->keep_scale));
    // seg_conv2
    TRY(apply(7, c->dy[7], 256, nullptr));
    TRY(wgrad(7, grads + L.off[14], 512));
    TRY(dgrad(7, c->seed + 1, c->thr16, c->keep_scale));
    // seg_conv1: point-feature columns by GEMM, global columns per cloud
    if (c->folded && !rag && c->fold6) {
        // folded: Q6 = dz6^T a1, coefficients (incl. the per-cloud gradient dcb of the pooled branch), dW / data-gradient
        // weights / per-cloud constant rows; dz6 itself already sits in dycat[:, 64:576]
        TRY(timed_gemm(c, O.wg_op[6], 32 + 6, s));
        Fold6Args f6;
        f6.Q = c->qraw[6];
        f6.W = c->wk[6];
        f6.G = c->gramf[6];
        f6.s = c->colsum[6];
        f6.part = c->part6;
        f6.cb = c->cb;
        f6.S1 = c->cloudsum6;
        f6.bnp = c->bnp[6];
        f6.coef = c->coef[6];
        f6.dgamma = grads + L.off[20 + 2 * 6];
        f6.dbeta = grads + L.off[21 + 2 * 6];
        f6.dbias = grads + L.off[2 * 6 + 1];
        f6.dcb = c->dcb;
        f6.gsum = c->gsum6;
        f6.ssum = c->ssum6;
        f6.dW = grads + L.off[12];
        f6.ld_dw = 1088;
        f6.wcat = c->wcat6;
        f6.ld_wcat = 640;
        f6.col0 = 64;
        f6.cst = c->cst6;
        f6.n = static_cast<double>(c->P);
        f6.N = c->N;
        f6.clouds = B;
        f6.Co = 512;
        f6.Ci = 64;
        {
            StampScope ts(c, 86, s);
            pdl_launch(k_fold6_coef, fold6_coef_blocks(512, 64), 256, 0, s, f6);
            LAUNCH_OK("k_fold6_coef");
        }
        {
            StampScope ts(c, 87, s);
            pdl_launch(k_fold6_bwd, fold6_bwd_blocks(512, 64, B), 256, 0, s, f6);
            LAUNCH_OK("k_fold6_bwd");
        }
    } else {
        TRY(apply(6, c->dycat + 64, 576, c->dcb));
        TRY(wgrad(6, grads + L.off[12], 1088));
    }
    {
        dim3 grid_dg(1024 / 32, B);
        pdl_launch(k_cloud_bwd_dg, grid_dg, 256, 0, s, c->dcb, params + L.off[12] + 64, 1088, B, 512, 1024, c->gmax, c->ystar, c->bnp[5], c->dzv,
                                              c->stats_b + c->stat_off[5]);
        LAUNCH_OK("k_cloud_bwd_dg");
        dim3 grid_dw(1024 / 256, 512);
        pdl_launch(k_cloud_bwd_dw, grid_dw, 256, 0, s, c->dcb, c->gmax, B, 512, 1024, grads + L.off[12] + 64, 1088);
        LAUNCH_OK("k_cloud_bwd_dw");
    }
    // global_feat
    if (c->folded && !rag) {
        // folded BatchNorm backward (neither y5 nor dy5 exist): coefficients from {sum dzv, sum dzv*yhat}, S5 on the tensor
        // cores, the max-pool gradient rows through the side buffer, ONE data-gradient GEMM over a4; then the Gram matrix of
        // a4 and the weight gradient in the epilogue of the W5 Gc4 GEMM
        Fold5Args f5;
        f5.stats_b = c->stats_b + c->stat_off[5];
        f5.bnp = c->bnp[5];
        f5.coef = c->coef[5];
        f5.dgamma = grads + L.off[20 + 2 * 5];
        f5.dbeta = grads + L.off[21 + 2 * 5];
        f5.dbias = grads + L.off[2 * 5 + 1];
        f5.W = c->wk[5];
        f5.WB = c->wb5;
        f5.cst = c->cstf[5];
        f5.n = static_cast<double>(c->P);
        f5.Co = 1024;
        f5.Ci = 1024;
        {
            StampScope ts(c, 82, s);
            pdl_launch(k_fold5_prep, 1024 / 8, 256, 0, s, f5);
            LAUNCH_OK("k_fold5_prep");
        }
        TRY(timed_gemm(c, c->s5_op, 48 + 6, s));
        {
            StampScope ts(c, 83, s);
            pdl_launch(k_pool_claim, (B * 1024 + 255) / 256, 256, 0, s, static_cast<const float*>(c->dzv), static_cast<const int*>(c->argidx), B * 1024,
                       1024, N, c->rowslot5);
            LAUNCH_OK("k_pool_claim");
        }
        {
            StampScope ts(c, 84, s);
            pdl_launch(k_pool_rows_own, 1024, 128, 0, s, static_cast<const float*>(c->dzv), static_cast<const int*>(c->argidx), B, 1024, N,
                       static_cast<const int*>(c->rowslot5), static_cast<const float4*>(c->coef[5]), static_cast<const bf16*>(c->wk[5]),
                       static_cast<const bf16*>(c->act[4]), c->side5, c->qraw[5]);
            LAUNCH_OK("k_pool_rows_own");
            pdl_launch(k_pool_rows_add, B * 1024, 128, 0, s, static_cast<const float*>(c->dzv), static_cast<const int*>(c->argidx), 1024, N,
                       static_cast<const int*>(c->rowslot5), static_cast<const float4*>(c->coef[5]), static_cast<const bf16*>(c->wk[5]),
                       c->side5);
            LAUNCH_OK("k_pool_rows_add");
        }
        TRY(dgrad(5, 0, 0, 1.f));
        TRY(timed_gemm(c, c->gram_op[5], 48 + 5, s));
        {
            StampScope ts(c, 85, s);
            pdl_launch(k_gram_center, 1024 * 1024 / 256, 256, 0, s, static_cast<const float*>(c->gramf[5]),
                       static_cast<const double*>(c->stats_b + c->stat_off[4] + 1024), static_cast<double>(c->P), 1024, c->gc5b);
            LAUNCH_OK("k_gram_center");
        }
        GemmOp t5 = c->t5_op;
        t5.p.out_f32 = grads + L.off[10];
        TRY(timed_gemm(c, t5, 32 + 5, s));
    } else {
        const int rps = apply_rows_per_strip(N, B, 1024);
        dim3 grid((N + rps - 1) / rps, B);
        if (rag)
            pdl_launch(k_bn_bwd_apply<true, true>, c->rag_strips, 256, 0, s, nullptr, 0, c->y[5], 1024, c->dy[5], 1024, N, 1024, rps, bwd_args(5),
                       grads + L.off[11], nullptr, c->argidx, c->dzv, rag_strips, c->rowmult);
        else
            pdl_launch(k_bn_bwd_apply<true, false>, grid, 256, 0, s, nullptr, 0, c->y[5], 1024, c->dy[5], 1024, N, 1024, rps, bwd_args(5),
                       grads + L.off[11], nullptr, c->argidx, c->dzv, nullptr, nullptr);
        LAUNCH_OK("k_bn_bwd_apply<sparse>");
        TRY(wgrad(5, grads + L.off[10], 1024));
        TRY(dgrad(5, 0, 0, 1.f));
    }
    }   // phase != 2
    if (phase == 1) return 0;
    // conv5
    if (c->folded && !rag) {
        // folded BatchNorm backward: neither y4 nor dy4 exist.  Q = dz4^T a3, per-channel coefficients, dW / S / const on
        // CUDA cores (K = 128), then ONE data-gradient GEMM over [dz4 | a3]
        TRY(timed_gemm(c, O.wg_op[4], 32 + 4, s));
        FoldArgs f;
        f.Q = c->qraw[4];
        f.W = c->wk[4];
        f.Wt = c->wt[4];
        f.Gc = c->gramf[4];
        f.s = c->colsum[4];
        f.sum_dz = c->stats_b + c->stat_off[4];
        f.bnp = c->bnp[4];
        f.coef = c->coef[4];
        f.dgamma = grads + L.off[20 + 2 * 4];
        f.dbeta = grads + L.off[21 + 2 * 4];
        f.dbias = grads + L.off[2 * 4 + 1];
        f.dW = grads + L.off[8];
        f.ld_dw = 128;
        f.Bw = c->bwf[4];
        f.ld_bw = 1024 + 128;
        f.cst = c->cstf[4];
        f.n = static_cast<double>(c->P);
        f.Co = 1024;
        f.Ci = 128;
        {
            StampScope ts(c, 86, s);
            pdl_launch(k_fold_coef, fold_coef_blocks(f.Co, f.Ci), 256, 0, s, f);
            LAUNCH_OK("k_fold_coef");
        }
        {
            StampScope ts(c, 87, s);
            pdl_launch(k_fold_bwd, fold_bwd_blocks(f.Co, f.Ci), 256, 0, s, f);
            LAUNCH_OK("k_fold_bwd");
        }
        TRY(timed_gemm(c, O.dg[4], 16 + 4, s));
    } else {
        TRY(apply(4, c->dy[4], 1024, nullptr));
        TRY(wgrad(4, grads + L.off[8], 128));
        TRY(dgrad(4, 0, 0, 1.f));
    }
    // conv4
    TRY(apply(3, c->dy[3], 128, nullptr));
    TRY(wgrad(3, grads + L.off[6], 64));
    TRY(dgrad(3, 0, 0, 1.f));
    // conv3 (its dy lives in the left 64 columns of dycat), then the skip join into conv2's output
    TRY(apply(2, c->dycat, 576, nullptr));
    TRY(wgrad(2, grads + L.off[4], 64));
    TRY(dgrad(2, 0, 0, 1.f));
    // conv2
    TRY(apply(1, c->dy[1], 64, nullptr));
    TRY(wgrad(1, grads + L.off[2], 64));
    TRY(dgrad(1, 0, 0, 1.f));
    // conv1 (no data gradient: the input does not require grad)
    TRY(apply(0, c->dy[0], 64, nullptr));
    {
        int grid = static_cast<int>((P + 31) / 32);
        if (grid > num_sms() * 2) grid = num_sms() * 2;
        pdl_launch(k_ingest_bwd, grid, 256, 0, s, c->dy[0], reinterpret_cast<const float4*>(x), P, grads + L.off[0]);
        LAUNCH_OK("k_ingest_bwd");
    }
    return 0;
}

extern "C" int pcseg_adam_step(float* params, const float* grads, float* m, float* v, long long n, int step, float lr, float b1,
                               float b2, float eps, float wd, float grad_scale, const pcseg_step_state* state, const double* grad_div,
                               void* stream) {
    g_pdl_call = !stream_is_capturing(stream);
    if (!params || !grads || !m || !v || n <= 0 || (step < 1 && !state)) return fail("pcseg_adam_step: bad arguments");
    if (step < 1) step = 1;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const float bc1 = 1.f - powf(b1, static_cast<float>(step));
    const float bc2 = 1.f - powf(b2, static_cast<float>(step));
    pdl_launch(k_adam, ew_grid(n), 256, 0, s, params, grads, m, v, n, lr, b1, b2, eps, wd, bc1, sqrtf(bc2), grad_scale,
                                      reinterpret_cast<const StepState*>(state), grad_div);
    LAUNCH_OK("k_adam");
    return 0;
}

extern "C" int pcseg_eval_metrics(const float* logits, const long long* labels, long long P, int C, const float* class_w,
                                  pcseg_ce_accum* ce, unsigned long long* confusion, long long* pred_out, void* stream) {
    g_pdl_call = true;
    if (!logits || P <= 0 || C < 1 || C > API_MAX_CLASSES) return fail("pcseg_eval_metrics: bad arguments");
    if (!labels && !pred_out) return fail("pcseg_eval_metrics: nothing to compute (no labels, no pred_out)");
    int grid = static_cast<int>((P + 255) / 256);
    if (grid > num_sms() * 8) grid = num_sms() * 8;
    if (C <= MAX_CLASSES)
        pdl_launch(k_eval_metrics<MAX_CLASSES>, grid, 256, 0, static_cast<cudaStream_t>(stream), logits, labels, P, C, class_w,
                   reinterpret_cast<CeAccum*>(ce), confusion, pred_out);
    else
        pdl_launch(k_eval_metrics<API_MAX_CLASSES>, grid, 256, 0, static_cast<cudaStream_t>(stream), logits, labels, P, C, class_w,
                   reinterpret_cast<CeAccum*>(ce), confusion, pred_out);
    LAUNCH_OK("k_eval_metrics");
    return 0;
}

extern "C" int pcseg_step_advance(pcseg_step_state* state, float b1, float b2, void* stream) {
    g_pdl_call = !stream_is_capturing(stream);
    if (!state) return fail("pcseg_step_advance: null state");
    pdl_launch(k_step_advance, 1, 1, 0, static_cast<cudaStream_t>(stream), reinterpret_cast<StepState*>(state), b1, b2);
    LAUNCH_OK("k_step_advance");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// gradient all-reduce over NVLink peer memory (peer_allreduce.cuh)
// ------------------------------------------------------------------------------------------------
struct pcseg_peer_ar {
    PeerArArgs a;
    uint32_t* local = nullptr;       // {arrive counter, generation, epoch, -} in local device memory
    void* opened[2 * AR_MAX_RANKS] = {};
    int num_opened = 0;
};

typedef CUresult (*GetAddressRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);

extern "C" int pcseg_ipc_export(const void* ptr, unsigned char* handle_out /* 64 bytes */, long long* offset_out) {
    if (!ptr || !handle_out || !offset_out) return fail("pcseg_ipc_export: null argument");
    static GetAddressRangeFn get_range = nullptr;
    if (!get_range) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn)
            return fail("pcseg_ipc_export: cuMemGetAddressRange unavailable");
        get_range = reinterpret_cast<GetAddressRangeFn>(fn);
    }
    CUdeviceptr base = 0;
    size_t size = 0;
    if (get_range(&base, &size, reinterpret_cast<CUdeviceptr>(ptr)) != CUDA_SUCCESS) return fail("pcseg_ipc_export: cuMemGetAddressRange failed");
    cudaIpcMemHandle_t h;
    CUDA_OK(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)));
    static_assert(sizeof(h) == 64, "IPC handle size");
    memcpy(handle_out, &h, 64);
    *offset_out = static_cast<long long>(reinterpret_cast<CUdeviceptr>(ptr) - base);
    return 0;
}

// bytes of the IPC-shared block of one rank: signals + publish buffer (one slice of the arena)
extern "C" long long pcseg_peer_ar_signal_bytes(long long n_floats, int world) {
    const long long n4 = n_floats / 4, per = (n4 + world - 1) / world;
    return static_cast<long long>((sizeof(PeerSignals) + 255) & ~size_t(255)) + per * 16 + 256;
}

extern "C" int pcseg_peer_ar_create(pcseg_peer_ar** out, int rank, int world, float* arena, long long n_floats, void* signals,
                                    void* counters /* 4 x uint32, zeroed */, const double* lw_in, double* lw_out) {
    // signals = [PeerSignals | publish buffer of ceil(n_floats / 4 / world) float4]
    if (!out || !arena || !signals || !counters) return fail("pcseg_peer_ar_create: null argument");
    if (world < 2 || world > AR_MAX_RANKS || (world != 2 && world != 4 && world != 8)) return fail("pcseg_peer_ar_create: world size %d unsupported (2, 4, 8)", world);
    if (rank < 0 || rank >= world) return fail("pcseg_peer_ar_create: bad rank");
    if (n_floats % 4 != 0 || (reinterpret_cast<uintptr_t>(arena) & 15)) return fail("pcseg_peer_ar_create: the arena must be 16-byte aligned and a multiple of 4 floats long");
    pcseg_peer_ar* h = new pcseg_peer_ar();
    memset(&h->a, 0, sizeof(h->a));
    h->a.rank = rank;
    h->a.world = world;
    h->a.n = n_floats;
    h->a.arena[rank] = arena;
    h->a.sig[rank] = static_cast<PeerSignals*>(signals);
    h->a.pub[rank] = reinterpret_cast<float*>(static_cast<char*>(signals) + ((sizeof(PeerSignals) + 255) & ~size_t(255)));
    h->local = static_cast<uint32_t*>(counters);
    h->a.epoch = h->local + 2;
    h->a.lw_in = lw_in;
    h->a.lw_out = lw_out;
    *out = h;
    return 0;
}

extern "C" int pcseg_peer_ar_open(pcseg_peer_ar* h, int peer, const unsigned char* arena_handle, long long arena_offset,
                                  const unsigned char* sig_handle, long long sig_offset) {
    if (!h || peer < 0 || peer >= h->a.world || peer == h->a.rank) return fail("pcseg_peer_ar_open: bad peer");
    cudaIpcMemHandle_t ha, hs;
    memcpy(&ha, arena_handle, 64);
    memcpy(&hs, sig_handle, 64);
    void* pa = nullptr;
    void* ps = nullptr;
    CUDA_OK(cudaIpcOpenMemHandle(&pa, ha, cudaIpcMemLazyEnablePeerAccess));
    h->opened[h->num_opened++] = pa;
    if (memcmp(&ha, &hs, 64) == 0) {
        ps = pa;                       // same allocation block (caching allocator): one mapping
    } else {
        CUDA_OK(cudaIpcOpenMemHandle(&ps, hs, cudaIpcMemLazyEnablePeerAccess));
        h->opened[h->num_opened++] = ps;
    }
    h->a.arena[peer] = reinterpret_cast<float*>(static_cast<char*>(pa) + arena_offset);
    h->a.sig[peer] = reinterpret_cast<PeerSignals*>(static_cast<char*>(ps) + sig_offset);
    h->a.pub[peer] = reinterpret_cast<float*>(static_cast<char*>(ps) + sig_offset + ((sizeof(PeerSignals) + 255) & ~size_t(255)));
    return 0;
}

extern "C" int pcseg_peer_ar_run(pcseg_peer_ar* h, void* stream) {
    if (!h) return fail("pcseg_peer_ar_run: null handle");
    for (int p = 0; p < h->a.world; ++p)
        if (!h->a.arena[p] || !h->a.sig[p]) return fail("pcseg_peer_ar_run: peer %d not opened", p);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int grid = 128;              // constant: the intra-rank barrier counts arrivals per launch (all CTAs must be co-resident)
    switch (h->a.world) {
        case 2: k_peer_allreduce<2><<<grid, 512, 0, s>>>(h->a, h->local); break;
        case 4: k_peer_allreduce<4><<<grid, 512, 0, s>>>(h->a, h->local); break;
        default: k_peer_allreduce<8><<<grid, 512, 0, s>>>(h->a, h->local); break;
    }
    LAUNCH_OK("k_peer_allreduce");
    return 0;
}

extern "C" int pcseg_peer_ar_destroy(pcseg_peer_ar* h) {
    if (!h) return 0;
    for (int i = 0; i < h->num_opened; ++i) cudaIpcCloseMemHandle(h->opened[i]);
    delete h;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// stand-alone GEMM for unit tests
// ------------------------------------------------------------------------------------------------
extern "C" int pcseg_gemm_test(int layout, int M, int N, int K, const void* A, int lda, const void* Bm, int ldb, void* D, int ldc,
                               const float* bias, int block_n, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    GemmOp op;
    if (layout == 0) {
        TRY(setup_gemm_kmajor(&op, EPI_BIAS_RELU, A, lda, Bm, ldb, M, N, K, D, ldc, nullptr, 0));
        if (block_n) {
            if (N % block_n) return fail("pcseg_gemm_test: N not a multiple of block_n");
            op.bn = block_n;
            op.p.num_n_tiles = N / block_n;
            TRY(make_tmap(&op.tmB, Bm, K, N, ldb, 64, block_n));
        }
        op.p.bias = bias;
        if (!bias) return fail("pcseg_gemm_test: layout 0 needs a bias vector");
    } else if (layout == 1) {
        TRY(setup_gemm_wgrad(&op, A, lda, M, Bm, ldb, N, K, static_cast<float*>(D), ldc));
        if (block_n) {
            if (N % block_n) return fail("pcseg_gemm_test: N not a multiple of block_n");
            op.bn = block_n;
            op.p.num_n_tiles = N / block_n;
        }
    } else {
        return fail("pcseg_gemm_test: unknown layout %d", layout);
    }
    return launch_gemm(op, s);
}

// ------------------------------------------------------------------------------------------------
// inspection hook for the layer-wise parity tests
// ------------------------------------------------------------------------------------------------
extern "C" int pcseg_debug_copy(pcseg_ctx* c, int kind, int layer, void* dst, long long dst_bytes, long long* rows, long long* cols,
                                int* elem_bytes, void* stream) {
    if (!c || !c->bound || !c->train) return fail("pcseg_debug_copy: context not bound in train mode");
    if ((kind <= 7 || kind >= 14) && (layer < 0 || layer >= NUM_BN)) return fail("pcseg_debug_copy: bad layer %d", layer);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const ConvDef* cv = c->L.conv;
    const long long P = c->P, B = c->B;
    const void* src = nullptr;
    long long r = 0, cc = 0, src_pitch = 0;
    int eb = 0;
    switch (kind) {
        case 0: src = c->y[layer]; r = P; cc = cv[layer].cout; eb = 2; break;
        case 1: src = c->act[layer]; r = P; cc = cv[layer].cout; eb = 2; break;
        case 2: src = c->dz[layer]; r = P; cc = cv[layer].cout; eb = 2; break;
        case 3:
            r = P; cc = cv[layer].cout; eb = 2;
            if (layer == 2) { src = c->dycat; src_pitch = 576 * 2; }
            else if (layer == 6) { src = c->dycat + 64; src_pitch = 576 * 2; }
            else src = c->dy[layer];
            break;
        case 4: src = c->bnp[layer]; r = cv[layer].cout; cc = 4; eb = 4; break;
        case 5: src = c->coef[layer]; r = cv[layer].cout; cc = 4; eb = 4; break;
        case 6: src = c->stats_f + c->stat_off[layer]; r = 2; cc = cv[layer].cout; eb = 8; break;
        case 7: src = c->stats_b + c->stat_off[layer]; r = 2; cc = cv[layer].cout; eb = 8; break;
        case 8: src = c->gmax; r = B; cc = 1024; eb = 4; break;
        case 9: src = c->ystar; r = B; cc = 1024; eb = 4; break;
        case 10: src = c->argidx; r = B; cc = 1024; eb = 4; break;
        case 11: src = c->cb; r = B; cc = 512; eb = 4; break;
        case 12: src = c->dcb; r = B; cc = 512; eb = 4; break;
        case 13: src = c->dzv; r = B; cc = 1024; eb = 4; break;
        // folded layers (layer = conv index of the layer whose y / dy are not materialised)
        case 14: src = c->gramf[layer]; r = cv[layer].cin; cc = cv[layer].cin; eb = 4; break;
        case 15: src = c->colsum[layer]; r = 1; cc = cv[layer].cin; eb = 8; break;
        case 16: src = c->qraw[layer]; r = cv[layer].cout; cc = (layer == 6) ? 64 : cv[layer].cin; eb = 4; break;
        case 17: src = c->bwf[layer]; r = cv[layer].cin; cc = cv[layer].cout + cv[layer].cin; eb = 2; break;
        case 18: src = c->cstf[layer]; r = 1; cc = cv[layer].cin; eb = 4; break;
        case 19: src = c->s5b; r = 1024; cc = 1024; eb = 2; break;
        case 20: src = c->gc5b; r = 1024; cc = 1024; eb = 2; break;
        case 21: src = c->side5; r = B * 1024; cc = 1024; eb = 4; break;
        case 22: src = c->rowslot5; r = 1; cc = P; eb = 4; break;
        case 23: src = c->cloudsum6; r = B; cc = 512; eb = 4; break;
        case 24: src = c->cst6; r = B; cc = 64; eb = 4; break;
        case 25: src = c->wcat6; r = 64; cc = 640; eb = 2; break;
        case 26: src = c->fold6 ? c->gramf[6] : nullptr; r = B * 64; cc = 64; eb = 4; break;
        case 27: src = c->fold6 ? c->colsum[6] : nullptr; r = B; cc = 64; eb = 8; break;
        default: return fail("pcseg_debug_copy: unknown kind %d", kind);
    }
    if (!src) return fail("pcseg_debug_copy: tensor kind %d layer %d is not materialised", kind, layer);
    if (rows) *rows = r;
    if (cols) *cols = cc;
    if (elem_bytes) *elem_bytes = eb;
    const long long bytes = r * cc * eb;
    if (!dst) return 0;      // size query
    if (dst_bytes < bytes) return fail("pcseg_debug_copy: destination too small (%lld < %lld)", dst_bytes, bytes);
    if (src_pitch) CUDA_OK(cudaMemcpy2DAsync(dst, cc * eb, src, src_pitch, cc * eb, r, cudaMemcpyDeviceToDevice, s));
    else CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// per-kernel event timing
// ------------------------------------------------------------------------------------------------
extern "C" int pcseg_profile_reset(pcseg_ctx* c) {
    if (!c) return fail("pcseg_profile_reset: null ctx");
    for (auto& st : c->stamps) { cudaEventDestroy(st.a); cudaEventDestroy(st.b); }
    c->stamps.clear();
    return 0;
}
extern "C" int pcseg_profile_enable(pcseg_ctx* c, int on) {
    if (!c) return fail("pcseg_profile_enable: null ctx");
    c->profiling = on != 0;
    return 0;
}
extern "C" int pcseg_profile_read(pcseg_ctx* c, int tag, double* total_ms, long long* launches) {
    if (!c) return fail("pcseg_profile_read: null ctx");
    CUDA_OK(cudaDeviceSynchronize());
    double tot = 0.0;
    long long n = 0;
    for (auto& st : c->stamps) {
        if (st.tag != tag) continue;
        float ms = 0.f;
        CUDA_OK(cudaEventElapsedTime(&ms, st.a, st.b));
        tot += ms;
        ++n;
    }
    if (total_ms) *total_ms = tot;
    if (launches) *launches = n;
    return 0;
}
