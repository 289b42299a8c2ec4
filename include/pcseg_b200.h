/* pcseg_b200 — C ABI of the B200-native point-cloud segmentation hot path.
 *
 * The reference has no FFI: its boundary is the Python class
 * PointNetSegmentation (point_cloud_segmentation.py:65-133, "pcs.py" below) plus the
 * loss / backward / optimizer calls of its training loop (pcs.py:241-255).  The entry points
 * below are what a binding for that path binds; every one cites the reference lines it
 * replaces.  All pointers except `ctx`, `out` arrays documented as host, and strings are
 * DEVICE pointers; `stream` is a cudaStream_t (NULL = legacy default stream).  Functions
 * return 0 on success, non-zero on error (message via pcseg_last_error()); no exceptions
 * cross this boundary and the library never allocates device memory itself: the caller
 * passes one workspace of pcseg_workspace_bytes().
 *
 * Layouts: x (B, N, 4) fp32 contiguous; logits (B, N, C) fp32 contiguous; labels (B, N) int64,
 * -1 = padding (pcs.py:54); parameters = one flat fp32 arena in state_dict order
 * (pcs.py:70-94: conv1.weight, conv1.bias, ..., seg_conv4.bias, bn1.weight, bn1.bias, ...,
 * bn_seg3.bias), BatchNorm running statistics = one flat fp32 arena
 * (bn1.running_mean, bn1.running_var, ..., bn_seg3.running_var).
 */
#ifndef PCSEG_B200_H
#define PCSEG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pcseg_ctx pcseg_ctx;

#define PCSEG_MAX_CLASSES 32          /* 1..8 classes run the fused head kernels, 9..32 the wide ones */
#define PCSEG_NUM_PARAM_TENSORS 38   /* 10 conv weight/bias pairs + 9 BN weight/bias pairs */
#define PCSEG_NUM_BN 9

/* Cross-entropy accumulators written by pcseg_forward_train (device struct, 32 bytes). */
typedef struct pcseg_ce_accum {
    double loss_num;             /* sum_i w[y_i] * -log softmax(z_i)[y_i] over labels != -1 */
    double w_sum;                /* sum_i w[y_i]                                            */
    unsigned long long correct;  /* argmax == label                                         */
    unsigned long long valid;    /* labels != -1                                            */
} pcseg_ce_accum;

/* Device-resident per-step state (32 bytes).  Lets a whole training step be captured in a CUDA graph: the dropout
 * seed, the Adam step count / bias corrections and the learning rate are read from device memory by the kernels
 * instead of being baked into launch arguments.  pcseg_step_advance moves it to the next step. */
typedef struct pcseg_step_state {
    unsigned long long seed;     /* added to the `seed` argument of pcseg_forward_train                     */
    long long step;              /* optimizer step count (1 after the first pcseg_step_advance)             */
    float lr;                    /* learning rate; the host may rewrite it between steps (StepLR, pcs.py:218) */
    float bias_corr1;            /* 1 - beta1^step                                                          */
    float bias_corr2_sqrt;       /* sqrt(1 - beta2^step)                                                    */
    float reserved;
} pcseg_step_state;

const char* pcseg_last_error(void);
const char* pcseg_version(void);

/* Model construction: replaces PointNetSegmentation.__init__(num_classes, input_dim=4), pcs.py:66-96. */
int pcseg_create(pcseg_ctx** out, int num_classes);
int pcseg_destroy(pcseg_ctx* ctx);

/* Flat-arena layout queries (host only, no GPU needed). */
long long pcseg_param_count(int num_classes);                 /* 1 927 621 for C = 5 */
long long pcseg_param_offset(int num_classes, int tensor);    /* tensor in [0, 38): state_dict order of parameters */
long long pcseg_param_numel(int num_classes, int tensor);
long long pcseg_bn_buffer_count(void);                        /* 6 528 floats */
long long pcseg_bn_buffer_offset(int bn, int which);          /* which: 0 running_mean, 1 running_var */
long long pcseg_workspace_bytes(int B, int N, int num_classes, int train);

/* Bind a batch shape and a caller-owned workspace (builds TMA descriptors; host only work). */
int pcseg_bind(pcseg_ctx* ctx, int B, int N, void* workspace, long long workspace_bytes, int train);

/* model.eval() weight preparation: folds BatchNorm running statistics into bf16 weights
 * (pcs.py:277/432 semantics).  Call again whenever parameters or buffers change. */
int pcseg_prepare_eval(pcseg_ctx* ctx, const float* params, const float* bn_buffers, void* stream);

/* Inference forward: PointNetSegmentation.forward under eval()/no_grad, pcs.py:98-133, 450-451.
 * labels_out (optional, int64 (B,N)) receives argmax over classes, pcs.py:452. */
int pcseg_forward_eval(pcseg_ctx* ctx, const float* x, float* logits, long long* labels_out, void* stream);

/* Point-sharded inference of clouds that are split over several GPUs ("within one cloud"): every rank
 * holds the same B clouds but its own slice of their points.  part 1 runs pcs.py:103-114 on the local points and leaves
 * the local max-pool result in the context; the caller reduces it over the ranks with MAX (e.g. ncclAllReduce on the
 * pointer returned by pcseg_pooled_feature: B x 1024 fp32, all values >= 0) and calls part 2 (pcs.py:117-131 + argmax)
 * for the local points.  The logits of a point are bit-identical to those of the un-sharded cloud. */
int pcseg_forward_eval_part(pcseg_ctx* ctx, const float* x, float* logits, long long* labels_out, int part, void* stream);
int pcseg_pooled_feature(pcseg_ctx* ctx, float** pooled);

/* Ragged (un-padded) execution of a zero-padded batch, SURVEY §8(f) rank 1.  x is the reference's padded batch
 * (B, N, 4) as built by collate_fn (pcs.py:44-63); lengths (HOST array, B ints) gives the number of real points
 * of every cloud -- rows lengths[b] .. N-1 of cloud b are the zero pad rows of pcs.py:53-56 (their content is not
 * read).  Only the real rows plus one representative pad row per cloud are computed; the result is the one the
 * padded batch gives: logits (B, N, C) bit-identical to pcseg_forward_eval on the same padded x, pad rows filled
 * with their cloud's pad-row logits.
 * nmax = the padded length N of THIS batch (x is (B, nmax, 4), logits (B, nmax, C)); it may be anything up to the N the
 * context is bound to (0 = exactly N), so that one binding of a generous capacity serves batches whose longest cloud
 * changes from batch to batch (pcs.py:50) without re-binding. */
int pcseg_forward_eval_ragged(pcseg_ctx* ctx, const float* x, const int* lengths, int nmax, float* logits,
                              long long* labels_out, void* stream);

/* Host-only planner of the packed layout used by the *_ragged calls (no CUDA call; the calls run it themselves -- it is
 * exported for inspection, tests and for callers that want the packed row count of a batch).  meta_out (may be NULL)
 * receives  len[B] | off[B+1] | tile_cloud[rows/128] | strips[n][4] = {cloud, first row, end row, first row of the cloud}:
 * cloud b owns packed rows off[b] .. off[b+1] (multiples of 128): its lengths[b] real rows, then ONE representative pad
 * row if lengths[b] < nmax, then filler.  Returns the number of ints of the plan, or -1 (pcseg_last_error). */
long long pcseg_ragged_plan(int B, int nmax, const int* lengths, int* meta_out, long long meta_capacity,
                            long long* rows_out, int* strips_out);

/* Training forward: pcs.py:98-133 under train() (batch statistics, running-stat update, dropout).
 * bn_buffers is updated in place.  dropout_p = 0 disables dropout.  If labels != NULL the weighted
 * cross-entropy of pcs.py:216,247-251 is accumulated into *ce (device, zeroed by this call);
 * class_w may be NULL (all ones).  state (device, may be NULL): its seed is added to `seed`. */
int pcseg_forward_train(pcseg_ctx* ctx, const float* x, const float* params, float* bn_buffers,
                        unsigned long long seed, float dropout_p, float* logits, const long long* labels,
                        const float* class_w, pcseg_ce_accum* ce, const pcseg_step_state* state, void* stream);

/* Ragged training forward (see pcseg_forward_eval_ragged).  The pad rows of the reference are real inputs of the
 * train-mode BatchNorm statistics: the representative pad row of every cloud enters every batch
 * sum, the max-pool and the gradients with multiplicity N - lengths[b], so that losses, logits of real points,
 * running statistics and parameter gradients equal those of the padded batch up to summation order (bf16 rounding
 * noise) when dropout is off.  With dropout the pad rows of one cloud share ONE mask instead of N - lengths[b]
 * independent ones: the BN batch sums of seg_conv2/3 are then an unbiased but noisier estimate of the padded ones.
 * labels is the padded (B, N) tensor (entries of pad rows are ignored and treated as -1).  pcseg_backward after this
 * call runs on the packed rows as well; a caller-supplied dlogits is the gradient of the PADDED logits (B, nmax, C): the
 * representative pad row receives the sum over its cloud's pad rows. */
int pcseg_forward_train_ragged(pcseg_ctx* ctx, const float* x, const int* lengths, int nmax, const float* params, float* bn_buffers,
                               unsigned long long seed, float dropout_p, float* logits, const long long* labels,
                               const float* class_w, pcseg_ce_accum* ce, const pcseg_step_state* state, void* stream);

/* Backward of the training forward: loss.backward(), pcs.py:254.  Gradients of all 38 parameter
 * tensors are written (not accumulated) into `grads`, laid out like `params`.
 * Either dlogits (B,N,C fp32) is given (autograd path), or dlogits == NULL and the gradient of
 * the weighted-mean cross-entropy is formed from the saved logits, labels, class_w and the
 * normaliser *wsum_total (device double: sum of class weights over ALL valid points of the
 * global batch, so data-parallel ranks can pass the all-reduced value).
 * phase 0 = whole backward; phase 1 = seg head .. global_feat (gradients of tensors 10..19 and
 * 30..37 are final afterwards); phase 2 = conv5 .. conv1 (tensors 0..9, 20..29).  The split lets a
 * data-parallel caller all-reduce the first bucket while phase 2 runs. */
int pcseg_backward(pcseg_ctx* ctx, const float* x, const float* params, const float* dlogits,
                   const float* logits, const long long* labels, const float* class_w,
                   const double* wsum_total, float* grads, int phase, void* stream);

/* optimizer.step() of torch.optim.Adam(lr, weight_decay) on the flat arena, pcs.py:217,255.  The gradient is taken as
 * grads * grad_scale / (*grad_div) (grad_div: device double, may be NULL): data-parallel ranks back-propagate the
 * un-normalised loss (pcseg_backward with *wsum_total == 1) and divide by the all-reduced sum of class weights here,
 * which is what the reference's single weighted-mean loss over the gathered logits amounts to (pcs.py:244-251). */
int pcseg_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                    int step, float lr, float beta1, float beta2, float eps, float weight_decay,
                    float grad_scale, const pcseg_step_state* state, const double* grad_div, void* stream);

/* state (device): seed += odd constant, step += 1, bias corrections recomputed.  If state is non-NULL in
 * pcseg_adam_step, `step` and `lr` are taken from it instead of from the arguments. */
int pcseg_step_advance(pcseg_step_state* state, float beta1, float beta2, void* stream);

/* Validation metrics in one pass over logits (pcs.py:292-304 loss / accuracy, pcs.py:319-343 F1 inputs):
 * *ce accumulates the weighted-CE sums and accuracy counters (caller zeroes it), confusion is a C x C matrix
 * (rows = true class, columns = argmax prediction; int64, caller zeroes it), pred_out (optional) receives the
 * argmax labels.  labels == -1 are ignored (pcs.py:54, 216). */
int pcseg_eval_metrics(const float* logits, const long long* labels, long long num_points, int num_classes,
                       const float* class_w, pcseg_ce_accum* ce, unsigned long long* confusion,
                       long long* pred_out, void* stream);

/* Stand-alone GEMM entry used by the unit tests of the tcgen05 kernel (bf16 in, fp32 accumulate).
 *   layout 0: D[M,N] = A[M,K] * B[N,K]^T         (A, B row-major, K contiguous), bf16 out = relu(D + bias)
 *   layout 1: D[M,N] = A[K,M]^T * B[K,N]         (A, B row-major, K = rows),     fp32 out += D (split-K atomics)
 * lda/ldb/ldc are row pitches in elements. */
int pcseg_gemm_test(int layout, int M, int N, int K, const void* A, int lda, const void* B, int ldb,
                    void* D, int ldc, const float* bias, int block_n, void* stream);

/* Test/inspection hook: copy one internal tensor of the bound TRAINING workspace into dst (device
 * memory, dense row-major).  kind: 0 y (pre-BN conv output, bf16 [P][C]), 1 act (post BN+ReLU(+dropout),
 * bf16), 2 dz (grad wrt BN output, bf16), 3 dy (grad wrt conv output, bf16), 4 bnp (float4 [C]:
 * scale, shift, invstd, -mean*invstd), 5 coef (float4 [C] BN-backward coefficients), 6 forward stats
 * (double [2][C]), 7 backward stats (double [2][C]), 8 g (float [B][1024] pooled feature), 9 ystar
 * (float [B][1024] pre-BN extremum), 10 argidx (int32 [B][1024]), 11 cb (float [B][512]), 12 dcb
 * (float [B][512]), 13 dzv (float [B][1024]).  layer = conv index 0..8 for kinds 0..7.
 * rows/cols/elem_bytes describe what was copied. */
int pcseg_debug_copy(pcseg_ctx* ctx, int kind, int layer, void* dst, long long dst_bytes, long long* rows,
                     long long* cols, int* elem_bytes, void* stream);

/* Per-kernel device timing (CUDA events on the launching stream) of the tcgen05 GEMMs of the training
 * step, used by bench.py for the roofline.  tag = conv index (1..8) + 0 forward, + 16 data gradient,
 * + 32 weight gradient.  pcseg_profile_read synchronises the device. */
int pcseg_profile_enable(pcseg_ctx* ctx, int on);
int pcseg_profile_read(pcseg_ctx* ctx, int tag, double* total_ms, long long* launches);
int pcseg_profile_reset(pcseg_ctx* ctx);

/* Number of kernels launched by this library since load (for the bench's gpu_launches claim). */
long long pcseg_launch_count(void);

/* Persistent kernels size their grids by the SM count.  n > 0 caps that count (process-wide) so that concurrently running
 * collectives (the NCCL gradient all-reduce that overlaps backward, replacing DataParallel's reduce-add of pcs.py:209-211)
 * find free SMs instead of delaying a full wave of a persistent GEMM; n <= 0 removes the cap. */
int pcseg_set_sm_limit(int n);

/* ---- data-parallel gradient exchange over NVLink peer memory (one node, 2 / 4 / 8 ranks, one process per GPU) ----
 * Replaces nn.DataParallel's reduce-add of the replica gradients (pcs.py:209-211, behind loss.backward() at pcs.py:254).
 * Every rank maps the gradient arenas and signal blocks of its peers (CUDA IPC) and launches ONE kernel per step on its
 * compute stream (capturable in a CUDA graph): cross-rank barrier, reduce-scatter of the arena with peer loads, barrier,
 * all-gather, barrier.  {loss numerator, sum of class weights} of the deferred loss normalisation travel along (lw_in ->
 * lw_out = sums over the ranks; see pcseg_adam_step's grad_div).
 *   pcseg_ipc_export      : 64-byte IPC handle of the allocation that holds `ptr` + the offset of `ptr` inside it
 *   pcseg_peer_ar_create  : arena = this rank's gradient arena (16-byte aligned, n_floats a multiple of 4), signals = zeroed,
 *                           256-byte aligned device block of pcseg_peer_ar_signal_bytes(n_floats, world) bytes (flags + the
 *                           publish buffer the peers gather the reduced slice from), counters = 4 zeroed uint32
 *   pcseg_peer_ar_open    : map peer `peer`'s arena / signal block from the handles it exported
 *   pcseg_peer_ar_run     : enqueue the all-reduce (every rank must call it once per step) */
typedef struct pcseg_peer_ar pcseg_peer_ar;
int pcseg_ipc_export(const void* ptr, unsigned char* handle_out, long long* offset_out);
long long pcseg_peer_ar_signal_bytes(long long n_floats, int world);
int pcseg_peer_ar_create(pcseg_peer_ar** out, int rank, int world, float* arena, long long n_floats, void* signals,
                         void* counters, const double* lw_in, double* lw_out);
int pcseg_peer_ar_open(pcseg_peer_ar* h, int peer, const unsigned char* arena_handle, long long arena_offset,
                       const unsigned char* sig_handle, long long sig_offset);
int pcseg_peer_ar_run(pcseg_peer_ar* h, void* stream);
int pcseg_peer_ar_destroy(pcseg_peer_ar* h);

#ifdef __cplusplus
}
#endif
#endif
