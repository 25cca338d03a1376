/*
 * gaitk.h -- C-ABI of libgaitk.so: the B200 (sm_100a) implementation of the gait
 * training hot path (per-modality temporal encoders + shared backbone + heads +
 * losses + CAGrad/private backward + SGD + window gather/normalise).
 *
 * Plain C: pointers and sizes only, no torch / C++ types.  Every device buffer
 * is allocated and owned by the caller (PyTorch); the library never frees or
 * keeps a pointer past the call.  All work is enqueued on the caller's CUDA
 * stream (passed as void* == cudaStream_t), with no allocation, no host sync
 * and no host read inside any call, so every entry is CUDA-graph capturable.
 *
 * Return value of every int entry: 0 = OK, >0 = cudaError_t, <0 = GAITK_E_*.
 * gaitk_last_error() returns a thread-local message for the last failure.
 * There is NO CPU fallback: a non-sm_100 device fails gaitk_plan_create with
 * GAITK_E_ARCH.
 *
 * Each entry cites the reference interface (file:line, relative to the
 * reference root) that it replaces.
 */
#ifndef GAITK_H
#define GAITK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GAITK_VERSION 100

#define GAITK_E_BADARG (-1)
#define GAITK_E_SHAPE  (-2)
#define GAITK_E_DTYPE  (-3)
#define GAITK_E_ARCH   (-4)
#define GAITK_E_STATE  (-5)

/* model families */
#define GAITK_FAMILY_WEARGAIT 0   /* data/WearGait/weargait_encoders.py:116-189 WearGaitThreeModal */
#define GAITK_FAMILY_FOG      1   /* train/feature_encoder.py:149-265 MultiModalMultiTaskModel   */
#define GAITK_FAMILY_STAGE    2   /* one stage of a fusion baseline (gaitk_stage_create)          */

#define GAITK_MAX_STREAMS 3
#define GAITK_DENOM_COUNT  24   /* denom[GAITK_DENOM_COUNT + s] = number of labels in stream s's GLOBAL label vector (KL batchmean) */
#define GAITK_DENOM_FLOATS 32   /* size of the `denom` device buffer: [0..3] results, the rest scratch of gaitk_loss_denominators */
#define GAITK_MAX_CLASSES 4
#define GAITK_MAX_PASSES  6

/* arithmetic of the conv/linear contractions */
#define GAITK_DTYPE_F32  0        /* fp32 FFMA, parity 1e-5                           */
#define GAITK_DTYPE_TF32 1        /* tensor-core tf32 inputs, fp32 accumulate, 1e-3   */
#define GAITK_DTYPE_BF16X3 2      /* split bf16 (hi + lo) operands, three tcgen05 passes, fp32 accumulate: operand error 2^-16;
                                     warp-specialised kernel, weight gradients on tcgen05 as well (stream_kernel_ws.cuh) */

/* simplex solver inside CAGrad */
#define GAITK_SOLVER_SLSQP 0      /* restatement of SciPy SLSQP's iteration (reference parity; default) */
#define GAITK_SOLVER_EXACT 1      /* true optimum (closed-form line minima + bisection)                 */
#define GAITK_SOLVER_MEAN  2      /* no CAGrad: shared grad = mean of the task rows, i.e. torch.stack(losses).mean()
                                     .backward() of step_cagrad_three's plain path (weargait_train.py:244-248); pass
                                     private_mult = 1/n_tasks and max_norm = 0 (that path does not clip)        */
/* OR-ed into `solver` of gaitk_step_update: diag has GAITK_DIAG_FLOATS entries and diag[GAITK_DIAG_EXCHANGE] is the sticky
 * status word of the data-parallel exchange (0 ok, -1 a peer never arrived): when it is negative the update leaves
 * parameters and momentum untouched, so replicas never apply an incomplete gradient sum. */
#define GAITK_SOLVER_FLAG_CHECK_EXCHANGE 0x100
#define GAITK_DIAG_FLOATS   24
#define GAITK_DIAG_EXCHANGE 16

typedef struct gaitk_model_desc {
    int32_t family;               /* GAITK_FAMILY_*                                                   */
    int32_t T;                    /* WearGait: win_len (weargait_train.py:655); FoG: pose_length      */
    int32_t enc_out_ch;           /* WearGait C (:671); FoG skeleton_output_dim == sensor_out_channels*/
    int32_t shared_out_ch;        /* S (:673 / configs.py shared_out_channels)                        */
    int32_t backbone_dim;         /* bdim (:672)                                                      */
    int32_t num_classes;          /* K                                                                */
    int32_t use_norm;             /* TaskHead LayerNorm (weargait_encoders.py:33)                     */
    int32_t use_cosine;           /* CosineLinear head (:19-28); implies norm                         */
    int32_t synchronized;         /* one shared head (:133-136 / feature_encoder.py:195-202)          */
    /* FoG / FBG only (configs.py:1-32) */
    int32_t skel_in_dim;          /* skeleton_input_dim (21 FoG, 51 FBG)                              */
    int32_t sensor_in_ch;         /* sensor_in_channels (6 / 3)                                       */
    int32_t sensor_len;           /* sensor_length (426 / 65): pooled to sensor_out_len iff T_in == it*/
    int32_t sensor_out_len;       /* SensorEncoder output_length (101)                                */
    int32_t reserved[3];          /* reserved[0] = proj_ch: SharedLatent3's per-stream Linear(C -> proj_ch) before the
                                     backbone (weargait_encoders.py:299-303); 0 = none                                */
} gaitk_model_desc;

/* One CE-family criterion: CE / weighted CE (weargait_train.py:121-130), GCL
 * (classification_losses.py:79-109: scale s, margin m on the true class) and
 * LDAM (:54-76: per-class margin).  loss = sum_b w[y_b] * nll_b / sum_b w[y_b],
 * nll of softmax(scale * (z - margin[y]*onehot - logit_off)). */
typedef struct gaitk_loss_desc {
    float scale;                              /* s (1 for CE)                                   */
    float margin[GAITK_MAX_CLASSES];          /* subtracted from the true-class logit           */
    float cls_weight[GAITK_MAX_CLASSES];      /* class weights; all 1 when unweighted           */
    int32_t nan_if_degenerate;                /* GCL with equal class counts: 0/0 at :104 -> NaN*/
    int32_t reserved[2];
} gaitk_loss_desc;

typedef struct gaitk_plan gaitk_plan;

int         gaitk_version(void);
const char* gaitk_last_error(void);

/* Plan = host metadata (shapes, parameter layout, launch geometry) for one model
 * on one device.  Replaces the constructors WearGaitThreeModal.__init__
 * (weargait_encoders.py:117-141) / MultiModalMultiTaskModel.__init__
 * (feature_encoder.py:157-220). */
int  gaitk_plan_create(const gaitk_model_desc* desc, int device, gaitk_plan** out);
void gaitk_plan_destroy(gaitk_plan* plan);

/* Canonical flat fp32 parameter layout == the reference's named_parameters()
 * order (state_dict keys, de-aliased).  group: 0 = shared (CAGrad,
 * get_shared_parameters :185-189 / feature_encoder.py:256-265), 1+s = private to
 * stream s, -1 = never receives a gradient (enc_i.ln1). */
int gaitk_param_count(const gaitk_plan* plan);
int gaitk_param_info(const gaitk_plan* plan, int index, char* name, size_t name_cap,
                     int64_t* offset, int64_t* numel, int32_t* group, int32_t* dims /*[4]*/);
int64_t gaitk_param_total(const gaitk_plan* plan);        /* floats in the flat buffer        */
int64_t gaitk_shared_total(const gaitk_plan* plan);       /* P = rows of G                    */
int     gaitk_num_streams(const gaitk_plan* plan);
int     gaitk_stream_in_dim(const gaitk_plan* plan, int stream);   /* channels of stream input */
int     gaitk_stream_in_len(const gaitk_plan* plan, int stream);   /* rows of stream input     */

/* Launch geometry of one stream kernel (introspection for DESIGN.md / profiling). */
int gaitk_stream_geometry(const gaitk_plan* plan, int stream, int dtype, int* ctas_per_sm, int* windows_per_tile,
                          size_t* smem_bytes);

/* Bytes of scratch (per-CTA partial gradients etc.) the step needs for a batch of B. */
size_t gaitk_workspace_bytes(const gaitk_plan* plan, int B);

/* forward: replaces model(xw, xi, xm) (weargait_encoders.py:148-156) /
 * model(skeleton, sensor) (feature_encoder.py:222-254).
 *   x[s]         (B, T_s, D_s) fp32, or a frame store when win_start[s] != NULL
 *   win_start[s] optional int64[B]: first frame of each window in the frame store
 *                (fused gather: dataloader_weargait.py:359-372 never materialised)
 *   enabled_mask bit s clear => stream s sees zeros (weargait_train.py:355-358)
 *   logits[s]    (B, K) fp32 out */
int gaitk_forward(gaitk_plan* plan, const float* params, const float* const* x,
                  const int64_t* const* win_start, int B, uint32_t enabled_mask,
                  float* const* logits, int dtype, void* stream);

/* criterion(logits, y): F.cross_entropy family (classification_losses.py:97-109,
 * weargait_train.py:125-130).  loss_out[1], correct_out[1] (argmax == y count,
 * weargait_train.py:313-315).  dlogits (B,K) optional: d loss / d logits. */
int gaitk_loss(const float* logits, const int64_t* y, int B, int K, const gaitk_loss_desc* desc,
               const float* logit_off, float* loss_out, int32_t* correct_out, float* dlogits,
               void* stream);

/* backward with external logit gradients (generic autograd path):
 * d/dparams of sum_b <dlogits[s][b], logits[s][b]> for stream s, accumulated into
 * grads (flat, same layout as params; += ).  Replaces loss.backward() through
 * one stream.  dx[s] optional (B,T_s,D_s) input gradient (NULL to skip). */
int gaitk_backward(gaitk_plan* plan, const float* params, const float* const* x,
                   const int64_t* const* win_start, int B, uint32_t enabled_mask,
                   const float* const* dlogits, float* grads, void* workspace, size_t workspace_bytes,
                   int dtype, void* stream);

/* Fused training step, phase 1 (replaces forward_batch + criteria + the n_task
 * backward sweeps of CAGrad.get_weighted_loss multitask_weighting.py:680-688 +
 * the private re-derivation weargait_train.py:218-242): one pass per stream over
 * the batch that computes logits, losses, and ALL gradients, leaving
 *   gbuf = [ G (P x n_tasks, column-major by task) | private grads (flat param layout, already
 *            multiplied by private_mult) | loss[n] | correct[n] ]
 * y[s] int64[B]; denom[s] = global sum_b w[y_b] (device float[1] per stream, see
 * gaitk_loss_denominators) so that shards of a data-parallel batch add up.
 * task_mask bit t clear => task t skipped (relaxed-input training with a dropped stream).
 * consistency_lambda != 0 (two-stream plans; fbg_fog_train.py:81-89,121-124): every loss gains
 * 0.5 lambda (KL(softmax_1 || softmax_0) + KL(softmax_0 || softmax_1)) ('batchmean' over the GLOBAL batch,
 * denom[GAITK_DENOM_COUNT]), differentiated through both streams: a logits pass, one coupling kernel, then one
 * recompute + backward pass per (task, stream). */
int gaitk_step_grads(gaitk_plan* plan, const float* params, const float* const* x,
                     const int64_t* const* win_start, const int64_t* const* y, int B,
                     const gaitk_loss_desc* loss /*[n_streams]*/, const float* const* logit_off,
                     const float* denom /*device [n_streams]*/, uint32_t enabled_mask, uint32_t task_mask,
                     float private_mult, float consistency_lambda, float* const* logits /*optional*/,
                     float* gbuf, void* workspace, size_t workspace_bytes, int dtype, void* stream);
int64_t gaitk_gbuf_floats(const gaitk_plan* plan);

/* sum_b w[y_b] per stream from (global) label vectors -> denom[0 .. n_streams) (device).  `denom` must hold
 * GAITK_DENOM_FLOATS floats: the tail is scratch (class histograms, tickets) zeroed by the call. */
int gaitk_loss_denominators(const int64_t* const* y, const int* counts, int n_streams,
                            const gaitk_loss_desc* loss, float* denom, void* stream);

/* Data-parallel exchange over NVLink peer memory, between gaitk_step_grads and gaitk_step_update (replaces the
 * gradient all-reduce a DistributedDataParallel wrapper of the reference trainers would issue).
 * peer_gbuf_dev / peer_flag_dev: DEVICE arrays [world] of peer-mapped pointers (e.g. from a torch symmetric-memory
 * rendezvous): every rank's gbuf of this step's parity (gbuf is double-buffered by step parity) and every rank's flag
 * word.  counter: local device word counting completed exchanges (zero at start; the call increments it on the
 * stream).  The kernel publishes this rank's step number, waits for all peers, and sums the gbufs in rank order into
 * the local gsum (gaitk_gbuf_floats floats), which gaitk_step_update then consumes.  Bit-identical on every rank.
 * If a peer does not arrive within ~4 s of SM clocks (GAITK_P2P_TIMEOUT_CYCLES overrides) the step is flagged in the
 * STICKY status word diag[GAITK_DIAG_EXCHANGE] = -1 (diag: device float[GAITK_DIAG_FLOATS]) instead of hanging the
 * device; gaitk_step_update called with GAITK_SOLVER_FLAG_CHECK_EXCHANGE then skips the parameter update. */
int gaitk_p2p_allreduce(gaitk_plan* plan, const float* const* peer_gbuf_dev, uint32_t* const* peer_flag_dev,
                        uint32_t* counter, int rank, int world, float* gsum, float* diag, void* stream);

/* Fused training step, phase 2 (replaces CAGrad.cagrad + overwrite_grad +
 * clip_grad_norm_ multitask_weighting.py:694-729,748-759,775 and
 * torch.optim.SGD.step weargait_train.py:248,560): on-device Gram matrix, simplex
 * solve, combine, clip, then SGD(momentum, weight decay) on the flat parameter
 * buffer.  Runs identically on every rank after gbuf has been all-reduced.
 * diag (optional, device float[16], or float[GAITK_DIAG_FLOATS] with GAITK_SOLVER_FLAG_CHECK_EXCHANGE): w[3], GTG[9],
 * pre-clip norm, objective, iters, clip factor. */
int gaitk_step_update(gaitk_plan* plan, float* params, float* momentum, const float* gbuf,
                      uint32_t task_mask, float cagrad_c, float max_norm, float lr, float mom,
                      float weight_decay, float* grads_out /*optional flat*/, float* diag, int solver, void* stream);

/* CAGrad alone on an explicit (P x n) column-major matrix (multitask_weighting.py:694-729). */
int gaitk_cagrad(const float* G, int P, int n_tasks, float c, float max_norm, float* shared_grad,
                 float* diag, int solver, void* stream);

/* The simplex solve alone, evaluated ON THE HOST by the same code the device runs (unit tests / debugging):
 * gram3x3 = fp32 G^T G with leading dimension 3; solver GAITK_SOLVER_SLSQP restates SciPy's SLSQP iteration
 * (scipy.optimize.minimize call site multitask_weighting.py:717), GAITK_SOLVER_EXACT is the true optimum.
 * Returns SLSQP's exit mode (0 converged, 8, 9), w_out[n_tasks] (float64). */
int gaitk_cagrad_solve_host(const float* gram3x3, int n_tasks, float alpha, int solver, double* w_out,
                            int* iters_out);

/* SGD alone (torch.optim.SGD, momentum/dampening 0/no nesterov); has_grad (host
 * uint8 per parameter of the plan) mirrors ".grad is None => skipped". */
int gaitk_sgd(gaitk_plan* plan, float* params, const float* grads, float* momentum,
              const uint8_t* has_grad, float lr, float mom, float weight_decay, void* stream);

/* ---- data path --------------------------------------------------------------- */
/* window_indices (dataloader_weargait.py:230-237): host integer arithmetic, bit exact.
 * Writes up to cap (wid,start,stop) triples, returns the count. */
int64_t gaitk_window_indices(int64_t n_frames, int64_t win, int64_t hop, int64_t* out, int64_t cap);

/* per-channel sum, sum of squares (fp64) and count over finite values of a (N, D)
 * fp64 frame matrix, accumulated into acc[3*D] (fit_stats_on_train :183-191). */
int gaitk_stats_accumulate(const double* frames, int64_t N, int D, double* acc, void* stream);
/* (mean, std) from acc (:205-209), std floored at 1e-6; device->device. */
int gaitk_stats_finalize(const double* acc, int D, double* mean, double* stdv, void* stream);
/* apply_stats (:212-227): NaN/Inf -> mean, (x-m)/max(s,1e-6), nan_to_num; fp64 in,
 * fp32 out (the cast WearGaitSyncDataset.__getitem__ :361 does per sample). */
int gaitk_normalize_frames(const double* frames, int64_t N, int D, const double* mean,
                           const double* stdv, float* out, void* stream);
/* gather B windows of T frames from a (N, D) fp32 frame store into (B, T, D); enabled==0 writes
 * zeros (_maybe_zero weargait_train.py:355-358).  Replaces Dataset.__getitem__ + collate + H2D
 * (dataloader_weargait.py:359-372). */
int gaitk_window_gather(const float* frames, int D, const int64_t* win_start, int B, int T,
                        int enabled, float* out, void* stream);
/* relaxed-input evaluation, all seven MASK_COMBOS (weargait_train.py:49-57) from one set of logits: replaces the
 * seven forward_batch_masked + softmax-ensemble passes of eval_with_mask / eval_all_masks (:360-433) and the per-stream
 * accuracies of eval_one_epoch (:322-350).  logits[3] are (B, K) fp32 device pointers, y[3] int64 labels per stream
 * (the same pointer three times in sync mode).  counts (device int32[10]) is ACCUMULATED into: [0..6] hits of the
 * softmax-mean ensemble per mask in MASK_COMBOS order (against y[0]), [7..9] argmax hits per stream. */
int gaitk_mask_eval(const float* const* logits, const int64_t* const* y, int B, int K, int32_t* counts, void* stream);
/* FoG clip preparation (dataloader_fbg_fog.py:24-37,93-113): centre on joint 0, per-clip
 * per-coordinate min-max, zero pad / trim to T_out; fp64 (L, J, 3) clips concatenated in `poses`
 * with clip_start[i], clip_len[i]; out (n_clips, T_out, J*3) fp32. */
int gaitk_fog_prepare_pose(const double* poses, const int64_t* clip_start, const int64_t* clip_len,
                           int n_clips, int J, int T_out, float* out, void* stream);
int gaitk_fog_prepare_sensor(const double* sens, const int64_t* clip_start, const int64_t* clip_len,
                             int n_clips, int D, int T_out, float* out, void* stream);

/* ---- fusion baselines as stages (EarlyFusion3 / CheapXAttn3 weargait_encoders.py:209-245,338-387; EarlyFusionModel /
 * LateFusionModel / ShareLatentModel / CheapXAttnModel feature_encoder.py:346-596; trained by baselines/fusion_train.py:188-202
 * and weargait_train.py --baseline).  These models couple the streams between encoder and backbone, so they run as
 *   encoder stage(s) -> fusion op (concat, or gaitk_xattn_*) -> trunk stage (backbone conv + ReLU + adaptive pool + flatten)
 *   -> gaitk_linear_* head -> gaitk_loss,
 * each stage one launch of the fused stream kernel with the intermediate tensor in HBM. */
#define GAITK_STAGE_CONV_GELU_LN   0   /* WalkwayEncoder / IMUEncoderShallow (weargait_encoders.py:40-69): params w1 b1 lng lnb       */
#define GAITK_STAGE_INSOLE         1   /* InsoleEncoderDeep (:71-101): w1 b1 w2 b2 lng lnb wsk bsk                                   */
#define GAITK_STAGE_LINEAR_LN_RELU 2   /* SkeletonMLP (feature_encoder.py:61-77): w1 b1 lng lnb                                      */
#define GAITK_STAGE_CONV_POOL      3   /* SensorEncoder (:27-58): w1 b1                                                              */
#define GAITK_STAGE_TRUNK          4   /* SharedBackbone (+ .flatten(1)) on a (B, T, CIN) tensor: wbb bbb                            */
typedef struct gaitk_stage_desc {
    int32_t enc;          /* GAITK_STAGE_*                                                         */
    int32_t CIN;          /* input channels of the stage                                           */
    int32_t H;            /* insole hidden width (2 C), else 0                                     */
    int32_t C;            /* encoder output channels (ignored by the trunk)                        */
    int32_t T_in, T;      /* input / output length (T_in != T only for the pooled sensor encoder)  */
    int32_t pool_sensor;  /* SensorEncoder: adaptive pool T_in -> T                                */
    int32_t S, bdim;      /* trunk: backbone channels and pooling bins (encoders: any valid pair)  */
    int32_t reserved[7];
} gaitk_stage_desc;
/* A stage is a one-stream plan (gaitk_plan_destroy / gaitk_param_info / gaitk_param_total / gaitk_workspace_bytes apply);
 * its parameters live in ONE flat buffer in the order listed above. */
int gaitk_stage_create(const gaitk_stage_desc* desc, int device, gaitk_plan** out);
/* out: encoder stage (B, T, C); trunk stage (B, bdim * S) = backbone(x).flatten(1) */
int gaitk_stage_forward(gaitk_plan* stage, const float* params, const float* x, const int64_t* win_start, int B, int zero_input,
                        float* out, void* stream);
/* dout: gradient of the stage output; grads: flat stage layout, gaitk_param_total + 8 floats, accumulated (+=);
 * dx (trunk only, optional): gradient of the trunk input (B, T, CIN).  The forward pass is recomputed inside. */
int gaitk_stage_backward(gaitk_plan* stage, const float* params, const float* x, const int64_t* win_start, int B, int zero_input,
                         const float* dout, float* dx, float* grads, void* workspace, size_t workspace_bytes, void* stream);
/* CheapCrossAttention (weargait_encoders.py:324-336): out = softmax(A B^T / sqrt(d)) B per window; A, B, out (n, T, d) fp32,
 * T <= 128, d in {3, 6, 8, 12, 16}.  Backward recomputes the scores; deterministic (no atomics). */
int gaitk_xattn_forward(const float* A, const float* B, float* out, int n_windows, int T, int d, void* stream);
int gaitk_xattn_backward(const float* A, const float* B, const float* dout, float* dA, float* dB, int n_windows, int T, int d, void* stream);
/* nn.Linear on rows: y (R, O) = x (R, I) W^T (O, I) + bias (heads: I = 128 / 256; projections: I = C).  I <= 256, O <= 32.
 * Backward: dx (optional), dW, db (optional) overwritten; deterministic two-stage batch reduction. */
int gaitk_linear_forward(const float* x, const float* W, const float* bias, float* y, int R, int I, int O, void* stream);
size_t gaitk_linear_workspace_bytes(int R, int I, int O);
int gaitk_linear_backward(const float* x, const float* W, const float* dy, float* dx, float* dW, float* db, int R, int I, int O,
                          void* workspace, size_t workspace_bytes, void* stream);
/* torch.optim.Adam.step over a table of tensors (baselines/fusion_train.py:202): HOST arrays of n_tensors device pointers /
 * element counts; `step` = 1 for the first update. */
int gaitk_adam(float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq, const int64_t* numel,
               int n_tensors, float lr, float beta1, float beta2, float eps, float weight_decay, int step, void* stream);

/* Hardware self-test of the tcgen05 (5th-gen tensor core) layer: runs nops tf32 MMAs (each op = 8 uint32:
 * a_off, a_lbo, a_sbo, b_off, b_lbo, b_sbo (bytes), accumulate, idesc) over shared-memory images of A and B
 * and returns the 128 x ncols fp32 accumulator.  Test infrastructure for the descriptor conventions. */
int gaitk_umma_selftest(const float* A, int nA, const float* B, int nB, const uint32_t* ops, int nops, int ncols,
                        float* D, void* stream);
/* Same harness for kind::f16 MMAs on bf16 operands (A, B = raw bf16 bit patterns).  accumulate: bit 0 = accumulate into D,
 * bits 8..15 = column offset of D in units of 8 TMEM columns.  Pins the K-major / MN-major no-swizzle layouts the
 * split-bf16 stream kernel (stream_kernel_ws.cuh) addresses. */
int gaitk_umma_selftest_bf16(const uint16_t* A, int nA, const uint16_t* B, int nB, const uint32_t* ops, int nops, int ncols,
                             float* D, void* stream);

/* Cost model probe: issues the op list `reps` times from one thread (kind::f16, all accumulating), returns SM clocks
 * {first issue -> completion, issue only} in cycles[2] (device int64).  Design evidence for DESIGN.md. */
int gaitk_umma_bench(const uint32_t* ops, int nops, int reps, int ncols, int smem_bytes, int64_t* cycles, void* stream);
/* the same with n_issuers (1..4) warps issuing concurrently into disjoint accumulator columns; cycles[2 w], cycles[2 w + 1] per issuer */
int gaitk_umma_bench_multi(const uint32_t* ops, int nops, int reps, int ncols, int smem_bytes, int n_issuers, int64_t* cycles, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GAITK_H */
