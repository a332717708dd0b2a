/*
 * udape.h — C-ABI of the B200-native (sm_100a) hot path of the mean-teacher + AdaIN
 * domain-adaptive pose trainer (reference: VisionLearningGroup/UDA_PoseEstimation).
 *
 * The reference has no FFI of its own: its "operator API" is a set of Python callables
 * (SURVEY.md §8b).  Every entry point below replaces one of those callables and cites
 * it as  <file>:<lines>  relative to the reference tree.  The Python package
 * uda_poseestimation_b200 binds these symbols with ctypes and re-exports the reference
 * names/signatures; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions (all entry points):
 *   - plain C types only; device pointers are raw `void*` owned by the caller;
 *   - the library never allocates, frees or retains device memory and keeps no global
 *     mutable state except a thread-local last-error string;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy
 *     default stream) of the calling thread's current device, asynchronously, with no
 *     host synchronisation — every call is CUDA-graph capturable;
 *   - tensors are dense, row-major NCHW ("planes" = N*C or B*K leading items, each a
 *     contiguous H*W plane);
 *   - return 0 on success, a negative UDAPE_ERR_* for argument errors, a positive
 *     cudaError_t for launch failures; udape_last_error() returns the message.
 *   - there is NO CPU fallback: without a CUDA device the calls fail with a cudaError_t.
 */
#ifndef UDAPE_H_
#define UDAPE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UDAPE_VERSION 302 /* major*10000 + minor*100 + patch */

#if defined(__GNUC__)
#define UDAPE_API __attribute__((visibility("default")))
#else
#define UDAPE_API
#endif

/* element type codes */
enum {
    UDAPE_F32 = 0,
    UDAPE_F16 = 1,
    UDAPE_BF16 = 2,
    UDAPE_U8 = 3 /* bool / uint8 masks */
};

/* error codes (negative); positive return values are cudaError_t */
enum {
    UDAPE_OK = 0,
    UDAPE_ERR_NULL = -1,   /* required pointer is NULL */
    UDAPE_ERR_DTYPE = -2,  /* unsupported element type code */
    UDAPE_ERR_SHAPE = -3,  /* non-positive / inconsistent extent */
    UDAPE_ERR_ALIGN = -4,  /* pointer not aligned to its element size */
    UDAPE_ERR_ARG = -5     /* scalar argument out of range */
};

/* ---- library ------------------------------------------------------------------- */
UDAPE_API int udape_version(void);
/* "udape-b200 <ver> sm_100a cuda <ver>" */
UDAPE_API const char* udape_build_info(void);
/* copies the calling thread's last error message (NUL-terminated) into buf; returns its length */
UDAPE_API int udape_last_error(char* buf, size_t buf_bytes);

/* Per-step scalars of a captured CUDA graph (the alpha of train_human.py:349,354 is drawn per step): one tiny launch
 * copies row (*counter % rows) of a device float32 table [rows, cols] to `out` and advances *counter, so that a
 * replay needs neither a host copy nor framework kernels in front of it. */
UDAPE_API int udape_table_feed(const float* table, int rows, int cols, uint32_t* counter, float* out, void* stream);

/* ---- a1: calc_mean_std — adain/function.py:3-11, lib/models/Style_net.py:4-12 ------
 * mean[p] = mean(feat[p,:]); std[p] = sqrt(var_unbiased(feat[p,:]) + eps), p < planes.
 * mean/std are written in `dtype` (the reference returns feat's dtype).  hw == 1 gives
 * std = NaN exactly as torch's unbiased var does. */
UDAPE_API int udape_mean_std(const void* feat, int dtype, int64_t planes, int64_t hw, float eps,
                   void* mean, void* std, void* stream);

/* ---- f4: backward of calc_mean_std — the style loss of the AdaIN decoder pre-training job
 * (adain/net.py:137-143: mse(mean_in, mean_tgt) + mse(std_in, std_tgt) on relu1_1..relu4_1, planes of
 * 256x256 .. 32x32) differentiates through a1:
 *   dfeat[p,i] = dmean[p] / hw + dstd[p] * (feat[p,i] - mean[p]) / ((hw - 1) * std[p])
 * mean / std are the outputs of udape_mean_std for the same feat; dmean / dstd (either may be NULL = 0)
 * are the upstream gradients, all [planes] of `dtype`.  One read and one write of the feature tensor. */
UDAPE_API int udape_mean_std_bwd(const void* feat, const void* mean, const void* std, const void* dmean,
                       const void* dstd, int dtype, int64_t planes, int64_t hw, void* dfeat, void* stream);

/* ---- a2+a3: adaptive_instance_normalization (+ alpha mix) ---------------------------
 * adain/function.py:14-22, lib/models/Style_net.py:21-29 and :167-168.
 * out[p,i] = alpha * ((c[p,i]-mean_c[p])/std_c[p]*std_s[p]+mean_s[p]) + (1-alpha)*c[p,i]
 * content has hw_c elements per plane, style hw_s (only N,C must agree, function.py:15).
 * alpha = 1 reproduces plain AdaIN.  If alpha_dev != NULL the scalar is read from that
 * device address instead (so a captured CUDA graph can change alpha between replays). */
UDAPE_API int udape_adain_mix(const void* content, const void* style, int dtype, int64_t planes,
                    int64_t hw_c, int64_t hw_s, float eps, float alpha,
                    const float* alpha_dev, void* out, void* stream);
/* Several independent jobs of ONE shape and dtype in one launch — the s2t and the t2s direction of a train step
 * (train_human.py:348-356): jobs is a HOST array of n_jobs <= UDAPE_MAX_ADAIN_JOBS entries of device pointers;
 * per job the arithmetic is udape_adain_mix's.  (A launch of this size spends ~10 % of its time ramping up and
 * draining; two directions in one launch pay that once.) */
#define UDAPE_MAX_ADAIN_JOBS 4
typedef struct udape_adain_job {
    const void* content;
    const void* style;
    void* out;
    const float* alpha_dev; /* optional: alpha read from device memory */
    float alpha;
} udape_adain_job;
UDAPE_API int udape_adain_mix_multi(const udape_adain_job* jobs, int n_jobs, int dtype, int64_t planes, int64_t hw_c,
                          int64_t hw_s, float eps, void* stream);

/* ---- a3': per-channel clamp after style transfer — train_human.py:276,351,356 -----------
 * (train_animal.py:301,376,381):  x = maximum(minimum(x.permute(0,2,3,1), recover_max),
 * recover_min).permute(0,3,1,2).  x[planes = N*C, hw] dense NCHW; lo/hi are DEVICE float
 * arrays of `channels` entries (recover_min / recover_max); channel of plane p = p % channels.
 * NaN in x or in a bound propagates like torch.minimum/maximum.  out may alias x. */
UDAPE_API int udape_channel_clamp(const void* x, int dtype, int64_t planes, int64_t channels, int64_t hw,
                        const float* lo, const float* hi, void* out, void* stream);

/* ---- a7 + a11 + a6: heatmap decode — lib/keypoint_detection.py:9-37 (get_max_preds),
 * utils.py:54-75 (get_max_preds_torch), train_human.py:376-383 (conf / pred_position /
 * conf_table), utils.py:77-109 (rectify).
 * Per plane p (= b*K+k) of hm[planes,h,w]: flat argmax with first-index tie-break and
 * NaN-is-maximum (numpy/torch semantics).  Every output pointer is optional (NULL = skip):
 *   idx[p]            int32 flat argmax
 *   preds[p,2]        float32 (x,y) = (idx % w, idx / w), zeroed when max <= 0 (or NaN)
 *   maxvals[p]        max in hm's dtype;  maxvals_f32[p] the same value as float32
 *   position[p,2]     int64 (x,y), never masked (train_human.py:380-381)
 *   conf_table[p]     uint8, max >= occlude_thresh (train_human.py:383)
 *   rectified[planes,h,w]  in hm's dtype: zeros + the (6*sigma+1)^2 unit-peak Gaussian
 *                     pasted at (preds.x, preds.y) with rectify's clipping rules. */
UDAPE_API int udape_decode(const void* hm, int dtype, int64_t planes, int64_t h, int64_t w,
                 int32_t* idx, float* preds, void* maxvals, float* maxvals_f32,
                 int64_t* position, float occlude_thresh, uint8_t* conf_table,
                 double sigma, void* rectified, void* stream);

/* udape_decode + udape_mask_select (below) in ONE launch: the CTA that finishes last selects the kth
 * smallest of maxvals_f32[planes] (required here) and writes thresh_out / tea_mask_out exactly as
 * udape_mask_select does.  ticket: one zeroed uint32, left zero (self-resetting). */
UDAPE_API int udape_decode_select(const void* hm, int dtype, int64_t planes, int64_t h, int64_t w,
                        int32_t* idx, float* preds, void* maxvals, float* maxvals_f32,
                        int64_t* position, float occlude_thresh, uint8_t* conf_table,
                        double sigma, void* rectified, int64_t kth, const float* tea_mask_in,
                        float* thresh_out, uint8_t* tea_mask_out, uint32_t* ticket, void* stream);

/* ---- a11: consistency mask — train_human.py:427-430 ----------------------------------
 * thresh = kth smallest (1-based kth, NaN sorts last: torch.kthvalue) of activates[n];
 * tea_mask_out[i] = (tea_mask_in[i] * activates[i]) > thresh   (tea_mask_in NULL = ones).
 * thresh_out (device float, optional) receives the threshold; no host sync. */
UDAPE_API int udape_mask_select(const float* activates, int64_t n, int64_t kth,
                      const float* tea_mask_in, float* thresh_out, uint8_t* tea_mask_out,
                      void* stream);

/* ---- a8: PCK counts — lib/keypoint_detection.py:40-94 (calc_dists, dist_acc, accuracy)
 * Decodes output[B,K,h,w] and target[B,K,h,w] (dtypes may differ), then per (b,k):
 * valid iff target (x>1 and y>1); hit iff ||pred/norm - target/norm||_2 < thr evaluated
 * in float64 with norm = (h/10, w/10) applied to (x, y) as the reference does (:79).
 * Every (b,k) writes its valid/hit bits to flags[B*K] (caller-owned scratch); the last CTA sums
 * them per joint into hits[K], valid[K] (int32, fully overwritten) — integer sums, so results
 * are exact and order-independent, with no atomics on the counts and no memset.
 * `ticket`: see "tickets" below.  pred[B,K,2] (float32, optional) = decoded output coordinates
 * (accuracy()'s 4th return value); tgt[B,K,2] optional likewise. */
UDAPE_API int udape_pck_counts(const void* output, int out_dtype, const void* target, int tgt_dtype,
                     int64_t batch, int64_t joints, int64_t h, int64_t w, double thr,
                     float* pred, float* tgt, int32_t* hits, int32_t* valid, uint8_t* flags,
                     uint32_t* ticket, void* stream);

/* ---- tickets ---------------------------------------------------------------------------
 * Calls that end in a grid-wide reduction (udape_pck_counts, udape_joints_mse_fwd with
 * loss_mean, udape_cons_fwd, udape_loss_step) take a `ticket`: a caller-owned, 4-byte aligned
 * device uint32 that MUST BE ZERO ON ENTRY.  The last CTA to finish reduces all partials in a
 * fixed order (deterministic results) and the counter wraps back to zero, so the same word can
 * be handed to the next call on the same stream without a memset.  Calls that may run
 * concurrently (different streams) must use different words. */

/* ---- a9: JointsMSELoss — lib/models/loss.py:11-49 ------------------------------------
 * fwd: plane_loss[p] = weight[p] * 0.5 * mean_i (o[p,i]-t[p,i])^2   (reduction='none');
 *      *loss_mean = mean_p plane_loss[p]                             (reduction='mean').
 * weight (optional, [planes]) may be f32/f16/bf16.  plane_loss is required (it doubles as
 * the deterministic reduction scratch); loss_mean optional; ticket (needed with loss_mean): see
 * "tickets" above.
 * bwd: grad_in[p,i] = c[p] * weight[p] * (o-t), c[p] = grad_out[0]/(planes*hw) for 'mean'
 * (grad_per_plane=0) or grad_out[p]/hw for 'none' (grad_per_plane=1); grad_out is a device
 * float pointer; grad_in is written in out_dtype. */
UDAPE_API int udape_joints_mse_fwd(const void* output, int out_dtype, const void* target, int tgt_dtype,
                         const void* weight, int w_dtype, int64_t planes, int64_t hw,
                         float* plane_loss, float* loss_mean, uint32_t* ticket, void* stream);
UDAPE_API int udape_joints_mse_bwd(const void* output, int out_dtype, const void* target, int tgt_dtype,
                         const void* weight, int w_dtype, int64_t planes, int64_t hw,
                         const float* grad_out, int grad_per_plane, void* grad_in,
                         void* stream);

/* ---- a10: ConsLoss — lib/models/loss.py:119-132 --------------------------------------
 * diff = (stu - tea) * tea_mask[b,k];  loss_map[b,i] = mean_k diff^2;
 * loss = mean over (b,i) (only where valid_mask[b,i] != 0 when valid_mask is given).
 * tea_mask optional [B*K] u8 or f32; valid_mask optional [B*hw] u8.
 * plane_partial[B*K] (float scratch), valid_count (int32 scratch, written when valid_mask
 * is given) are caller-owned; ticket: see "tickets" above.  bwd writes
 * grad_stu = 2*g*diff*mask/(K*Nvalid) (0 where invalid) in stu_dtype. */
UDAPE_API int udape_cons_fwd(const void* stu, int stu_dtype, const void* tea, int tea_dtype,
                   const void* tea_mask, int mask_dtype, const uint8_t* valid_mask,
                   int64_t batch, int64_t joints, int64_t hw, float* plane_partial,
                   int32_t* valid_count, float* loss, uint32_t* ticket, void* stream);
UDAPE_API int udape_cons_bwd(const void* stu, int stu_dtype, const void* tea, int tea_dtype,
                   const void* tea_mask, int mask_dtype, const uint8_t* valid_mask,
                   int64_t batch, int64_t joints, int64_t hw, const float* grad_out,
                   const int32_t* valid_count, void* grad_stu, void* stream);

/* ---- a9 + a10 fused: both criteria + the seed of the scaled backward in ONE launch ------
 * train_human.py:425-436:  loss_s = criterion(y_s, label_s, weight_s);
 *                          loss_c = con_criterion(y_t_stu, rectify(y_t_tea), tea_mask=tea_mask);
 *                          loss_all = loss_s + lambda_c*loss_c;  scaler.scale(loss_all).backward()
 * Reads every operand once and writes, besides losses[3] = {loss_all, loss_s, loss_c}:
 *   grad_y_s     = d(grad_scale*loss_all)/d y_s      [planes_s, h, w]   in stu_dtype (optional)
 *   grad_y_t_stu = d(grad_scale*loss_all)/d y_t_stu  [batch_t*joints, h, w] in stu_dtype (optional)
 * y_s / y_t_stu share stu_dtype, label / tea share tgt_dtype.  grad_scale_dev (optional device
 * float, e.g. GradScaler's scale tensor) overrides grad_scale.  If tea == NULL the rectified
 * teacher map is never read: it is evaluated on the fly from tea_preds[batch_t*joints,2]
 * (udape_decode's `preds`) and sigma with the same device functions udape_decode uses to
 * materialise `rectified`, i.e. bit-identical values (utils.py:77-109).  Either pair may be
 * absent (planes_s == 0 or batch_t == 0).  partial[planes_s + batch_t*joints] floats are
 * caller-owned scratch; ticket: see "tickets" above; reductions are deterministic. */
UDAPE_API int udape_loss_step(const void* y_s, const void* label, const void* weight, int w_dtype,
                    int64_t planes_s, const void* y_t_stu, const void* tea, const float* tea_preds,
                    double sigma, const void* tea_mask, int mask_dtype, int64_t batch_t,
                    int64_t joints, int64_t h, int64_t w, int stu_dtype, int tgt_dtype,
                    float lambda_c, float grad_scale, const float* grad_scale_dev, float* partial,
                    float* losses, uint32_t* ticket, void* grad_y_s, void* grad_y_t_stu,
                    void* stream);

/* ---- a4: generate_target (batched) — lib/datasets/util.py:12-70 ----------------------
 * joints[planes,2] float64 image pixels, vis[planes] float32.  mu = trunc(j/stride+0.5)
 * in float64, stride = image/heatmap; out of bounds -> weight 0; window pasted iff
 * weight > 0.5.  target[planes,hm_h,hm_w] float32 (fully written), weight[planes] f32. */
UDAPE_API int udape_gauss_target(const double* joints, const float* vis, int64_t planes, int64_t hm_w,
                       int64_t hm_h, double sigma, double image_w, double image_h,
                       float* target, float* weight, void* stream);

/* ---- a5: draw_labelmap_ori (batched) — lib/datasets/util.py:326-363 ------------------
 * pts[planes,2] int32 (already truncated as pt.to(int32)).  A window touching the border
 * is rejected (vis_out=0, plane untouched/zero).  kind 0 = Gaussian, 1 = Cauchy, evaluated
 * in float64 and stored as float32.  zero_fill != 0: write the whole plane (zeros +
 * window); zero_fill == 0: overwrite only the window inside the existing img (in place). */
UDAPE_API int udape_labelmap(const int32_t* pts, int64_t planes, int64_t h, int64_t w, double sigma,
                   int kind, int zero_fill, float* img, int32_t* vis_out, void* stream);

/* ---- f4: every target set a loader builds per sample, in one launch --------------------------
 * lib/datasets/rendered_hand_pose_mt.py:99,103,115,134,147 calls generate_target five times per sample (the
 * student's, the un-augmented and the teacher view's (64,64) targets plus two (8,8) "small" targets);
 * lib/datasets/real_animal_all_mt.py:275-283,306-311 calls draw_labelmap_ori per joint for three views inside
 * `if tpts[i, 1] > 0`.  A job is one such set over the whole batch: planes = B*K (sample, joint) pairs, the same
 * for every job; jobs is a HOST array of n_jobs <= UDAPE_MAX_TARGET_JOBS entries of DEVICE pointers.  Per plane
 * the arithmetic is udape_gauss_target's / udape_labelmap's (zero-filling form).  gate (optional, uint8 [planes]):
 * 0 = the caller's `if` skipped the joint — the plane is all zero and vis_out is 1, so that the
 * `target_weight *= vis` of :282-283 leaves the weight alone. */
#define UDAPE_MAX_TARGET_JOBS 8
typedef struct udape_target_job {
    const double* joints; /* [planes, 2] image pixels */
    const float* vis;     /* [planes] */
    float* target;        /* [planes, hm_h, hm_w] */
    float* weight;        /* [planes] */
    int32_t hm_w, hm_h;
} udape_target_job;
typedef struct udape_labelmap_job {
    const int32_t* pts;  /* [planes, 2] heatmap pixels, already truncated to int32 (util.py:332) */
    const uint8_t* gate; /* optional [planes] */
    float* img;          /* [planes, h, w] */
    int32_t* vis_out;    /* optional [planes] */
} udape_labelmap_job;
UDAPE_API int udape_gauss_target_multi(const udape_target_job* jobs, int n_jobs, int64_t planes, double sigma,
                             double image_w, double image_h, void* stream);
UDAPE_API int udape_labelmap_multi(const udape_labelmap_job* jobs, int n_jobs, int64_t planes, int64_t h, int64_t w,
                         double sigma, int kind, void* stream);

/* ---- a12/a13: EMA teacher update — utils.py:9-25 (OldWeightEMA), lib/models/ema.py ----
 * A chunk is a contiguous run of one parameter tensor.  udape_ema_plan (host-only helper,
 * no CUDA) splits n_tensors tensors of elem_bytes-sized elements into chunks of
 * <= chunk_elems elements and writes them to out[capacity]; returns the number of chunks
 * (the number required if capacity is too small / out is NULL), negative on error.  The caller copies the table to device
 * memory once and passes it to udape_ema_multi every step.
 * mode 0: dst = dst*a + src*b with three roundings (mul, mul, add — bit-identical to the
 *         reference's  p.mul_(alpha); p.add_(src*(1-alpha))  in fp32);
 * mode 1: dst = src (ModelEMA buffer copy, ema.py:31-36). `numel` counts elements of
 *         `dtype` (use UDAPE_U8 with byte counts to copy buffers of any type). */
typedef struct udape_ema_chunk {
    void* dst;
    const void* src;
    int64_t numel;
} udape_ema_chunk;

UDAPE_API int64_t udape_ema_plan(void* const* dst, const void* const* src, const int64_t* numel,
                       int64_t n_tensors, int64_t elem_bytes, int64_t chunk_elems,
                       udape_ema_chunk* out, int64_t capacity);
UDAPE_API int udape_ema_multi(const udape_ema_chunk* chunks_dev, int64_t n_chunks, int64_t chunk_elems,
                    float a, float b, int dtype, int mode, void* stream);

/* ---- f2: student update + teacher EMA in one pass — train_human.py:436-440 -----------------
 * Replaces  scaler.step(stu_optimizer); tea_optimizer.step()  (GradScaler.unscale_ + torch.optim.Adam
 * / SGD(momentum=0.9, nesterov=True), train_human.py:136-139, + OldWeightEMA.step, utils.py:21-25).
 * A chunk is a contiguous run of one float32 parameter tensor with its gradient, optimizer state and
 * teacher copy.  udape_opt_plan (host-only, no CUDA) splits n_tensors tensors into chunks of
 * <= chunk_elems elements like udape_ema_plan; grad / state1 / state2 / ema may be NULL tables or hold
 * NULL entries (no gradient: the parameter is only folded into the EMA; no teacher: no EMA).
 *   state1 = Adam exp_avg | SGD momentum buffer (NULL when momentum == 0),  state2 = Adam exp_avg_sq.
 * udape_grad_check: *found_inf = 1.0f if any gradient element is non-finite else 0.0f (overwrites;
 * the read-only half of torch._amp_foreach_non_finite_check_and_unscale_).  ws: 2 zeroed uint32 words,
 * left zero.
 * udape_student_step, per element and in torch's single-tensor op order:
 *   g = grad * (1 / *grad_scale)                       (grad_scale NULL: g = grad)
 *   g += weight_decay * p
 *   Adam:  m += (1-beta1)(g-m);  v = beta2 v + (1-beta2) g^2;  p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
 *   SGD :  buf = *fresh ? g : momentum*buf + (1-dampening) g;  g = nesterov ? g + momentum*buf : buf;  p -= lr g
 *          (fresh: per-TENSOR int32 word, non-zero while the tensor's momentum buffer has never been written —
 *          torch clones the gradient into it the first time the parameter is updated, per parameter, and a
 *          loaded checkpoint's buffers are never fresh; NULL = not fresh)
 *   ema = fl(fl(ema*ema_a) + fl(p*ema_b))              (the updated p, three roundings like utils.py:24-25)
 * bc1 = 1-beta1^step, bc2 = 1-beta2^step in double.  step = *step_dev + 1 when step_dev is given, else
 * hyper->step (1-based).  With a ticket (one zeroed, self-resetting uint32) the last CTA of the launch, once
 * every CTA has read them and only if the update was applied, advances *step_dev (advance_step != 0; several
 * param groups share one counter: only the last launch of a step advances it) and zeroes the n_fresh words at
 * fresh_flags (the launch's own tensors).  ticket == NULL: counter and flags are only read.  lr_dev (optional) overrides hyper->lr from device memory (CUDA-graph replays across
 * MultiStepLR milestones).  If *found_inf != 0 the student, its state and the step counter are left
 * untouched and only the EMA runs, as scaler.step() + tea_optimizer.step() do.  float32 only. */
enum { UDAPE_OPT_ADAM = 0, UDAPE_OPT_SGD = 1 };
typedef struct udape_opt_chunk {
    void* param;
    const void* grad;
    void* state1;
    void* state2;
    void* ema;
    const int32_t* fresh; /* SGD: the tensor's "momentum buffer not written yet" word (same for all its chunks) */
    int64_t numel;
} udape_opt_chunk;
typedef struct udape_opt_hyper {
    double lr;
    double beta1;        /* Adam beta1 | SGD momentum */
    double beta2;        /* Adam beta2 | SGD dampening */
    double eps;
    double weight_decay;
    float ema_a, ema_b;  /* teacher = teacher*ema_a + student*ema_b */
    int32_t step;        /* 1-based number of this update (ignored when step_dev is given) */
    int32_t nesterov;
} udape_opt_hyper;

UDAPE_API int64_t udape_opt_plan(void* const* param, const void* const* grad, void* const* state1,
                       void* const* state2, void* const* ema, const int32_t* const* fresh,
                       const int64_t* numel, int64_t n_tensors, int64_t chunk_elems, udape_opt_chunk* out,
                       int64_t capacity);
UDAPE_API int udape_grad_check(const udape_opt_chunk* chunks_dev, int64_t n_chunks, float* found_inf,
                     uint32_t* ws, void* stream);
UDAPE_API int udape_student_step(const udape_opt_chunk* chunks_dev, int64_t n_chunks, int algo,
                       const udape_opt_hyper* hyper, const float* lr_dev, const float* grad_scale,
                       const float* found_inf, int32_t* step_dev, int advance_step, int32_t* fresh_flags,
                       int64_t n_fresh, uint32_t* ticket, void* stream);

/* ---- e: the data-parallel tail of the step over peer memory (NVLink / NVSwitch) ----------------------
 * One process per GPU replaces nn.DataParallel (train_human.py:145-148: replicate, scatter, reduce-add of
 * the gradients onto GPU 0) and the  scaler.step(stu_optimizer); tea_optimizer.step()  tail (:436-438).
 * Every rank keeps its flat float32 gradient bucket, its flat student parameters, a shadow of its slice and a signal pad of
 * UDAPE_DP_PAD_BYTES (zeroed once) in memory the other ranks have mapped (udape_peer_*, CUDA IPC; any other
 * mapping of peer memory works as well — the kernels only see pointers).  udape_dp_peers holds, for every
 * rank, those four addresses AS MAPPED IN THE CALLING PROCESS (entry `rank` is the local memory).
 * n_total = elements of the flat buffers (multiple of 4); rank r owns the slice
 *   [r*S, min((r+1)*S, n_total)),  S = udape_dp_shard_elems(n_total, world)  (a multiple of 4096).
 * One step on one stream of every rank, all ranks taking the same sequence (epoch_dev: a uint32 on this
 * device, zero at start, advanced by udape_dp_gather_ema; flags in the pads are these step numbers):
 *   udape_dp_barrier(READY)         every rank's bucket is complete (backward done) before anyone reads it
 *   udape_dp_reduce_step            g[i] = (g_0[i] + g_1[i] + ... in rank order) * (1/world) for the own slice, by
 *                                   128-bit loads from every rank's bucket; non-finite test of g; and, in the same
 *                                   registers, udape_student_step's arithmetic (unscale, Adam | SGD, same scalars)
 *                                   SPECULATIVELY: new parameters -> this rank's `shadow` slice (S floats, peer-
 *                                   mapped), new state -> the other half of state1 / state2 ([2][S] floats each:
 *                                   exp_avg | momentum buffer / exp_avg_sq shards, half (*step_dev & 1) is current).
 *                                   Posts REDUCED + the non-finite flag to every rank.  ws: 2 zeroed uint32 (left zero).
 *                                   (SGD: buf = grad while *step_dev == 0.)
 *   udape_dp_wait(REDUCED, &found)  every slice is done (and nobody reads this rank's bucket any more);
 *                                   *found_inf = 1.0f if any rank's slice has a non-finite value else 0.0f
 *   udape_dp_gather_ema             the commit: params_local[i] = shadow_owner(i)[i - owner's lo] by 128-bit peer
 *                                   loads (own slice: local), and teacher[i] = fl(fl(teacher[i]*ema_a) +
 *                                   fl(params[i]*ema_b)) in the same pass (teacher NULL: gather only).  *found_inf != 0:
 *                                   nothing is committed, the EMA runs on the unchanged parameters, the state halves do
 *                                   not flip — scaler.step() skipping the update.  Advances *step_dev when committed,
 *                                   and *epoch_dev.
 * Results do not depend on the world size when it is a power of two (exact 1/world) and are bit-identical on
 * every rank; the replicated form is udape_student_step on the same averaged gradient.
 * udape_dp_allreduce_counts: out[i] = sum over ranks (rank order) of counts[i], i < n <= 64 int32 — the PCK
 * hits || valid exchange (keypoint_detection.py:82-92 forms the ratios AFTER summing) in ONE single-CTA launch;
 * its own epoch word (zero at start).  Every wait is bounded by timeout_ns (0 = unbounded): on expiry the
 * pad's error word (UDAPE_DP_PAD_ERR) is set to 1 + phase and the kernel carries on — a missing rank costs a
 * wrong result plus an error the caller can read, never a hung GPU. */
#define UDAPE_DP_MAX_RANKS 8
#define UDAPE_DP_PAD_BYTES 8192
enum { UDAPE_DP_READY = 0, UDAPE_DP_REDUCED = 1, UDAPE_DP_PARAMS = 2, UDAPE_DP_COUNTS = 3, UDAPE_DP_PHASES = 4 };
#define UDAPE_DP_PAD_ERR (UDAPE_DP_PHASES * UDAPE_DP_MAX_RANKS + UDAPE_DP_MAX_RANKS) /* uint32 index of the error word */
typedef struct udape_dp_peers {
    int32_t rank, world;
    float* grads[UDAPE_DP_MAX_RANKS];
    float* params[UDAPE_DP_MAX_RANKS];
    float* shadow[UDAPE_DP_MAX_RANKS]; /* S floats: the rank's slice after the speculative update */
    uint32_t* pads[UDAPE_DP_MAX_RANKS];
} udape_dp_peers;

UDAPE_API int64_t udape_dp_shard_elems(int64_t n_total, int world);
UDAPE_API int udape_dp_barrier(const udape_dp_peers* peers, int phase, const uint32_t* epoch_dev, uint64_t timeout_ns,
                     void* stream);
UDAPE_API int udape_dp_wait(const udape_dp_peers* peers, int phase, const uint32_t* epoch_dev, float* found_inf,
                  uint64_t timeout_ns, void* stream);
UDAPE_API int udape_dp_reduce_step(const udape_dp_peers* peers, int64_t n_total, int algo, const udape_opt_hyper* hyper,
                         const float* lr_dev, const float* grad_scale, const int32_t* step_dev, float* state1,
                         float* state2, const uint32_t* epoch_dev, uint32_t* ws, void* stream);
UDAPE_API int udape_dp_gather_ema(const udape_dp_peers* peers, int64_t n_total, float* teacher, float ema_a, float ema_b,
                        const float* found_inf, int32_t* step_dev, uint32_t* epoch_dev, uint32_t* ticket,
                        void* stream);
UDAPE_API int udape_dp_allreduce_counts(const udape_dp_peers* peers, const int32_t* counts, int n, int32_t* out,
                              uint32_t* epoch_dev, uint64_t timeout_ns, void* stream);
/* Peer-mappable device memory: cudaMalloc'ed (zeroed) arena, 64-byte CUDA IPC handle to hand to the other
 * processes of the node, which map it with udape_peer_open (peer access is enabled on first use). */
UDAPE_API int udape_peer_alloc(size_t bytes, void** ptr);
UDAPE_API int udape_peer_free(void* ptr);
UDAPE_API int udape_peer_export(const void* ptr, unsigned char* handle64);
UDAPE_API int udape_peer_open(const unsigned char* handle64, void** ptr);
UDAPE_API int udape_peer_close(void* ptr);

/* ---- f1: batched multi-stage nearest-neighbour affine re-warp ------------------------------
 * Replaces the per-sample loops of train_human.py:361-372 (teacher recon: k views x three
 * tF.affine(nearest) calls, mean over views), :418-423 (student recon under autocast, needs
 * backward) and :385-412 (occlusion: three-stage warp, patch paste, one-stage warp back);
 * same code in train_animal.py.  A chain of nearest-neighbour resamplings is a composition of
 * integer source-index maps, so the whole chain is one gather out[p] = in[s1(s2(s3(p)))].
 * in[v] / theta[v] are HOST arrays of `views` DEVICE pointers: in[v] is [B,C,H,W] of `dtype`,
 * theta[v] is float32 [B,stages,6] holding, per sample and stage IN EVALUATION ORDER (the stage
 * applied last by the reference comes first), torchvision's rescaled inverse matrix
 * theta[r][k] / (0.5*size_r) rounded in that stage's grid dtype.  Bit s of half_mask marks
 * stage s as computed on a grid_dtype (UDAPE_F16 / UDAPE_BF16) grid — the first tF.affine of a
 * half tensor under autocast; all other arithmetic is float32 in torchvision's / ATen's CPU op
 * order (x*r0 -> fma(y,r1,.) -> +r2; ((g+1)*size-1)/2; rint).  out = mean over views
 * (sequential float32 sum, one division), written in `dtype`.
 * paste (optional, device int32 [B,6] = dst row0,row1,col0,col1, src row0,col0) is the patch copy
 * of :409, applied after `paste_after` evaluated stages.  active (optional, device uint8 [B],
 * views == 1): samples with 0 are copied through unchanged.  The gather cannot run in place.
 * inverse_plan (optional, device, B * udape_rewarp_plan_elems(H, W, elem_bytes) uint16, 16-byte aligned):
 * what autograd saves for the backward — per sample a shared-memory SLOT for every output pixel (the first
 * contributor of a source pixel goes to that pixel's accumulator, later ones to a tail ordered by source pixel
 * and rank) and the list of source pixels with two or more contributors, so that udape_rewarp_bwd is one
 * conflict-free push, one ordered fold per such pixel and one read-out.  Opaque to the caller; valid for the
 * theta / shape / element size it was built for.  Single view only.
 * out == NULL with an inverse_plan builds the plan alone (in[0] may then be NULL too: the plan depends
 * only on theta and the shape), e.g. on a second stream beside the gather. */
UDAPE_API int udape_rewarp_fwd(const void* const* in, const float* const* theta, int views, int stages,
                     int half_mask, int grid_dtype, const int32_t* paste, int paste_after,
                     const uint8_t* active, int64_t B, int64_t C, int64_t H, int64_t W, int dtype,
                     void* out, uint16_t* inverse_plan, void* stream);
/* uint16 elements per sample of the inverse plan for H x W planes of elem_bytes-sized elements;
 * 0 if the plan route does not apply (planes above 4096 pixels, rows that are not 16-byte multiples). */
UDAPE_API int64_t udape_rewarp_plan_elems(int64_t H, int64_t W, int elem_bytes);
/* udape_rewarp_fwd (one view, no paste) + udape_decode_select in ONE launch, for a re-warped map that is only ever
 * decoded — the teacher chain of the step: train_human.py:359-372 (three tF.affine per sample) feeding :376-383
 * (conf / position / conf_table) and :427-430 (activates, k-th value mask).  The re-warped map is never written:
 * every plane is arg-maxed where it is gathered (same ordered keys, first output pixel wins ties, pixels that left
 * the image count as 0), so the outputs equal udape_decode_select(udape_rewarp_fwd(in)) bit for bit while the stage
 * moves B*C*H*W*E bytes instead of 3x that.  Outputs as udape_decode (maxvals_f32 required; idx / preds / position /
 * conf_table optional).  kth = 0: no select (ticket may be NULL); kth in [1, B*C]: as udape_decode_select.
 * Shapes: planes of exactly 4096 pixels, rows a power-of-two number of 16-byte chunks, C <= 64, `in` 16-byte aligned;
 * otherwise UDAPE_ERR_SHAPE / UDAPE_ERR_ALIGN — call the two entries. */
UDAPE_API int udape_rewarp_decode_select(const void* in, const float* theta, int stages, int half_mask, int grid_dtype,
                               int64_t B, int64_t C, int64_t H, int64_t W, int dtype, int32_t* idx, float* preds,
                               float* maxvals_f32, int64_t* position, float occlude_thresh, uint8_t* conf_table,
                               int64_t kth, const float* tea_mask_in, float* thresh_out, uint8_t* tea_mask_out,
                               uint32_t* ticket, void* stream);

/* Gradient of the single-view re-warp w.r.t. its input: grad_in[s] = sum of grad_out[p] over
 * {p : source(p) = s}, float32 accumulation in ascending p, one rounding to `dtype` (deterministic:
 * integer counting / max-key rounds decide the order, no float atomics).  H*W <= 25600.
 * inverse_plan (optional): the plan udape_rewarp_fwd wrote for the same theta / shape / element size;
 * the backward then neither builds nor inverts the map (same sums, same order, same bits; 5x faster). */
UDAPE_API int udape_rewarp_bwd(const void* grad_out, const float* theta, int stages, int half_mask,
                     int grid_dtype, int64_t B, int64_t C, int64_t H, int64_t W, int dtype,
                     void* grad_in, const uint16_t* inverse_plan, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UDAPE_H_ */
