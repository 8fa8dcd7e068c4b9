/* libryolo_b200 — C ABI of the B200-native R-YOLOv4 hot path.
 *
 * Every entry point takes plain device pointers, sizes and a CUDA stream (as void*); none of them
 * allocates, synchronises or touches torch.  All return 0 on success, non-zero on error
 * (ryolo_last_error() gives the message).  Caller-provided workspaces are sized by the matching
 * *_workspace() function.  Reference interfaces replaced are cited per function (paths relative to
 * the reference repository yingkunwu/R-YOLOv4); INTEGRATION.md shows the reference-side bindings.
 */
#ifndef RYOLO_B200_H_
#define RYOLO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ------------------------------------------------------------------------------- */
int ryolo_abi_version(void);
const char* ryolo_last_error(void);
void ryolo_set_error(const char* msg);
int ryolo_check_device(int device); /* 0 only on a compute-capability 10.x device */
/* Process-wide tuning / timing-experiment switch (no reference counterpart).  Keys: "halo", "dbg", "wg_split",
 * "wg_dbg", "epi_tma", "epi_maxbn", "wg_tapgrp", "bn_bwd", "wg_trans", "sw64", "nacc", "pdl" (see csrc/lib.cu); defaults come from the environment
 * variable RYOLO_<KEY>. */
int ryolo_tune(const char* key, int value);
int ryolo_knob(int id);

/* ---- label encoding ---------------------------------------------------------------------------
 * datasets/base_dataset.py:137-154: xyxyxyxy2xywha (lib/general.py:70-104) + gaussian_label (base_dataset.py:13-31).
 * polys fp32 [T,10] = (image index, class, x1,y1,..,x4,y4 clockwise); out fp32 [T,187] (csl) or [T,7].          */
int ryolo_encode_labels(const float* polys, long long T, int csl, float* out, void* stream);

/* ---- rotated IoU / NMS ------------------------------------------------------------------------
 * detectron2.layers.rotated_boxes.pairwise_iou_rotated  (reference call site test.py:7,135)
 *   a [n,5], b [m,5] fp32 (cx, cy, w, h, angle DEGREES) -> out [n,m] fp32 skew IoU             */
size_t ryolo_pairwise_iou_rotated_workspace(int64_t n, int64_t m);
int ryolo_pairwise_iou_rotated(const float* a, int64_t n, const float* b, int64_t m, float* out,
                               void* workspace, size_t ws_bytes, void* stream);

/* test.py:100-149 get_batch_statistics for a whole batch in ONE launch (eval matching):
 *   dets [B,max_det,7] (x,y,w,h,theta RAD,score,class) with n_det[B] valid rows each (= ryolo_post_process outputs),
 *   targets [T,tcols>=7] (image,class,x,y,w,h,theta RAD), iouv [niou<=32] ascending -> tp uint8 [B,max_det,niou].
 *   Reference semantics verbatim: per target class ascending, detections in index order claim the target of their
 *   class with the highest skew IoU (first maximum) if IoU > iouv[0] and unclaimed; tp[d,k] = IoU > iouv[k].
 *   status int32[1] (zeroed by the caller) is set to 1 when an image has more than 1024 targets.                  */
size_t ryolo_eval_match_workspace(int64_t B, int max_det);
int ryolo_eval_match(const float* dets, const int32_t* n_det, int64_t B, int max_det, const float* targets, int64_t T,
                     int tcols, const float* iouv, int niou, uint8_t* tp, int32_t* status, void* workspace,
                     size_t ws_bytes, void* stream);

/* detectron2.layers.nms.nms_rotated  (reference call site lib/general.py:4,177)
 *   boxes5 [n,5] (cx,cy,w,h,deg), scores [n]; keep int64[n] (indices into boxes5, score-descending,
 *   stable), n_keep int32[1]; suppression when IoU > iou_thr (CUDA-build semantics).            */
size_t ryolo_nms_rotated_workspace(int64_t n);
int ryolo_nms_rotated(const float* boxes5, const float* scores, int64_t n, float iou_thr, int64_t* keep,
                      int32_t* n_keep, void* workspace, size_t ws_bytes, void* stream);

/* post_process  (lib/general.py:136-183), whole batch, no host sync
 *   pred [B,R,6+nc] fp32 (x,y,w,h,theta rad,obj,cls...); class columns *= obj in place if mutate
 *   dets_out [B,max_det,7] (x,y,w,h,theta,score,cls), rows_out int64 [B,max_det] (source row of
 *   each detection), n_out int32 [B].  Reference constants: max_nms 5000, max_det 1500, max_wh 4096 */
size_t ryolo_post_process_workspace(int64_t B, int64_t R, int nc, int max_nms);
int ryolo_post_process(float* pred, int64_t B, int64_t R, int nc, float conf_thres, float iou_thres,
                       int max_nms, int max_det, float max_wh, int mutate, float* dets_out,
                       int64_t* rows_out, int32_t* n_out, void* workspace, size_t ws_bytes, void* stream);

/* ---- rotated-box decode -----------------------------------------------------------------------
 * YoloCSLLayer.forward eval branch (model/yololayer.py:28-56): level [B,3,gs,gs,nc+185] ->
 * rows [row0, row0+3*gs*gs) of out [B,R,nc+6].  anchors_wh: HOST float[6], grid units.          */
int ryolo_decode_csl(const float* level, int64_t B, int gs, int nc, float stride, const float* anchors_wh,
                     float* out, int64_t row0, int64_t R, void* stream);
/* YoloKFIoULayer.forward eval branch (model/yololayer.py:79-105): level [B,na,gs,gs,nc+6];
 * anchors_dev: DEVICE float[na,3] (w,h,rad), grid units.                                       */
int ryolo_decode_kfiou(const float* level, int64_t B, int na, int gs, int nc, float stride,
                       const float* anchors_dev, float* out, int64_t row0, int64_t R, void* stream);

/* ---- target assignment + loss -------------------------------------------------------------------
 * Positive record written by ryolo_build_targets (48 bytes):
 *   int32 b, a, gj, gi; float bx, by, bw, bh; float angle; int32 cls; int32 row; int32 cell      */
size_t ryolo_pos_record_bytes(void);
size_t ryolo_loss_workspace(int64_t B, int na, const int32_t* grid_hw /* host [3][2] */, int64_t T);

/* KFLoss()(pred, target) -> (loss, KFIoU) (lib/loss.py:81-150) on N pairs, O(N) (the reference's accidental
 * [N,1]+[N] -> [N,N] broadcast at :114,:148 is reproduced in value only).  pred/target device fp32 [N,5]
 * (x,y,w,h,theta rad); kfiou [N]; grad = d loss/d pred [N,5] or NULL; loss [1]; workspace >= 16 bytes.           */
int ryolo_kfloss(const float* pred, const float* target, int64_t N, float* kfiou, float* grad, float* loss,
                 void* workspace, size_t ws_bytes, void* stream);

/* ComputeCSLLoss.build_targets (lib/loss.py:270-331, rotated=0) /
 * ComputeKFIoULoss.build_targets (lib/loss.py:427-492, rotated=1).
 *   targets device [T,tcols] (img, cls, x, y, w, h, theta, ...); anchors device [3,na,3] (w,h,rad)
 *   pos device [3][5*na*T] records in the reference's emission order; counts device int32[3]    */
int ryolo_build_targets(const float* targets, int64_t T, int tcols, int rotated, const float* anchors, int na,
                        int64_t B, const int32_t* grid_hw, void* pos, int32_t* counts, void* workspace,
                        size_t ws_bytes, void* stream);

/* ComputeCSLLoss.__call__ (lib/loss.py:191-268, mode 0) / ComputeKFIoULoss.__call__ (:368-425, mode 1)
 * value and gradient in one pass.
 *   levels 3 device pointers [B,na,gh,gw,ch], ch = nc+185 (csl) | nc+6 (kfiou)
 *   grads  3 device pointers (same shapes), fully overwritten with d loss / d level; NULL = value only
 *   hyp    HOST float[7] = box, obj, cls, obj_pw, cls_pw, fl_gamma, lambda_theta
 *   items  device float[8] = reg, theta, conf, cls, total, n_pos[0..2]                           */
int ryolo_loss(int mode, const float* const* levels, float* const* grads, int64_t B, int na, int nc,
               const int32_t* grid_hw, const float* targets, int64_t T, int tcols, const float* anchors,
               const float* hyp, float* items, void* workspace, size_t ws_bytes, void* stream);


/* ---- conv stack ---------------------------------------------------------------------------------
 * Activations are NHWC bf16 "views": a pointer to the first channel of a slice of a (possibly wider,
 * concat) buffer plus the buffer's channel pitch in elements.  Weights are bf16 [Cout][kh][kw][Cin]
 * (derived from the reference's OIHW fp32 state-dict tensors by ryolo_pack_weights).              */
enum { RYOLO_ACT_LINEAR = 0, RYOLO_ACT_LEAKY = 1, RYOLO_ACT_MISH = 2, RYOLO_ACT_SWISH = 3 };
enum { RYOLO_OUT_NHWC_BF16 = 0, RYOLO_OUT_HEAD_F32 = 1 };

/* Fused train-mode BatchNorm2d statistics (model/utils.py:16-17): the raw-output conv epilogue accumulates the
 * per-channel sum / sum of squares of what it stores, and the last CTA to finish turns them into
 * scale = gamma*rsqrt(var+eps), shift = beta-mean*scale, updates the running statistics (momentum, unbiased
 * variance) and num_batches_tracked.  The result is bit-reproducible run to run (like the reference's
 * cudnn.deterministic): within a CTA the order is fixed, across CTAs the sums are 64-bit fixed point.
 * partial = scratch of >= 4*Cout floats (2*Cout 64-bit accumulators), 8-byte aligned, ZERO on entry and left zero on
 * exit (so one zero-initialised scratch serves every layer of a stream); counter (u32) must be zero on entry, or NULL
 * for a DEFERRED finalize (the kernel then only accumulates into partial; see ryolo_scale_shift_act_bn);
 * sum / sumsq (fp32[Cout]) are optional outputs.                                                              */
#define RYOLO_BN_PARTIAL_ROWS 160   /* legacy sizing constant: callers allocate RYOLO_BN_PARTIAL_ROWS*2*Cout floats */
typedef struct ryolo_bn_fuse {
  float* partial; float* sum; float* sumsq; unsigned int* counter;
  const float* gamma; const float* beta;
  float* running_mean; float* running_var; long long* num_batches;     /* nullable */
  float eps, momentum;
  float* scale; float* shift;                                           /* outputs fp32[Cout] */
  float* save_mean; float* save_invstd;                                 /* nullable outputs */
} ryolo_bn_fuse;

typedef struct ryolo_conv_desc {
  const void* x;          /* bf16 NHWC input view                                   */
  int N, H, W, Cin;       /* Cin = channels of the view (multiple of 8)             */
  long long x_cpitch;     /* channel pitch of the input buffer (elements)           */
  const void* w;          /* bf16 [Cout][ksize*ksize*Cin]                           */
  int Cout, ksize, stride;/* ksize 1|3, stride 1|2, pad = (ksize-1)/2               */
  void* out;              /* mode 0: bf16 NHWC view; mode 1: fp32 [N,na,Ho,Wo,ch]   */
  int out_mode;
  long long out_cpitch;
  const float* scale;     /* per-Cout multiplier or NULL  (folded BN gamma*rsqrt)   */
  const float* shift;     /* per-Cout addend or NULL      (folded BN beta / bias)   */
  int act;                /* RYOLO_ACT_*                                            */
  const void* residual;   /* optional bf16 NHWC view added after the activation     */
  long long res_cpitch;
  int head_na, head_ch;   /* mode 1: Cout == head_na*head_ch                        */
  const ryolo_bn_fuse* bn;/* optional (HOST pointer): fused BN statistics, raw epilogue only */
} ryolo_conv_desc;

/* model/utils.py:6-32 Conv (Conv2d + eval-folded BN + activation) / raw conv for train-mode BN.
 * tcgen05 implicit GEMM (csrc/conv.cu).                                                           */
int ryolo_conv2d_forward(const ryolo_conv_desc* d, void* stream);
/* Backward-data of the same Conv2d (autograd's conv backward at train.py:198): dx[N,H,W,Cin] (+)= conv^T(dy, w).
 * dy bf16 NHWC view [N,Ho,Wo,Cout]; wt bf16 [Cin][kh][kw][Cout] (ryolo_pack_weights, layout 2).           */
int ryolo_conv2d_dgrad(const void* dy, long long dy_cpitch, int N, int H, int W, int Cin, int Cout, int ksize,
                       int stride, const void* wt, void* dx, long long dx_cpitch, int accumulate, void* stream);
/* same contract on CUDA cores: device-side checker for the tensor-core path, never a fallback     */
int ryolo_conv2d_reference(const ryolo_conv_desc* d, void* stream);

/* nn.BatchNorm2d train-mode forward, split in three (model/utils.py:16-17):
 *   ryolo_bn_stats     per-channel sum / sum of squares of a bf16 NHWC view (sum, sumsq pre-zeroed, fp32[C])
 *   ryolo_bn_finalize  scale = gamma*rsqrt(var+eps), shift = beta-mean*scale; running-stat EMA (unbiased var),
 *                      num_batches_tracked += 1; optional save_mean / save_invstd for the backward pass
 *   ryolo_scale_shift_act  y = act(x*scale+shift [+ x2*scale2+shift2]) [+ residual]   (bf16 NHWC views, P pixels)
 *                      also used for RepConv's two-branch sum (model/utils.py:209-215) and the Bottleneck
 *                      shortcut (model/utils.py:45-46)                                                   */
int ryolo_bn_stats(const void* x, long long pitch, long long P, int C, float* sum, float* sumsq, void* stream);
/* y = act(BN_train(x)) [+ residual] with the FINALIZE of the producing conv's fused statistics folded in: the conv ran
 * with a ryolo_bn_fuse whose counter == NULL (deferred finalize: it only adds its per-channel totals to bn->partial, a
 * zeroed per-layer accumulator of 4*C floats that nobody clears afterwards); this pass turns the totals into scale /
 * shift per thread, publishes bn->scale / shift / save_mean / save_invstd and moves the running statistics
 * (model/utils.py:16-23 in train mode).  count = N*Ho*Wo.                                                       */
int ryolo_scale_shift_act_bn(const void* x, long long xp, const ryolo_bn_fuse* bn, double count, int act,
                             const void* residual, long long rp, void* y, long long yp, long long P, int C,
                             void* stream);
int ryolo_bn_finalize(const float* sum, const float* sumsq, double count, int C, const float* gamma, const float* beta,
                      float eps, float momentum, float* running_mean, float* running_var, long long* num_batches,
                      float* scale, float* shift, float* save_mean, float* save_invstd, void* stream);
int ryolo_scale_shift_act(const void* x, long long xp, const float* scale, const float* shift, const void* x2,
                          long long x2p, const float* scale2, const float* shift2, int act, const void* residual,
                          long long rp, void* y, long long yp, long long P, int C, void* stream);
/* nn.MaxPool2d (model/utils.py:152,231-233,252) on a bf16 NHWC view; padding acts as -inf              */
int ryolo_maxpool(const void* x, long long xp, int N, int H, int W, int C, int k, int stride, int pad, void* y,
                  long long yp, void* stream);
/* nn.Upsample(scale_factor=factor, nearest) (model/neck.py:9,19) / strided copy into a concat slice     */
int ryolo_resize_copy(const void* x, long long xp, int N, int H, int W, int C, int factor, void* y, long long yp,
                      void* stream);
/* fp32 NCHW image [N,3,H,W] -> bf16 [N,Ho,Wo,Kpad] k x k / stride patches (3*k*k values + zero pad), pad=(k-1)/2:
 * the 3-channel stem conv (3x3/s1 in yolov4/v7, 6x6/s2 in yolov5) becomes a Kpad-channel 1x1 conv             */
int ryolo_stem_im2col(const float* img, int N, int H, int W, int k, int stride, int Kpad, void* y, void* stream);
/* state-dict OIHW fp32 -> bf16: layout 0 [Cout][kh][kw][Cin]; 2 (dgrad) [Cin][kh][kw][Cout]; >= 8 (stem)
 * [Cout][Kpad = layout] in the im2col channel order                                                                     */
int ryolo_pack_weights(const float* w, int Cout, int Cin, int k, int stem, void* out, void* stream);
/* all conv weights of a model in one launch.  table (DEVICE array, sorted by `first`): tensor i holds flat
 * elements [first, first + Cout*Cin*k*k) of the launch; dst = layout 0 (or the stem's [Cout][Kpad], padding
 * pre-zeroed by the caller), dst_t = layout 2 or NULL.  Every tensor's element count (hence every `first`) must be
 * a multiple of 8: a thread packs 8 elements into one 16-byte piece of each layout.                             */
typedef struct ryolo_pack_entry {
  const float* src; void* dst; void* dst_t;
  long long first;
  int Cout, Cin, k, stem;
} ryolo_pack_entry;
int ryolo_pack_weights_multi(const ryolo_pack_entry* table_dev, int n, long long total, void* stream);
int ryolo_unpack_wgrad_multi(const ryolo_pack_entry* table_dev, int n, long long total, void* stream);

/* ---- conv stack backward (autograd's backward of the reference, train.py:198) ------------------------------
 * dwk (fp32, K-major [Cout][kh*kw][Cin] like the packed forward weights; stem: x = the Kpad-channel im2col tensor,
 * dwk = [Cout][Kpad]) += conv_backward_weight(x, dy).  x bf16 NHWC view [N,H,W,Cin]; dy bf16 NHWC view
 * [N,Ho,Wo,Cdy], Cdy >= Cout.  tcgen05 GEMM over pixels with MN-major operands, split-K across CTAs, vector fp32
 * atomics into dwk (csrc/wgrad.cu).  ryolo_unpack_wgrad_multi then adds every dwk into its OIHW gradient
 * (table as for ryolo_pack_weights_multi with src = dwk (fp32), dst = OIHW fp32 gradient, dst_t unused) and, with
 * the default tiled kernel (knob ssa != 0), leaves every dwk element it folded ZEROED for the next step.           */
int ryolo_conv2d_wgrad(const void* x, long long x_cpitch, int N, int H, int W, int Cin, const void* dy,
                       long long dy_cpitch, int Cdy, int Cout, int ksize, int stride, float* dwk, void* stream);
/* d raw = backward of act(BatchNorm2d_train(raw)) given d out; d gamma, d beta (fp32[C], nullable) are ACCUMULATED
 * into (autograd .grad semantics: gradient accumulation over micro-batches, reference train.py:198-202).
 * scale/shift/mean/invstd are what the forward pass saved; sums = fp32[2C] zeroed scratch.  dout may be overwritten
 * with dout*act'(.) (knob bn_bwd < 2): treat it as dead afterwards.                                                                        */
int ryolo_bn_act_bwd(void* dout, long long dp, const void* raw, long long rp, const float* scale,
                     const float* shift, const float* mean, const float* invstd, int act, long long P, int C,
                     float* sums, void* draw, long long op, float* dgamma, float* dbeta, void* dres, long long resp,
                     void* stream);
/* dres (nullable, bf16 view with channel pitch resp): receives a COPY of d out — the gradient of the residual operand
 * of out = residual + act(BN(raw)) (Bottleneck, model/utils.py:35-46) when nothing has been accumulated into it yet. */
/* ds = dout * act'(x1*s1+b1 + x2*s2+b2): backward through RepConv's SiLU of two summed BN branches          */
int ryolo_act_bwd2(const void* dout, long long dp, const void* x1, long long p1, const float* s1, const float* b1,
                   const void* x2, long long p2, const float* s2, const float* b2, int act, void* ds, long long op,
                   long long P, int C, void* stream);
/* dst (+)= src on bf16 NHWC views (gradient fan-in: residuals, concat slices, multiple consumers)            */
int ryolo_add_into(void* dst, long long dpitch, const void* src, long long sp, long long P, int C, int accumulate,
                   void* stream);
/* nn.MaxPool2d backward: dx (+)= route(dy) (first maximum of each window); scratch = fp32[N*H*W*C]           */
int ryolo_maxpool_bwd(const void* x, long long xp, const void* dy, long long dyp, int N, int H, int W, int C, int k,
                      int stride, int pad, void* dx, long long dxp, int accumulate, float* scratch, void* stream);
/* SPP backward (model/utils.py:231-241): the three stride-1 "same" pools (odd k0/k1/k2) of ONE map with H*W <= 1024 in
   one launch: dx (+)= route(dy0,k0) + route(dy1,k1) + route(dy2,k2); the dy views share one channel pitch           */
int ryolo_spp_bwd(const void* x, long long xp, const void* dy0, const void* dy1, const void* dy2, long long dyp, int N, int H,
                  int W, int C, int k0, int k1, int k2, void* dx, long long dxp, int accumulate, void* stream);
/* nearest x2 upsample backward: dx[N,H,W,C] (+)= 2x2 block sums of dy[N,2H,2W,C]                             */
int ryolo_upsample2x_bwd(const void* dy, long long dyp, int N, int H, int W, int C, void* dx, long long dxp,
                         int accumulate, void* stream);
/* loss gradient fp32 [B,na,H,W,ch] -> bf16 NHWC [B,H,W,Cpad] (x mul[c] if given) + dbias[c] += column sums.
 * yolov7's implicit head y = im * (conv(x + ia) + b) (model/utils.py:163-186, model/neck.py:201,208,215), all nullable:
 *   yhead  the head's forward output (same layout as glev): dmul[c] += sum glev * yhead / mul[c]  (= d ImplicitM)
 *   dsum   fp32[>= na*ch], zeroed by the caller: receives the same column sums as dbias (= sum over pixels of d pre,
 *          from which the caller forms d ImplicitA = W^T dsum and the (sum d pre) (x) ia term of d W)            */
int ryolo_head_grad_pack(const float* glev, int B, int na, int H, int W, int ch, int Cpad, const float* mul, void* out,
                         float* dbias, const float* yhead, float* dsum, float* dmul, void* stream);
/* torch.optim.SGD step (train.py:156: momentum, nesterov) on flat fp32 buffers:
 *   g = grad + wd*p;  buf = first ? g : momentum*buf + g;  p -= lr * (nesterov ? g + momentum*buf : buf)      */
int ryolo_sgd_step(float* param, const float* grad, float* buf, long long n, float lr, float momentum,
                   float weight_decay, int nesterov, int first, void* stream);
/* The same SGD update with the learning rate read from DEVICE memory (momentum buffer must already hold valid numbers,
 * zeros before the first step): lets a whole training step captured in a CUDA graph follow the reference's warm-up /
 * one-cycle schedule (train.py:189-193,220) by rewriting one float between replays.                                 */
int ryolo_sgd_step_lrdev(float* param, const float* grad, float* buf, long long n, const float* lr_dev, float momentum,
                         float weight_decay, int nesterov, void* stream);
/* torch.optim.Adam step (train.py:153-154) on flat fp32 buffers; step counts from 1 (bias correction)           */
int ryolo_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                    float beta1, float beta2, float eps, float weight_decay, int step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RYOLO_B200_H_ */
