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

/* ---- rotated IoU / NMS ------------------------------------------------------------------------
 * detectron2.layers.rotated_boxes.pairwise_iou_rotated  (reference call site test.py:7,135)
 *   a [n,5], b [m,5] fp32 (cx, cy, w, h, angle DEGREES) -> out [n,m] fp32 skew IoU             */
size_t ryolo_pairwise_iou_rotated_workspace(int64_t n, int64_t m);
int ryolo_pairwise_iou_rotated(const float* a, int64_t n, const float* b, int64_t m, float* out,
                               void* workspace, size_t ws_bytes, void* stream);

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

#ifdef __cplusplus
}
#endif
#endif /* RYOLO_B200_H_ */
