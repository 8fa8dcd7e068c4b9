// Test harness: compiles the PRODUCT's skew-IoU source (r-yolov4_b200/csrc/rotated_iou.cuh — the same text the CUDA
// NMS kernel uses) for the host, so that the fast-path gate and the early-outs of nms_mask_kernel can be checked
// against the oracle over millions of adversarial pairs without a GPU.
// Built by tests/test_rotated_iou_host.py with g++ -ffp-contract=off (nms.cu is compiled with --fmad=false).
#include <cstdint>
#include "rotated_iou.cuh"

using namespace ryolo;

extern "C" {

// a, b: [n,5] (cx, cy, w, h, deg), pair i = (a[i], b[i]).
// out_fast / out_full: the two IoU values;  gate: rbox_fast_ok;  decide: what nms_mask_kernel decides for the pair
// (bounding-circle reject -> area-ratio / separating-axis bound -> rbox_iou_exceeds), for threshold thr.
void hr_pairs(int64_t n, const float* a, const float* b, float thr, float* out_fast, float* out_full, uint8_t* gate,
              uint8_t* decide, float* out_pair) {
  const bool use_reject = thr >= 0.f, use_bounds = thr >= 1e-6f;
  for (int64_t i = 0; i < n; i++) {
    const float* p = a + 5 * i;
    const float* q = b + 5 * i;
    RPrep A = rprep(p[0], p[1], p[2], p[3], p[4]), B = rprep(q[0], q[1], q[2], q[3], q[4]);
    out_fast[i] = rbox_iou_fast(A, B);
    out_full[i] = rbox_iou_full(A, B);
    if (out_pair) out_pair[i] = rbox_iou(A, B);         // what pairwise_iou_kernel stores
    gate[i] = rbox_fast_ok(A, B) ? 1 : 0;
    bool live = true;
    if (use_reject) {
      live = !rbox_far(A, B);
      if (live && use_bounds)
        live = !rbox_cannot_exceed(A.cx, A.cy, A.w, A.h, A.c2, A.s2, A.area, B.cx, B.cy, B.w, B.h, B.c2, B.s2, B.area, thr);
    }
    decide[i] = (live && rbox_iou_exceeds(A, B, thr)) ? 1 : 0;
  }
}
}
