// Label encoding on the device: polygon targets -> the rows the loss consumes.  Replaces the
// per-image Python of datasets/base_dataset.py:137-154:
//   xyxyxyxy2xywha  (lib/general.py:70-104, a Python loop over boxes)  -> (x, y, w, h, theta), h the long side,
//                   theta in [-pi/2, pi/2) by norm_angle (lib/general.py:7-20)
//   gaussian_label  (datasets/base_dataset.py:13-31) with label = theta*180/pi + 90, 180 classes, sigma 6:
//                   csl[j] = g[(int(90 - label) + j) mod 180],  g[k] = exp(-(k - 90)^2 / 72)
// One warp per target; lanes write the 180-bin row coalesced.  fp32 op order follows the reference (--fmad=false).
#include "common.cuh"
#include "ryolo_b200.h"
#include <math.h>

namespace {

struct CslTable { float g[180]; };

constexpr float kPiF = 3.14159274101257324f;       // fl32(np.pi)
constexpr float kHalfPiF = 1.57079637050628662f;   // fl32(np.pi / 2)

__device__ __forceinline__ float norm2(float a, float b) {
  return sqrtf(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)));
}

__global__ void __launch_bounds__(256)
encode_labels_kernel(const float* __restrict__ polys, long long T, int csl, float* __restrict__ out, CslTable tab) {
  const int lane = threadIdx.x & 31;
  const long long t = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (t >= T) return;
  const float* p = polys + t * 10;
  const float x1 = p[2], y1 = p[3], x2 = p[4], y2 = p[5], x3 = p[6], y3 = p[7], x4 = p[8], y4 = p[9];
  const float x = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(x1, x2), x3), x4), 4.f);
  const float y = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(y1, y2), y3), y4), 4.f);
  float w = __fdiv_rn(__fadd_rn(norm2(x2 - x3, y2 - y3), norm2(x1 - x4, y1 - y4)), 2.f);
  float h = __fdiv_rn(__fadd_rn(norm2(x1 - x2, y1 - y2), norm2(x4 - x3, y4 - y3)), 2.f);
  float th = __fdiv_rn(-__fadd_rn(atan2f(y1 - y2, x1 - x2), atan2f(y4 - y3, x4 - x3)), 2.f);
  if (w >= h) {                                    // the long side is h (lib/general.py:92-99)
    const float tmp = w; w = h; h = tmp;
    th = th > 0.f ? th - kHalfPiF : th + kHalfPiF;
  }
  if (th >= kHalfPiF) th = th - kPiF;              // norm_angle
  else if (th < -kHalfPiF) th = th + kPiF;
  const int width = csl ? 187 : 7;
  float* o = out + t * width;
  if (lane == 0) {
    o[0] = p[0]; o[1] = p[1]; o[2] = x; o[3] = y; o[4] = w; o[5] = h; o[6] = th;
  }
  if (csl) {
    const float label = __fadd_rn(__fdiv_rn(__fmul_rn(th, 180.f), kPiF), 90.f);   // base_dataset.py:145
    int idx = (int)__fsub_rn(90.f, label);                                        // int(num_class/2 - label)
    idx %= 180;
    if (idx < 0) idx += 180;
    for (int j = lane; j < 180; j += 32) {
      int k = idx + j;
      if (k >= 180) k -= 180;
      o[7 + j] = tab.g[k];
    }
  }
}

}  // namespace

extern "C" {

// polys: fp32 [T,10] rows (image index, class, x1,y1, x2,y2, x3,y3, x4,y4), vertices clockwise (the reference's
// `targets` after collate_fn, base_dataset.py:161-167).  out: fp32 [T,187] (csl != 0) or [T,7].
int ryolo_encode_labels(const float* polys, long long T, int csl, float* out, void* stream) {
  RY_CHECK_ARG(T >= 0, "encode_labels: negative row count");
  if (T == 0) return RYOLO_OK;
  RY_CHECK_ARG(polys && out, "encode_labels: null pointer");
  static CslTable tab;
  static bool init = false;
  if (!init) {
    for (int k = 0; k < 180; k++) {
      const double xk = (double)k - 90.0;
      tab.g[k] = (float)exp(-(xk * xk) / (2.0 * 6.0 * 6.0));
    }
    init = true;
  }
  const long long threads = T * 32;
  encode_labels_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(polys, T, csl, out, tab);
  RY_CHECK_LAUNCH();
  return RYOLO_OK;
}

}  // extern "C"
