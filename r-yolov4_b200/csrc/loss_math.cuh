// Per-element math of the CSL / KFIoU losses with analytic gradients.
//
// Everything here is __host__ __device__ so that the same source can be compiled for the host by
// tests (tests/test_loss_math_host.py builds a tiny harness) and checked against the oracle's
// torch autograd without a GPU.  References:
//   bbox_ciou           lib/loss.py:36-78
//   KFLoss.forward      lib/loss.py:100-150, xywhr2xywhrsigma lib/general.py:107-133
//   BCEWithLogits/Focal lib/loss.py:10-33 (ATen binary_cross_entropy_with_logits formula)
//   norm_angle          lib/general.py:7-20
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define RY_HD __host__ __device__ __forceinline__
#else
#define RY_HD inline
#endif

// Device builds use the SFU intrinsics (|error| ~1e-6 for arguments within +-pi, far inside the 1e-4 parity bar);
// host builds (tests/host) use libm.
#if defined(__CUDA_ARCH__)
#define RY_SINF(x) __sinf(x)
#define RY_COSF(x) __cosf(x)
#if defined(RY_KF_FAST_MATH)
// Standalone KFLoss (csrc/kfloss.cu): the kernel sits on the instruction-issue roof (ncu: 440 instructions per pair,
// issue slots 74 % busy at 58 % of HBM), and nine IEEE-rounded reciprocals + accurate logf / expf / sqrtf are ~110 of
// them.  The SFU forms are 1-2 ulp (1e-7 relative), three orders of magnitude inside the 1e-4 parity bar.
#define RY_RCP(x) __fdividef(1.f, (x))
#define RY_LOGF(x) __logf(x)
#define RY_EXPF(x) __expf(x)
static __device__ __forceinline__ float ry_sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
#define RY_SQRTF(x) ry_sqrt_approx(x)
#else
#define RY_RCP(x) __frcp_rn(x)
#define RY_LOGF(x) logf(x)
#define RY_EXPF(x) expf(x)
#define RY_SQRTF(x) sqrtf(x)
#endif
#else
#define RY_SINF(x) sinf(x)
#define RY_COSF(x) cosf(x)
#define RY_RCP(x) (1.f / (x))
#define RY_LOGF(x) logf(x)
#define RY_EXPF(x) expf(x)
#define RY_SQRTF(x) sqrtf(x)
#endif

namespace ryolo {

constexpr float kPi = 3.14159274101257324f;       // fl32(np.pi)
constexpr float kHalfPi = 1.57079637050628662f;   // fl32(np.pi / 2)

RY_HD float ry_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

// ---- BCE-with-logits (+ optional focal modulation), value and d/dx --------------------------------
// loss = (1-t)*x + lw*(log1p(exp(-|x|)) + max(-x,0)),  lw = 1 + (pos_weight-1)*t
RY_HD void bce_logits(float x, float t, float pos_weight, float gamma, float* loss, float* dx) {
  const float lw = 1.f + (pos_weight - 1.f) * t;
  const float sp = log1pf(expf(-fabsf(x))) + fmaxf(-x, 0.f);   // softplus(-x)
  const float p = ry_sigmoid(x);
  float l = (1.f - t) * x + lw * sp;
  float g = (1.f - t) + lw * (p - 1.f);
  if (gamma > 0.f) {  // FocalLoss, alpha = 0.25 (lib/loss.py:11)
    const float alpha = 0.25f;
    const float pt = t * p + (1.f - t) * (1.f - p);
    const float af = t * alpha + (1.f - t) * (1.f - alpha);
    const float om = 1.f - pt;
    const float m = powf(om, gamma);
    const float dm = (om > 0.f) ? -gamma * powf(om, gamma - 1.f) * (2.f * t - 1.f) * p * (1.f - p) : 0.f;
    g = af * (g * m + l * dm);
    l = l * af * m;
  }
  *loss = l;
  *dx = g;
}

// ---- CIoU of (x,y,w,h) pairs: value and gradient wrt the predicted box ----------------------------
struct Box4 { float x, y, w, h; };

RY_HD float ciou_fwd_bwd(const Box4 p, const Box4 t, Box4* grad /* d ciou / d p, may be null */) {
  const float l1 = p.x - p.w / 2, r1 = p.x + p.w / 2, t1 = p.y - p.h / 2, b1 = p.y + p.h / 2;
  const float l2 = t.x - t.w / 2, r2 = t.x + t.w / 2, t2 = t.y - t.h / 2, b2 = t.y + t.h / 2;
  const float iw_raw = fminf(r1, r2) - fmaxf(l1, l2), ih_raw = fminf(b1, b2) - fmaxf(t1, t2);
  const float iw = fmaxf(iw_raw, 0.f), ih = fmaxf(ih_raw, 0.f);
  const float inter = iw * ih;
  const float dxc = t.x - p.x, dyc = t.y - p.y;
  const float centre = dxc * dxc + dyc * dyc;
  const float ow_raw = fmaxf(r1, r2) - fminf(l1, l2), oh_raw = fmaxf(b1, b2) - fminf(t1, t2);
  const float ow = fmaxf(ow_raw, 0.f), oh = fmaxf(oh_raw, 0.f);
  const float diag = ow * ow + oh * oh;
  const float uni = p.w * p.h + t.w * t.h - inter;
  const float ud = diag + 1e-15f, id = uni + 1e-15f;
  const float u = centre / ud;
  const float iou = inter / id;
  const float D = atanf(t.w / t.h) - atanf(p.w / p.h);
  const float kv = 4.f / (kPi * kPi);
  const float v = kv * D * D;
  const float alpha = v / ((1.f - iou) + v);   // detached (lib/loss.py:71-73)
  const float raw = iou - (u + alpha * v);
  const float c = fminf(fmaxf(raw, -1.f), 1.f);
  if (grad) {
    Box4 g = {0.f, 0.f, 0.f, 0.f};
    if (raw >= -1.f && raw <= 1.f) {
      // intersection extents
      const float iwx = (iw_raw > 0.f) ? ((r1 < r2 ? 1.f : 0.f) - (l1 > l2 ? 1.f : 0.f)) : 0.f;
      const float iww = (iw_raw > 0.f) ? 0.5f * ((r1 < r2 ? 1.f : 0.f) + (l1 > l2 ? 1.f : 0.f)) : 0.f;
      const float ihy = (ih_raw > 0.f) ? ((b1 < b2 ? 1.f : 0.f) - (t1 > t2 ? 1.f : 0.f)) : 0.f;
      const float ihh = (ih_raw > 0.f) ? 0.5f * ((b1 < b2 ? 1.f : 0.f) + (t1 > t2 ? 1.f : 0.f)) : 0.f;
      const float dI[4] = {ih * iwx, iw * ihy, ih * iww, iw * ihh};
      const float dU[4] = {-dI[0], -dI[1], p.h - dI[2], p.w - dI[3]};
      // enclosing extents
      const float owx = (ow_raw > 0.f) ? ((r1 > r2 ? 1.f : 0.f) - (l1 < l2 ? 1.f : 0.f)) : 0.f;
      const float oww = (ow_raw > 0.f) ? 0.5f * ((r1 > r2 ? 1.f : 0.f) + (l1 < l2 ? 1.f : 0.f)) : 0.f;
      const float ohy = (oh_raw > 0.f) ? ((b1 > b2 ? 1.f : 0.f) - (t1 < t2 ? 1.f : 0.f)) : 0.f;
      const float ohh = (oh_raw > 0.f) ? 0.5f * ((b1 > b2 ? 1.f : 0.f) + (t1 < t2 ? 1.f : 0.f)) : 0.f;
      const float dDg[4] = {2.f * ow * owx, 2.f * oh * ohy, 2.f * ow * oww, 2.f * oh * ohh};
      const float dC[4] = {-2.f * dxc, -2.f * dyc, 0.f, 0.f};
      const float q = p.w * p.w + p.h * p.h;
      const float dV[4] = {0.f, 0.f, -2.f * kv * D * (p.h / q), 2.f * kv * D * (p.w / q)};
      float o[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const float diou = (dI[k] * id - inter * dU[k]) / (id * id);
        const float du = (dC[k] * ud - centre * dDg[k]) / (ud * ud);
        o[k] = diou - du - alpha * dV[k];
      }
      g.x = o[0]; g.y = o[1]; g.w = o[2]; g.h = o[3];
    }
    *grad = g;
  }
  return c;
}

// ---- KFIoU pair terms: xy_loss, kf_loss, KFIoU and gradients wrt the predicted (x,y,w,h,r) ---------
struct Box5 { float x, y, w, h, r; };

RY_HD float norm_angle1(float th) {   // single wrap, lib/general.py:14-15
  if (th >= kHalfPi) th = th - kPi;
  if (th < -kHalfPi) th = th + kPi;
  return th;
}

// returns xy_loss + kf_loss (each clamped at 0, see oracle/hotpath.py kf_loss); *kfiou gets KFIoU.
RY_HD void kf_fwd_bwd(const Box5 p, const Box5 t, float* xy_loss, float* kf_loss, float* kfiou,
                      Box5* grad /* d(xy_loss+kf_loss)/dp, may be null */) {
  const float wp = fminf(fmaxf(p.w, 1e-4f), 1e4f), hp = fminf(fmaxf(p.h, 1e-4f), 1e4f);
  const float wt = fminf(fmaxf(t.w, 1e-4f), 1e4f), ht = fminf(fmaxf(t.h, 1e-4f), 1e4f);
  const float c = RY_COSF(t.r), s = RY_SINF(t.r);
  const float a = (0.5f * wt) * (0.5f * wt), b = (0.5f * ht) * (0.5f * ht);
  const float s00 = c * c * a + s * s * b, s01 = c * s * (a - b), s11 = s * s * a + c * c * b;
  const float idet = RY_RCP(s00 * s11 - s01 * s01);
  const float dx = p.x - t.x, dy = p.y - t.y;
  const float quad = (dx * dx * s11 - 2.f * dx * dy * s01 + dy * dy * s00) * idet;
  const float xl = RY_LOGF(quad + 1.f);
  const float P = wp * wp, Q = hp * hp, U = wt * wt, V = ht * ht;
  const float iP = RY_RCP(P), iQ = RY_RCP(Q), iU = RY_RCP(U), iV = RY_RCP(V);
  const float dr = p.r - t.r;
  const float cd = RY_COSF(dr), sd = RY_SINF(dr);
  const float c2 = cd * cd, s2 = sd * sd;
  const float A2 = 1.f + (P * Q) * (iU * iV) + (P * iU + Q * iV) * c2 + (P * iV + Q * iU) * s2;
  const float B2 = 1.f + (U * V) * (iP * iQ) + (U * iP + V * iQ) * c2 + (U * iQ + V * iP) * s2;
  const float A = RY_SQRTF(A2), B = RY_SQRTF(B2);
  const float K = RY_RCP(A + B - 3.f);
  const float e = RY_EXPF(1.f - K);
  const float kl = e - 1.f;
  *xy_loss = fmaxf(xl, 0.f);
  *kf_loss = fmaxf(kl, 0.f);
  *kfiou = K;
  if (grad) {
    Box5 g = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (xl >= 0.f) {
      const float iq = RY_RCP(quad + 1.f);
      g.x = 2.f * (dx * s11 - dy * s01) * idet * iq;
      g.y = 2.f * (dy * s00 - dx * s01) * idet * iq;
    }
    if (kl >= 0.f) {
      const float G = e * K * K;                      // d kf_loss / dA = d kf_loss / dB
      const float hA = 0.5f * RY_RCP(A), hB = 0.5f * RY_RCP(B);
      const float dAdP = (Q * (iU * iV) + c2 * iU + s2 * iV) * hA;
      const float dAdQ = (P * (iU * iV) + c2 * iV + s2 * iU) * hA;
      const float dBdP = -((U * V) * (iP * iP * iQ) + (U * c2 + V * s2) * (iP * iP)) * hB;
      const float dBdQ = -((U * V) * (iP * iQ * iQ) + (V * c2 + U * s2) * (iQ * iQ)) * hB;
      const float s2d = 2.f * sd * cd;                // sin(2 dr)
      const float dAdr = s2d * ((P * iV + Q * iU) - (P * iU + Q * iV)) * hA;
      const float dBdr = s2d * ((U * iQ + V * iP) - (U * iP + V * iQ)) * hB;
      const bool w_in = (p.w >= 1e-4f && p.w <= 1e4f), h_in = (p.h >= 1e-4f && p.h <= 1e4f);
      g.w = w_in ? G * (dAdP + dBdP) * 2.f * wp : 0.f;
      g.h = h_in ? G * (dAdQ + dBdQ) * 2.f * hp : 0.f;
      g.r = G * (dAdr + dBdr);
    }
    *grad = g;
  }
}

}  // namespace ryolo
