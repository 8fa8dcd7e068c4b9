// Skew (rotated) IoU device function — K12 in SURVEY.md §2.1.
//
// Implements the algorithm of detectron2's single_box_iou_rotated (the arithmetic behind the
// reference's nms_rotated call at lib/general.py:177 and pairwise_iou_rotated at test.py:135) as
// frozen in SURVEY.md Appendix B: fp32 arithmetic, angle->radian + trig in double, every compare
// against a literal done in double, CUDA-build exchange sort inside the Graham scan.
//
// All fp32 arithmetic goes through __f{add,sub,mul}_rn / __fdiv_rn so that nvcc can never contract
// a multiply-add into an FMA: keep/suppress decisions are compared bit-for-bit with a CPU build.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace ryolo {

struct RPrep {     // 32 bytes / box, produced once per box (trig is per box, not per pair)
  float cx, cy, w, h;
  float c2, s2;    // (float)cos(theta)*0.5f, (float)sin(theta)*0.5f
  float area;      // w*h
  float reach;     // conservative bounding radius (incl. the +-EPS slack of Appendix B step 3)
};

struct V2 {
  float x, y;
};

__device__ __forceinline__ float fm(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fa(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fs(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ V2 vsub(V2 a, V2 b) { return {fs(a.x, b.x), fs(a.y, b.y)}; }
__device__ __forceinline__ float vdot(V2 a, V2 b) { return fa(fm(a.x, b.x), fm(a.y, b.y)); }
__device__ __forceinline__ float vcross(V2 a, V2 b) { return fs(fm(a.x, b.y), fm(b.x, a.y)); }

__device__ __forceinline__ RPrep rprep(float cx, float cy, float w, float h, float deg) {
  RPrep r;
  r.cx = cx; r.cy = cy; r.w = w; r.h = h;
  double theta = (double)deg * 0.01745329251;
  r.c2 = fm((float)cos(theta), 0.5f);
  r.s2 = fm((float)sin(theta), 0.5f);
  r.area = fm(w, h);
  float mn = fminf(fabsf(w), fabsf(h));
  // circumscribed radius, inflated; the last term covers the EPS/|edge| slack of the
  // "corner inside" test for very thin boxes (inf for degenerate boxes => never rejected early)
  r.reach = 0.5f * sqrtf(w * w + h * h) * 1.001f + 1e-3f + 2e-5f / mn;
  return r;
}

// true => the two boxes are certainly disjoint (IoU == 0 exactly under Appendix B)
__device__ __forceinline__ bool rbox_far(const RPrep& a, const RPrep& b) {
  float dx = a.cx - b.cx, dy = a.cy - b.cy, rr = a.reach + b.reach;
  return dx * dx + dy * dy > rr * rr;
}

// Sound early-outs for the NMS decision  IoU(A,B) > thr  (thr >= 1e-6): true => the reference IoU certainly does not
// exceed thr, so the pair can skip the polygon clip.
//   (1) area bound: IoU <= min(area)/max(area); 0.1 % margin covers fp32 rounding of the clipped area
//   (2) separating-axis test on the four box axes with a margin far above the +-EPS slack of Appendix B step 3
//       (a disjoint pair can only produce a sliver of IoU << 1e-6)
__device__ __forceinline__ bool rbox_cannot_exceed(float acx, float acy, float aw, float ah, float ac2, float as2,
                                                   float aarea, float bcx, float bcy, float bw, float bh, float bc2,
                                                   float bs2, float barea, float thr) {
  const float amin = fminf(aarea, barea), amax = fmaxf(aarea, barea);
  if (amin < thr * 0.999f * amax) return true;
  const float dx = bcx - acx, dy = bcy - acy;
  // half-extent vectors: w-axis (c2*w, -s2*w), h-axis (s2*h, c2*h)   [c2 = cos/2, s2 = sin/2]
  const float awx = ac2 * aw, awy = -as2 * aw, ahx = as2 * ah, ahy = ac2 * ah;
  const float bwx = bc2 * bw, bwy = -bs2 * bw, bhx = bs2 * bh, bhy = bc2 * bh;
  const float margin = 0.01f + 1e-4f * (fabsf(aw) + fabsf(ah) + fabsf(bw) + fabsf(bh));
  // unit axes of A: (2*ac2, -2*as2) and (2*as2, 2*ac2); of B likewise
  {
    const float nx = 2.f * ac2, ny = -2.f * as2;
    const float sep = fabsf(nx * dx + ny * dy) - (0.5f * fabsf(aw) + fabsf(nx * bwx + ny * bwy) + fabsf(nx * bhx + ny * bhy));
    if (sep > margin) return true;
  }
  {
    const float nx = 2.f * as2, ny = 2.f * ac2;
    const float sep = fabsf(nx * dx + ny * dy) - (0.5f * fabsf(ah) + fabsf(nx * bwx + ny * bwy) + fabsf(nx * bhx + ny * bhy));
    if (sep > margin) return true;
  }
  {
    const float nx = 2.f * bc2, ny = -2.f * bs2;
    const float sep = fabsf(nx * dx + ny * dy) - (0.5f * fabsf(bw) + fabsf(nx * awx + ny * awy) + fabsf(nx * ahx + ny * ahy));
    if (sep > margin) return true;
  }
  {
    const float nx = 2.f * bs2, ny = 2.f * bc2;
    const float sep = fabsf(nx * dx + ny * dy) - (0.5f * fabsf(bh) + fabsf(nx * awx + ny * awy) + fabsf(nx * ahx + ny * ahy));
    if (sep > margin) return true;
  }
  return false;
}

__device__ __forceinline__ void rcorners(float cx, float cy, const RPrep& b, V2 (&p)[4]) {
  p[0].x = fa(fa(cx, fm(b.s2, b.h)), fm(b.c2, b.w));
  p[0].y = fs(fa(cy, fm(b.c2, b.h)), fm(b.s2, b.w));
  p[1].x = fa(fs(cx, fm(b.s2, b.h)), fm(b.c2, b.w));
  p[1].y = fs(fs(cy, fm(b.c2, b.h)), fm(b.s2, b.w));
  p[2].x = fs(fm(2.f, cx), p[0].x);
  p[2].y = fs(fm(2.f, cy), p[0].y);
  p[3].x = fs(fm(2.f, cx), p[1].x);
  p[3].y = fs(fm(2.f, cy), p[1].y);
}

__device__ __noinline__ float rbox_iou_full(const RPrep& A, const RPrep& B) {
  const double EPS = 1e-5;
  // step 1: shift both centres by the pair midpoint
  double sx = (double)fa(A.cx, B.cx) / 2.0, sy = (double)fa(A.cy, B.cy) / 2.0;
  float ax = (float)((double)A.cx - sx), ay = (float)((double)A.cy - sy);
  float bx = (float)((double)B.cx - sx), by = (float)((double)B.cy - sy);
  if ((double)A.area < 1e-14 || (double)B.area < 1e-14) return 0.f;

  V2 pa[4], pb[4], ea[4], eb[4];
  rcorners(ax, ay, A, pa);
  rcorners(bx, by, B, pb);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    ea[i] = vsub(pa[(i + 1) & 3], pa[i]);
    eb[i] = vsub(pb[(i + 1) & 3], pb[i]);
  }
  V2 pts[24];
  int n = 0;
  // step 3a: edge x edge
#pragma unroll
  for (int i = 0; i < 4; i++) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      float det = vcross(eb[j], ea[i]);
      if ((double)fabsf(det) <= 1e-14) continue;
      V2 d = vsub(pb[j], pa[i]);
      float t1 = __fdiv_rn(vcross(eb[j], d), det);
      float t2 = __fdiv_rn(vcross(ea[i], d), det);
      if ((double)t1 > -EPS && (double)t1 < 1.0f + EPS && (double)t2 > -EPS && (double)t2 < 1.0f + EPS) {
        pts[n].x = fa(pa[i].x, fm(ea[i].x, t1));
        pts[n].y = fa(pa[i].y, fm(ea[i].y, t1));
        n++;
      }
    }
  }
  // step 3b: corners of A inside B, then corners of B inside A
  {
    V2 AB = eb[0], DA = eb[3];
    float abab = vdot(AB, AB), adad = vdot(DA, DA);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      V2 AP = vsub(pa[i], pb[0]);
      float pab = vdot(AP, AB), pad = -vdot(AP, DA);
      if (((double)pab > -EPS) && ((double)pad > -EPS) && ((double)pab < (double)abab + EPS) &&
          ((double)pad < (double)adad + EPS))
        pts[n++] = pa[i];
    }
  }
  {
    V2 AB = ea[0], DA = ea[3];
    float abab = vdot(AB, AB), adad = vdot(DA, DA);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      V2 AP = vsub(pb[i], pa[0]);
      float pab = vdot(AP, AB), pad = -vdot(AP, DA);
      if (((double)pab > -EPS) && ((double)pad > -EPS) && ((double)pab < (double)abab + EPS) &&
          ((double)pad < (double)adad + EPS))
        pts[n++] = pb[i];
    }
  }
  if (n <= 2) return 0.f;

  // step 4: Graham scan (CUDA-build flavour)
  int t = 0;
  for (int i = 1; i < n; i++)
    if (pts[i].y < pts[t].y || (pts[i].y == pts[t].y && pts[i].x < pts[t].x)) t = i;
  V2 start = pts[t];
  V2 q[24];
  float dist[24];
  for (int i = 0; i < n; i++) q[i] = vsub(pts[i], start);
  { V2 tmp = q[0]; q[0] = q[t]; q[t] = tmp; }
  for (int i = 0; i < n; i++) dist[i] = vdot(q[i], q[i]);
  for (int i = 1; i < n - 1; i++) {
    for (int j = i + 1; j < n; j++) {
      float cp = vcross(q[i], q[j]);
      if (((double)cp < -1e-6) || ((double)fabsf(cp) < 1e-6 && dist[i] > dist[j])) {
        V2 tq = q[i]; q[i] = q[j]; q[j] = tq;
        float td = dist[i]; dist[i] = dist[j]; dist[j] = td;
      }
    }
  }
  int k;
  for (k = 1; k < n; k++)
    if ((double)dist[k] > 1e-8) break;
  float inter = 0.f;
  if (k < n) {
    q[1] = q[k];
    int m = 2;
    for (int i = k + 1; i < n; i++) {
      while (m > 1) {
        V2 q1 = vsub(q[i], q[m - 2]), q2 = vsub(q[m - 1], q[m - 2]);
        if (fm(q1.x, q2.y) >= fm(q2.x, q1.y)) m--; else break;
      }
      q[m++] = q[i];
    }
    // step 5: polygon area
    if (m > 2) {
      float area = 0.f;
      for (int i = 1; i < m - 1; i++) area = fa(area, fabsf(vcross(vsub(q[i], q[0]), vsub(q[i + 1], q[0]))));
      inter = (float)((double)area / 2.0);
    }
  }
  return __fdiv_rn(inter, fs(fa(A.area, B.area), inter));
}

// Fast fp32 estimate of the same skew IoU by Sutherland-Hodgman clipping of A's rectangle against B's four edges
// (<= 8 vertices, no sort, no fp64).  It is NOT bit-identical with Appendix B; the NMS kernel only trusts it when the
// estimate is far (> 2e-3) from the threshold and re-runs rbox_iou_full otherwise, so decisions stay exact.
__device__ __forceinline__ float rbox_iou_fast(const RPrep& A, const RPrep& B) {
  if (A.area < 1e-12f || B.area < 1e-12f) return 0.f;
  // work in A-centred coordinates
  V2 pa[4], pb[4];
  rcorners(0.f, 0.f, A, pa);
  rcorners(B.cx - A.cx, B.cy - A.cy, B, pb);
  float px[9], py[9], qx[9], qy[9];
  int n = 4;
#pragma unroll
  for (int i = 0; i < 4; i++) { px[i] = pa[i].x; py[i] = pa[i].y; }
  const float orient = (pb[1].x - pb[0].x) * (pb[2].y - pb[1].y) - (pb[1].y - pb[0].y) * (pb[2].x - pb[1].x);
  const float sgn = orient >= 0.f ? 1.f : -1.f;
#pragma unroll
  for (int e = 0; e < 4; e++) {
    const float ex = pb[(e + 1) & 3].x - pb[e].x, ey = pb[(e + 1) & 3].y - pb[e].y;
    const float bx = pb[e].x, by = pb[e].y;
    int m = 0;
    float sx = px[n - 1], sy = py[n - 1];
    float ds = sgn * (ex * (sy - by) - ey * (sx - bx));
    for (int i = 0; i < n; i++) {
      const float cx = px[i], cy = py[i];
      const float dc = sgn * (ex * (cy - by) - ey * (cx - bx));
      if ((dc >= 0.f) != (ds >= 0.f)) {            // edge s->c crosses the clip line
        const float t = ds / (ds - dc);
        qx[m] = sx + t * (cx - sx);
        qy[m] = sy + t * (cy - sy);
        m++;
      }
      if (dc >= 0.f) { qx[m] = cx; qy[m] = cy; m++; }
      sx = cx; sy = cy; ds = dc;
    }
    n = m;
    if (n == 0) return 0.f;
    for (int i = 0; i < n; i++) { px[i] = qx[i]; py[i] = qy[i]; }
  }
  float a2 = 0.f;
  for (int i = 0; i < n; i++) {
    const int j = (i + 1 == n) ? 0 : i + 1;
    a2 += px[i] * py[j] - px[j] * py[i];
  }
  const float inter = 0.5f * fabsf(a2);
  return inter / (A.area + B.area - inter);
}

// IoU(A, B) with A = the higher-scored ("row") box, B = the candidate ("column") box.
__device__ __forceinline__ float rbox_iou(const RPrep& A, const RPrep& B) {
  if (rbox_far(A, B)) return 0.f;
  return rbox_iou_full(A, B);
}

}  // namespace ryolo
