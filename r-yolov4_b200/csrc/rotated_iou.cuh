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
#include <math.h>
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define RY_RDEV __device__ __forceinline__
#define RY_RDEV_NOINLINE __device__ __noinline__
#else
// Host build (tests/host/rotated_iou_host.cpp, g++ -ffp-contract=off): the same source decides the same pairs on the
// CPU, so the fast-path gate below can be checked against the oracle over millions of adversarial pairs without a GPU.
#define RY_RDEV static inline
#define RY_RDEV_NOINLINE static
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
#endif

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

RY_RDEV float fm(float a, float b) { return __fmul_rn(a, b); }
RY_RDEV float fa(float a, float b) { return __fadd_rn(a, b); }
RY_RDEV float fs(float a, float b) { return __fsub_rn(a, b); }
RY_RDEV V2 vsub(V2 a, V2 b) { return {fs(a.x, b.x), fs(a.y, b.y)}; }
RY_RDEV float vdot(V2 a, V2 b) { return fa(fm(a.x, b.x), fm(a.y, b.y)); }
RY_RDEV float vcross(V2 a, V2 b) { return fs(fm(a.x, b.y), fm(b.x, a.y)); }

RY_RDEV RPrep rprep(float cx, float cy, float w, float h, float deg) {
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

// Appendix B shifts both centres by the pair's midpoint before it builds the corners, so each box ends up ~|d|/2 from the
// origin.  A box whose sides are below the fp32 spacing at that magnitude (ulp(|d|/2) ~ 6e-8 |d|) collapses to a point
// there; its edge vectors become exactly 0 and the "corner inside" test (0 > -EPS && 0 < 0 + EPS) then accepts EVERY
// corner of the other box, wherever it is: the frozen spec reports IoU = area ratio >> 1 for such a pair even 80 000
// units apart (a sub-0.01 px box against any box of another class, 4096*cls offsets).  Geometric early-outs are only
// sound when both boxes stay resolvable: smallest side > 160 spacings of the centre distance.
RY_RDEV bool rbox_resolvable(float smin, float dx, float dy) {
  return smin > 1e-5f * fmaxf(fabsf(dx), fabsf(dy));
}

// true => the two boxes are certainly disjoint (IoU == 0 exactly under Appendix B)
RY_RDEV bool rbox_far_soa(float acx, float acy, float aw, float ah, float areach, float bcx, float bcy, float bw,
                          float bh, float breach) {
  const float dx = acx - bcx, dy = acy - bcy, rr = areach + breach;
  if (!(dx * dx + dy * dy > rr * rr)) return false;
  const float smin = fminf(fminf(fabsf(aw), fabsf(ah)), fminf(fabsf(bw), fabsf(bh)));
  return rbox_resolvable(smin, dx, dy);
}
RY_RDEV bool rbox_far(const RPrep& a, const RPrep& b) {
  return rbox_far_soa(a.cx, a.cy, a.w, a.h, a.reach, b.cx, b.cy, b.w, b.h, b.reach);
}

// Sound early-outs for the NMS decision  IoU(A,B) > thr  (thr >= 1e-6): true => the reference IoU certainly does not
// exceed thr, so the pair can skip the polygon clip.
//   (1) area bound: IoU <= min(area)/max(area); 0.1 % margin covers fp32 rounding of the clipped area
//   (2) separating-axis test on the four box axes with a margin far above the +-EPS slack of Appendix B step 3
//       (a disjoint pair can only produce a sliver of IoU << 1e-6)
RY_RDEV bool rbox_cannot_exceed(float acx, float acy, float aw, float ah, float ac2, float as2,
                                                   float aarea, float bcx, float bcy, float bw, float bh, float bc2,
                                                   float bs2, float barea, float thr) {
  // Appendix B's absolute EPS = 1e-5 slack (on squared-length quantities) dominates sub-unit boxes: disjoint slivers
  // can "overlap" there, so no geometric bound is sound for them.
  const float smin = fminf(fminf(fabsf(aw), fabsf(ah)), fminf(fabsf(bw), fabsf(bh)));
  if (!(smin >= 1.f) || !rbox_resolvable(smin, bcx - acx, bcy - acy)) return false;
  // The area bound assumes intersection <= min(area).  With (nearly) parallel edges Appendix B can mis-order its
  // near-duplicate points and report up to a few times the smaller area (see rbox_fast_ok), so it only applies to
  // skewed pairs.
  const float sd = 4.f * (as2 * bc2 - ac2 * bs2), cd = 4.f * (ac2 * bc2 + as2 * bs2);
  const float amin = fminf(aarea, barea), amax = fmaxf(aarea, barea);
  if (fminf(fabsf(sd), fabsf(cd)) >= 0.02f && amin < thr * 0.999f * amax) return true;
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

RY_RDEV void rcorners(float cx, float cy, const RPrep& b, V2 (&p)[4]) {
  p[0].x = fa(fa(cx, fm(b.s2, b.h)), fm(b.c2, b.w));
  p[0].y = fs(fa(cy, fm(b.c2, b.h)), fm(b.s2, b.w));
  p[1].x = fa(fs(cx, fm(b.s2, b.h)), fm(b.c2, b.w));
  p[1].y = fs(fs(cy, fm(b.c2, b.h)), fm(b.s2, b.w));
  p[2].x = fs(fm(2.f, cx), p[0].x);
  p[2].y = fs(fm(2.f, cy), p[0].y);
  p[3].x = fs(fm(2.f, cx), p[1].x);
  p[3].y = fs(fm(2.f, cy), p[1].y);
}

RY_RDEV_NOINLINE float rbox_iou_full(const RPrep& A, const RPrep& B) {
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
// (<= 8 vertices, no sort, no fp64).  It is NOT bit-identical with Appendix B; the NMS kernel only trusts it under
// rbox_fast_ok() and when the estimate is farther than kFastBand from the threshold (see rbox_iou_exceeds).
RY_RDEV float rbox_iou_fast(const RPrep& A, const RPrep& B) {
  if (A.area < 1e-12f || B.area < 1e-12f) return 0.f;
  // work in A-centred coordinates
  V2 pa[4], pb[4];
  rcorners(0.f, 0.f, A, pa);
  rcorners(B.cx - A.cx, B.cy - A.cy, B, pb);
  float px[9], py[9], qx[9], qy[9];
  int n = 4;
#pragma unroll
  for (int i = 0; i < 4; i++) { px[i] = pa[i].x; py[i] = pa[i].y; }
  // (explicit fmaf: nms.cu is compiled with --fmad=false for the exact path; the estimate wants the fused forms)
  const float orient = fmaf(pb[1].x - pb[0].x, pb[2].y - pb[1].y, -(pb[1].y - pb[0].y) * (pb[2].x - pb[1].x));
  const float sgn = orient >= 0.f ? 1.f : -1.f;
#pragma unroll
  for (int e = 0; e < 4; e++) {
    const float ex = pb[(e + 1) & 3].x - pb[e].x, ey = pb[(e + 1) & 3].y - pb[e].y;
    const float bx = pb[e].x, by = pb[e].y;
    int m = 0;
    float sx = px[n - 1], sy = py[n - 1];
    const float sex = sgn * ex, sey = sgn * ey;
    float ds = fmaf(sex, sy - by, -sey * (sx - bx));
    for (int i = 0; i < n; i++) {
      const float cx = px[i], cy = py[i];
      const float dc = fmaf(sex, cy - by, -sey * (cx - bx));
      if ((dc >= 0.f) != (ds >= 0.f)) {            // edge s->c crosses the clip line
        const float t = ds / (ds - dc);
        qx[m] = fmaf(t, cx - sx, sx);
        qy[m] = fmaf(t, cy - sy, sy);
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
    a2 = fmaf(px[i], py[j], fmaf(-px[j], py[i], a2));
  }
  const float inter = 0.5f * fabsf(a2);
  return inter / (A.area + B.area - inter);
}

// When may the NMS decision trust the fast estimate?  Two error sources separate it from Appendix B:
//   (1) Appendix B itself departs from the true polygon overlap through its absolute EPS = 1e-5 slack on SQUARED-length
//       quantities (a corner up to 1e-5/|edge| outside still counts as inside) and the 1e-6 / 1e-8 hull cut-offs:
//       negligible once every side is >= 1 unit (<= 4e-5 units of slack), unbounded for sub-unit boxes;
//   (2) the estimate's own fp32 rounding, ~C * 2^-24 * L^2 of absolute area error with L the coordinate range in the
//       A-centred frame (<= 2 (reach_A + reach_B)), to be compared with the union >= max(area).
// The gate keeps both below ~3e-4 of IoU (tests/test_rotated_iou_host.py measures the worst case over adversarial
// pairs: thin slivers, sub-unit boxes, near-coincident edges, near-duplicates, 4096*cls offsets); everything else,
// and every estimate within kFastBand of the threshold, takes the bit-exact path.
constexpr float kFastBand = 2e-3f;
// true when one of the four values |a +- S|, |a +- T| lies within delta of `half` (a corner on an edge line)
RY_RDEV bool rbox_axis_touch(float a, float p, float q, float half, float delta) {
  const float S = p + q, T = fabsf(p - q);
  return fabsf(a + S - half) < delta || fabsf(fabsf(a - S) - half) < delta || fabsf(a + T - half) < delta ||
         fabsf(fabsf(a - T) - half) < delta;
}
#ifndef RY_FAST_MIN_SKEW
#define RY_FAST_MIN_SKEW 0.002f
#endif
#ifndef RY_FAST_CORNER_DELTA
#define RY_FAST_CORNER_DELTA 2.5e-4f
#endif
constexpr float kFastMinSkew = RY_FAST_MIN_SKEW;     // |sin| / |cos| of the angle between the boxes (0.002 ~ 0.11 degrees)
RY_RDEV bool rbox_fast_ok(const RPrep& A, const RPrep& B) {
  const float sa = fminf(fabsf(A.w), fabsf(A.h)), sb = fminf(fabsf(B.w), fabsf(B.h));
  if (!(sa >= 1.f && sb >= 1.f) || !rbox_resolvable(fminf(sa, sb), A.cx - B.cx, A.cy - B.cy)) return false;
  const float L = 2.f * (A.reach + B.reach);
  if (!(L * L <= 256.f * fmaxf(A.area, B.area))) return false;
  // (3) near-parallel edges: Appendix B intersects (almost) coincident edge lines with a tiny determinant and then
  //     de-duplicates / orders the resulting points with absolute cut-offs, which for near-duplicate boxes yields
  //     "IoU" values such as 1/3, 3 or 55 where the geometric overlap is ~1 (tests/test_rotated_iou_host.py shows
  //     them).  Those values ARE the frozen spec, so such pairs must take the bit-exact path.  Edges of the two
  //     rectangles are parallel when sin or cos of the angle difference vanishes.
  const float sd = 4.f * (A.s2 * B.c2 - A.c2 * B.s2), cd = 4.f * (A.c2 * B.c2 + A.s2 * B.s2);
  if (!(fminf(fabsf(sd), fabsf(cd)) >= kFastMinSkew)) return false;
  // (4) a corner of one box (almost) on an edge line of the other: two or three of Appendix B's candidate points then
  //     (almost) coincide, its Graham scan orders the near-duplicate of the start point by a meaningless angle and
  //     drops real vertices (observed: 0.318 where the overlap is 0.877, corner 3e-5 px from the edge).  The damage
  //     stops ~2.5e-5 of the box size away from the singular position; the gate keeps 10x that distance (a sweep over
  //     21.6 M adversarial decisions found no disagreement down to 4x; flush / nearly coincident edges are this case
  //     too, which is why the skew gate (3) can be two orders of magnitude tighter than the corner gate).
  const float delta = RY_FAST_CORNER_DELTA * (A.reach + B.reach);
  const float dx = A.cx - B.cx, dy = A.cy - B.cy;
  // half-extent vectors (see rcorners): w-axis (c2 w, -s2 w), h-axis (s2 h, c2 h); unit axes are 2*(c2, -s2), 2*(s2, c2)
  const float apx = A.c2 * A.w, apy = -A.s2 * A.w, aqx = A.s2 * A.h, aqy = A.c2 * A.h;
  const float bpx = B.c2 * B.w, bpy = -B.s2 * B.w, bqx = B.s2 * B.h, bqy = B.c2 * B.h;
  // The four corners of A sit at uD +- uP +- uQ along an axis of B (D = centre offset, P / Q = A's half-extent vectors):
  // their absolute values are |a + S|, |a - S|, |a + T|, |a - T| with a = |uD|, S = |uP| + |uQ|, T = ||uP| - |uQ||.
  {  // corners of A in B's frame
    const float ux = 2.f * B.c2, uy = -2.f * B.s2, vx = 2.f * B.s2, vy = 2.f * B.c2;
    if (rbox_axis_touch(fabsf(fmaf(ux, dx, uy * dy)), fabsf(fmaf(ux, apx, uy * apy)), fabsf(fmaf(ux, aqx, uy * aqy)),
                        0.5f * fabsf(B.w), delta)) return false;
    if (rbox_axis_touch(fabsf(fmaf(vx, dx, vy * dy)), fabsf(fmaf(vx, apx, vy * apy)), fabsf(fmaf(vx, aqx, vy * aqy)),
                        0.5f * fabsf(B.h), delta)) return false;
  }
  {  // corners of B in A's frame
    const float ux = 2.f * A.c2, uy = -2.f * A.s2, vx = 2.f * A.s2, vy = 2.f * A.c2;
    if (rbox_axis_touch(fabsf(fmaf(ux, dx, uy * dy)), fabsf(fmaf(ux, bpx, uy * bpy)), fabsf(fmaf(ux, bqx, uy * bqy)),
                        0.5f * fabsf(A.w), delta)) return false;
    if (rbox_axis_touch(fabsf(fmaf(vx, dx, vy * dy)), fabsf(fmaf(vx, bpx, vy * bpy)), fabsf(fmaf(vx, bqx, vy * bqy)),
                        0.5f * fabsf(A.h), delta)) return false;
  }
  return true;
}

// NMS decision "IoU(A, B) > thr" for a pair that survived the early-outs, from the fast estimate alone:
// 1 = exceeds, 0 = does not, -1 = undecided (gate closed, or estimate within kFastBand of thr): take the exact path.
RY_RDEV int rbox_iou_exceeds_quick(const RPrep& A, const RPrep& B, float thr) {
  if (!rbox_fast_ok(A, B)) return -1;
  const float v = rbox_iou_fast(A, B);
  if (fabsf(v - thr) <= kFastBand) return -1;
  return v > thr ? 1 : 0;
}

// The complete decision: bit-identical with Appendix B.  (The NMS kernel runs the two halves as separate passes so
// that the ~10x more expensive exact path executes on compacted, fully populated warps instead of stalling 31 lanes
// whenever one lane of a warp is undecided.)
RY_RDEV bool rbox_iou_exceeds(const RPrep& A, const RPrep& B, float thr) {
  const int q = rbox_iou_exceeds_quick(A, B, thr);
  if (q >= 0) return q != 0;
  return rbox_iou_full(A, B) > thr;
}

// IoU(A, B) with A = the higher-scored ("row") box, B = the candidate ("column") box.
RY_RDEV float rbox_iou(const RPrep& A, const RPrep& B) {
  if (rbox_far(A, B)) return 0.f;
  return rbox_iou_full(A, B);
}

}  // namespace ryolo
