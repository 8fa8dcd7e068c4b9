// ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported, linked or executed by the product path
// (r-yolov4_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library.
//
// CPU restatement of the two detectron2 operators the reference calls but does not vendor:
//   * detectron2.layers.nms.nms_rotated            (reference call site: lib/general.py:4,177)
//   * detectron2.layers.rotated_boxes.pairwise_iou_rotated (reference call site: test.py:7,135)
// detectron2 is an unpinned git-HEAD dependency of the reference (Readme.md:51,
// docker/Dockerfile:33) and is absent from /root/reference and from this image, so this file
// restates its published algorithm (layers/csrc/box_iou_rotated/box_iou_rotated_utils.h and
// layers/csrc/nms_rotated/nms_rotated_cuda.cu) as frozen in SURVEY.md Appendix B.
//
// PARITY UNPINNED for this file: the reference ships no golden vectors / tests at this boundary
// and detectron2 cannot run here.  It is anchored by known-answer tests (tests/test_oracle_*.py)
// and an fp32 cross-check against cv2.rotatedRectangleIntersection.
//
// Frozen spec: CUDA-build semantics — the O(n^2) exchange sort inside the
// Graham scan, suppression on IoU  >  threshold (strict), stable score order (lower index first on
// ties), all arithmetic in fp32 without FMA contraction except the places upstream promotes to
// double (angle->radian conversion + trig, and every comparison against a double literal).
//
// Build: g++ -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (see oracle/Makefile).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>

namespace {

struct P2 {
  float x, y;
};
inline P2 sub(P2 a, P2 b) { return {a.x - b.x, a.y - b.y}; }
inline P2 add(P2 a, P2 b) { return {a.x + b.x, a.y + b.y}; }
inline P2 scale(P2 a, float s) { return {a.x * s, a.y * s}; }
inline float dot2(P2 a, P2 b) { return a.x * b.x + a.y * b.y; }
inline float cross2(P2 a, P2 b) { return a.x * b.y - b.x * a.y; }

struct RBox {
  float cx, cy, w, h, a;
};

// Appendix B step 2: corners; degrees->radians and trig in double, then fp32.
void corners(const RBox& b, P2 (&p)[4]) {
  double theta = b.a * 0.01745329251;
  float c2 = (float)std::cos(theta) * 0.5f;
  float s2 = (float)std::sin(theta) * 0.5f;
  p[0].x = b.cx + s2 * b.h + c2 * b.w;
  p[0].y = b.cy + c2 * b.h - s2 * b.w;
  p[1].x = b.cx - s2 * b.h + c2 * b.w;
  p[1].y = b.cy - c2 * b.h - s2 * b.w;
  p[2].x = 2 * b.cx - p[0].x;
  p[2].y = 2 * b.cy - p[0].y;
  p[3].x = 2 * b.cx - p[1].x;
  p[3].y = 2 * b.cy - p[1].y;
}

// Appendix B step 3: edge/edge crossings + contained corners (<= 24 points, duplicates allowed).
int gather_points(const P2 (&a)[4], const P2 (&b)[4], P2 (&out)[24]) {
  const double EPS = 1e-5;
  P2 ea[4], eb[4];
  for (int i = 0; i < 4; i++) {
    ea[i] = sub(a[(i + 1) & 3], a[i]);
    eb[i] = sub(b[(i + 1) & 3], b[i]);
  }
  int n = 0;
  for (int i = 0; i < 4; i++) {
    for (int j = 0; j < 4; j++) {
      float det = cross2(eb[j], ea[i]);
      if ((double)std::fabs(det) <= 1e-14) continue;  // parallel
      P2 d = sub(b[j], a[i]);
      float t1 = cross2(eb[j], d) / det;
      float t2 = cross2(ea[i], d) / det;
      if ((double)t1 > -EPS && (double)t1 < 1.0f + EPS && (double)t2 > -EPS &&
          (double)t2 < 1.0f + EPS) {
        out[n++] = add(a[i], scale(ea[i], t1));
      }
    }
  }
  {  // corners of a inside b
    P2 AB = eb[0], DA = eb[3];
    float ABAB = dot2(AB, AB), ADAD = dot2(DA, DA);
    for (int i = 0; i < 4; i++) {
      P2 AP = sub(a[i], b[0]);
      float pab = dot2(AP, AB);
      float pad = -dot2(AP, DA);
      if (((double)pab > -EPS) && ((double)pad > -EPS) && ((double)pab < (double)ABAB + EPS) &&
          ((double)pad < (double)ADAD + EPS)) {
        out[n++] = a[i];
      }
    }
  }
  {  // corners of b inside a
    P2 AB = ea[0], DA = ea[3];
    float ABAB = dot2(AB, AB), ADAD = dot2(DA, DA);
    for (int i = 0; i < 4; i++) {
      P2 AP = sub(b[i], a[0]);
      float pab = dot2(AP, AB);
      float pad = -dot2(AP, DA);
      if (((double)pab > -EPS) && ((double)pad > -EPS) && ((double)pab < (double)ABAB + EPS) &&
          ((double)pad < (double)ADAD + EPS)) {
        out[n++] = b[i];
      }
    }
  }
  return n;
}

// Appendix B step 4: Graham scan, CUDA-build flavour (exchange sort), points left shifted to start.
int hull(const P2* p, int n, P2* q) {
  int t = 0;
  for (int i = 1; i < n; i++) {
    if (p[i].y < p[t].y || (p[i].y == p[t].y && p[i].x < p[t].x)) t = i;
  }
  P2 start = p[t];
  for (int i = 0; i < n; i++) q[i] = sub(p[i], start);
  std::swap(q[0], q[t]);
  float dist[24];
  for (int i = 0; i < n; i++) dist[i] = dot2(q[i], q[i]);
  for (int i = 1; i < n - 1; i++) {
    for (int j = i + 1; j < n; j++) {
      float cp = cross2(q[i], q[j]);
      if (((double)cp < -1e-6) || ((double)std::fabs(cp) < 1e-6 && dist[i] > dist[j])) {
        std::swap(q[i], q[j]);
        std::swap(dist[i], dist[j]);
      }
    }
  }
  int k;
  for (k = 1; k < n; k++) {
    if ((double)dist[k] > 1e-8) break;
  }
  if (k == n) {
    q[0] = p[t];
    return 1;
  }
  q[1] = q[k];
  int m = 2;
  for (int i = k + 1; i < n; i++) {
    while (m > 1) {
      P2 q1 = sub(q[i], q[m - 2]), q2 = sub(q[m - 1], q[m - 2]);
      // upstream compares the two rounded products directly (no fused multiply-subtract)
      if (q1.x * q2.y >= q2.x * q1.y)
        m--;
      else
        break;
    }
    q[m++] = q[i];
  }
  return m;
}

float poly_area(const P2* q, int m) {
  if (m <= 2) return 0.f;
  float area = 0.f;
  for (int i = 1; i < m - 1; i++) area += std::fabs(cross2(sub(q[i], q[0]), sub(q[i + 1], q[0])));
  return (float)(area / 2.0);
}

float intersection_area(const RBox& b1, const RBox& b2) {
  P2 pa[4], pb[4], pts[24], ord[24];
  corners(b1, pa);
  corners(b2, pb);
  int n = gather_points(pa, pb, pts);
  if (n <= 2) return 0.f;
  int m = hull(pts, n, ord);
  return poly_area(ord, m);
}

// Appendix B steps 1 + 5.
float iou_one(const float* r1, const float* r2) {
  double sx = (r1[0] + r2[0]) / 2.0;
  double sy = (r1[1] + r2[1]) / 2.0;
  RBox b1{(float)(r1[0] - sx), (float)(r1[1] - sy), r1[2], r1[3], r1[4]};
  RBox b2{(float)(r2[0] - sx), (float)(r2[1] - sy), r2[2], r2[3], r2[4]};
  float a1 = b1.w * b1.h, a2 = b2.w * b2.h;
  if ((double)a1 < 1e-14 || (double)a2 < 1e-14) return 0.f;
  float inter = intersection_area(b1, b2);
  return inter / (a1 + a2 - inter);
}

}  // namespace

extern "C" {

// boxes: [n,5] / [m,5] fp32 (cx, cy, w, h, angle in DEGREES); out: [n,m] fp32.
void oracle_pairwise_iou_rotated(const float* a, int64_t n, const float* b, int64_t m, float* out) {
  for (int64_t i = 0; i < n; i++)
    for (int64_t j = 0; j < m; j++) out[i * m + j] = iou_one(a + 5 * i, b + 5 * j);
}

// Greedy NMS over skew IoU.  Returns number kept; keep[] holds indices into the input, in
// descending-score order.  strict != 0 -> suppress when IoU > thr (CUDA build, the frozen spec);
// strict == 0 -> IoU >= thr (CPU build), kept for documentation of the upstream discrepancy.
int64_t oracle_nms_rotated(const float* boxes, const float* scores, int64_t n, float thr, int strict,
                           int64_t* keep) {
  std::vector<int64_t> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(),
                   [&](int64_t x, int64_t y) { return scores[x] > scores[y]; });
  std::vector<uint8_t> dead(n, 0);
  int64_t nk = 0;
  for (int64_t ii = 0; ii < n; ii++) {
    int64_t i = order[ii];
    if (dead[i]) continue;
    keep[nk++] = i;
    for (int64_t jj = ii + 1; jj < n; jj++) {
      int64_t j = order[jj];
      if (dead[j]) continue;
      float v = iou_one(boxes + 5 * i, boxes + 5 * j);
      if (strict ? (v > thr) : (v >= thr)) dead[j] = 1;
    }
  }
  return nk;
}

// Number of candidate pairs (i<j in score order, both visited by the greedy scan) whose IoU lies
// within `tol` of thr — used by the generators to report how fragile a fixture is.
int64_t oracle_nms_near_threshold(const float* boxes, int64_t n, float thr, float tol) {
  int64_t c = 0;
  for (int64_t i = 0; i < n; i++)
    for (int64_t j = i + 1; j < n; j++) {
      float v = iou_one(boxes + 5 * i, boxes + 5 * j);
      if (std::fabs(v - thr) <= tol) c++;
    }
  return c;
}

}  // extern "C"
