// Anchor/target assignment rule of ComputeCSLLoss.build_targets (lib/loss.py:270-331) and
// ComputeKFIoULoss.build_targets (lib/loss.py:427-492), one candidate entry at a time.
//
// Candidate entries of one pyramid level are enumerated as  e = (o*na + a)*T + t  with
//   o in {centre, x-left j, y-up k, x-right l, y-down m}, a = anchor, t = target row,
// which is exactly the order in which the reference emits its positives (boolean-mask indexing of
// a [5, na*T] tensor).  The arithmetic mirrors the reference's fp32 op sequence; the translation
// unit is compiled with --fmad=false so no multiply-add is contracted (indices are bit-exact).
#pragma once
#include "loss_math.cuh"

namespace ryolo {

struct Pos {          // one positive, 48 bytes
  int b, a, gj, gi;   // image, anchor, grid y, grid x          (lib/loss.py:324)
  float bx, by, bw, bh;  // tbox = (gxy - gij, gwh)             (lib/loss.py:325)
  float ang;          // target angle, radians                  (lib/loss.py:326 / :488)
  int cls;            // target class                           (lib/loss.py:330)
  int row;            // source row in `targets`
  int cell;           // ((b*na + a)*gh + gj)*gw + gi
};

RY_HD float ry_remainder1(float x) {   // torch.remainder(x, 1.0)
  float m = fmodf(x, 1.0f);
  if (m != 0.f && m < 0.f) m += 1.0f;
  return m;
}

// tgt points at (img, cls, x, y, w, h, theta, ...); anc at (w, h[, rad]) in grid units.
RY_HD bool assign_entry(const float* tgt, int o, int a, const float* anc, int rotated, int na, int gh, int gw,
                        int row, int nimg, Pos* out) {
  {  // the reference would raise IndexError on an image index outside the batch; we drop the row
    const int b = (int)tgt[0];
    if (b < 0 || b >= nimg) return false;
  }
  const float gwf = (float)gw, ghf = (float)gh;
  const float tx = tgt[2] * gwf, ty = tgt[3] * ghf, tw = tgt[4] * gwf, th = tgt[5] * ghf;   // :292
  const float rw = tw / anc[0], rh = th / anc[1];                                           // :297
  const float worst = fmaxf(fmaxf(rw, 1.0f / rw), fmaxf(rh, 1.0f / rh));
  if (!(worst < 4.0f)) return false;                                                        // :298
  if (rotated) {
    const float d = fabsf(cosf(tgt[6] - anc[2]));                                           // :458
    if (!(d > 0.866f)) return false;                                                        // :459
  }
  float ox = 0.f, oy = 0.f;
  if (o == 1) {
    if (!((ry_remainder1(tx) < 0.5f) && (tx > 1.0f))) return false;                         // :305
    ox = 0.5f;
  } else if (o == 2) {
    if (!((ry_remainder1(ty) < 0.5f) && (ty > 1.0f))) return false;
    oy = 0.5f;
  } else if (o == 3) {
    const float ix = gwf - tx;                                                              // :304
    if (!((ry_remainder1(ix) < 0.5f) && (ix > 1.0f))) return false;                         // :306
    ox = -0.5f;
  } else if (o == 4) {
    const float iy = ghf - ty;
    if (!((ry_remainder1(iy) < 0.5f) && (iy > 1.0f))) return false;
    oy = -0.5f;
  }
  if (out) {
    long long gi = (long long)(tx - ox), gj = (long long)(ty - oy);                         // :319 trunc
    gi = gi < 0 ? 0 : (gi > gw - 1 ? gw - 1 : gi);                                          // :324 clamp
    gj = gj < 0 ? 0 : (gj > gh - 1 ? gh - 1 : gj);
    out->b = (int)tgt[0];
    out->cls = (int)tgt[1];
    out->a = a;
    out->gi = (int)gi;
    out->gj = (int)gj;
    out->bx = tx - (float)gi;                                                               // :325
    out->by = ty - (float)gj;
    out->bw = tw;
    out->bh = th;
    out->ang = tgt[6];
    out->row = row;
    out->cell = ((out->b * na + a) * gh + (int)gj) * gw + (int)gi;
  }
  return true;
}

}  // namespace ryolo
