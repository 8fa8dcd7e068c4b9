// Rotated-box decode (K6 in SURVEY.md §2.1): replaces the eval branch of
// YoloCSLLayer.forward (model/yololayer.py:28-56) and YoloKFIoULayer.forward (:79-105).
//
// Input  : one pyramid level already in the reference's [B, na, gs, gs, ch] fp32 layout.
// Output : rows [row0, row0 + na*gs*gs) of every image in the fused [B, R, nc+6] tensor,
//          row = a*gs*gs + gy*gs + gx  (yololayer.py:51-52,56).
// One pass over HBM: read ch floats per cell, write nc+6.  fp32 op order mirrors the reference
// (compiled with --fmad=false so that mul/add pairs are not contracted).
#include "common.cuh"

namespace {

constexpr float kPiF = 3.14159274101257324f;  // fl32(np.pi)

__device__ __forceinline__ float sigmoidf_acc(float x) { return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x))); }

// ---- CSL: ch = nc + 185 (x,y,w,h,obj, nc classes, 180 angle bins). One warp per cell, grid-stride: the 180 logits of
// the NEXT cell are requested before the current one is reduced (a warp that handles one cell and exits leaves the
// kernel latency-bound: ncu showed 1.2 TB/s with every warp waiting on its own six loads).
__global__ void __launch_bounds__(256)
decode_csl_kernel(const float* __restrict__ lvl, int64_t cells_total, int na, int gs, int nc, float stride,
                  float aw0, float ah0, float aw1, float ah1, float aw2, float ah2,
                  float* __restrict__ out, int64_t row0, int64_t R) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  int64_t cell = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (cell >= cells_total) return;
  const int ch = nc + 185;
  const int64_t per_img = (int64_t)na * gs * gs;
  float xn[6], hn = 0.f;                              // next cell: angle logits of this lane, head value (lanes 0..4+nc)
  {
    const float* p = lvl + cell * ch;
#pragma unroll
    for (int q = 0; q < 6; q++) { const int j = lane + 32 * q; xn[q] = j < 180 ? p[5 + nc + j] : -INFINITY; }
    if (lane < 5 + nc) hn = p[lane];
  }
  for (; cell < cells_total; cell += nwarps) {
    float xv[6];
#pragma unroll
    for (int q = 0; q < 6; q++) xv[q] = xn[q];
    const float hv = hn;
    const float* p = lvl + cell * ch;
    if (cell + nwarps < cells_total) {
      const float* pn = lvl + (cell + nwarps) * ch;
#pragma unroll
      for (int q = 0; q < 6; q++) { const int j = lane + 32 * q; xn[q] = j < 180 ? pn[5 + nc + j] : -INFINITY; }
      if (lane < 5 + nc) hn = pn[lane];
    }
    int64_t b, r;
    int a, rem;
    if (cells_total < (1ll << 31)) {                  // 32-bit index arithmetic (64-bit divisions cost ~100 instructions)
      const unsigned c32 = (unsigned)cell, pi = (unsigned)per_img, g2 = (unsigned)(gs * gs);
      const unsigned b32 = c32 / pi, r32 = c32 - b32 * pi, a32 = r32 / g2;
      b = b32; r = r32; a = (int)a32; rem = (int)(r32 - a32 * g2);
    } else {
      b = cell / per_img;
      r = cell - b * per_img;
      a = (int)(r / ((int64_t)gs * gs));
      rem = (int)(r - (int64_t)a * gs * gs);
    }
    const int gy = rem / gs, gx = rem - gy * gs;

    // angle bins: argmax over the fp32 SIGMOID values, first maximum wins (yololayer.py:39,48).  The sigmoid is monotone,
    // so only logits that can TIE with the largest one (m) after rounding need their sigmoid evaluated: 180 exponentials
    // + IEEE divisions per cell shrink to (almost always) one.  Two logits can compare equal — or, with expf's 2-ulp
    // error, reversed — only if their true sigmoids are within ~8 ulps: m - x < 8 ulp(s) / s'(m) <= 2^-19 (1 + e^m).
    // Beyond m = 14 the window opens quickly and from 16.64 on every logit maps to 1.0f: the cut is then a flat 11.7
    // (s(11.7) = 1 - 8e-6 is 100 ulps below s(14)).  Flat distributions (random init: all 180 logits within 0.1 of each
    // other) are why the window must be this tight to pay off.
    float m = -INFINITY;
#pragma unroll
    for (int q = 0; q < 6; q++) m = fmaxf(m, xv[q]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const float mc = fminf(m, 14.f);
    const float cut = mc - 1.9073486e-6f * (1.f + __expf(mc));
    float best = -1.f;
    int bi = 0x7fffffff;
#pragma unroll
    for (int q = 0; q < 6; q++) {
      if (xv[q] >= cut) {                              // NaN logits never qualify, exactly like `s > best` below
        const float s = sigmoidf_acc(xv[q]);
        if (s > best) { best = s; bi = lane + 32 * q; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ob = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    // lanes 0..3: box, lane 4: objectness, lanes 5..: classes (the general nc > 27 tail reads its logits directly)
    const float sg = sigmoidf_acc(hv);
    float* o = out + (b * R + row0 + r) * (nc + 6);
    const float aw = a == 0 ? aw0 : (a == 1 ? aw1 : aw2);
    const float ah = a == 0 ? ah0 : (a == 1 ? ah1 : ah2);
    if (lane < 2) {
      o[lane] = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(sg, 2.f), 0.5f), (float)(lane ? gy : gx)), stride);
    } else if (lane < 4) {
      const float w2 = __fmul_rn(sg, 2.f);
      o[lane] = __fmul_rn(__fmul_rn(__fmul_rn(w2, w2), lane == 2 ? aw : ah), stride);
    } else if (lane == 4) {
      o[4] = __fmul_rn(__fdiv_rn((float)(bi - 90), 180.f), kPiF);
      o[5] = sg;
    } else if (lane < 5 + nc) {
      o[1 + lane] = sg;
    }
    for (int c = 27 + lane; c < nc; c += 32) o[6 + c] = sigmoidf_acc(p[5 + c]);
  }
}

// ---- KFIoU: ch = nc + 6 (x,y,w,h,angle,obj, nc classes). One thread per cell.
__global__ void __launch_bounds__(256)
decode_kfiou_kernel(const float* __restrict__ lvl, int64_t cells_total, int na, int gs, int nc, float stride,
                    const float* __restrict__ anchors /* [na,3] device */, float* __restrict__ out,
                    int64_t row0, int64_t R) {
  const int64_t cell = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (cell >= cells_total) return;
  const int ch = nc + 6;
  const float* p = lvl + cell * ch;
  const int64_t per_img = (int64_t)na * gs * gs;
  const int64_t b = cell / per_img;
  const int64_t r = cell - b * per_img;
  const int a = (int)(r / ((int64_t)gs * gs));
  const int rem = (int)(r - (int64_t)a * gs * gs);
  const int gy = rem / gs, gx = rem - gy * gs;
  const float aw = anchors[a * 3 + 0], ah = anchors[a * 3 + 1], aa = anchors[a * 3 + 2];
  float v[8];
  if (ch == 8 && ((((uintptr_t)lvl) & 15) == 0)) {
    float4 lo = ry_ld_stream(reinterpret_cast<const float4*>(p));
    float4 hi = ry_ld_stream(reinterpret_cast<const float4*>(p) + 1);
    v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
  } else {
#pragma unroll
    for (int i = 0; i < 6; i++) v[i] = p[i];
  }
  float* o = out + (b * R + row0 + r) * (nc + 6);
  float sx = sigmoidf_acc(v[0]), sy = sigmoidf_acc(v[1]), sw = sigmoidf_acc(v[2]), sh = sigmoidf_acc(v[3]);
  float sa = sigmoidf_acc(v[4]);
  float r0 = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(sx, 2.f), 0.5f), (float)gx), stride);
  float r1 = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(sy, 2.f), 0.5f), (float)gy), stride);
  float w2 = __fmul_rn(sw, 2.f), h2 = __fmul_rn(sh, 2.f);
  float r2 = __fmul_rn(__fmul_rn(__fmul_rn(w2, w2), aw), stride);
  float r3 = __fmul_rn(__fmul_rn(__fmul_rn(h2, h2), ah), stride);
  float r4 = __fadd_rn(__fmul_rn(__fsub_rn(sa, 0.5f), 0.5236f), aa);  // yololayer.py:96
  float r5 = sigmoidf_acc(v[5]);
  if (ch == 8 && ((((uintptr_t)out) & 15) == 0)) {
    reinterpret_cast<float4*>(o)[0] = make_float4(r0, r1, r2, r3);
    reinterpret_cast<float4*>(o)[1] = make_float4(r4, r5, sigmoidf_acc(v[6]), sigmoidf_acc(v[7]));
  } else {
    o[0] = r0; o[1] = r1; o[2] = r2; o[3] = r3; o[4] = r4; o[5] = r5;
    for (int c = 0; c < nc; c++) o[6 + c] = sigmoidf_acc(p[6 + c]);
  }
}

}  // namespace

extern "C" {

// level: [B,3,gs,gs,nc+185]; anchors_wh: host float[6] in grid units; out: [B,R,nc+6].
int ryolo_decode_csl(const float* level, int64_t B, int gs, int nc, float stride, const float* anchors_wh,
                     float* out, int64_t row0, int64_t R, void* stream) {
  RY_CHECK_ARG(B >= 0 && gs > 0 && nc >= 1, "decode_csl: bad shape");
  const int64_t cells = B * 3 * (int64_t)gs * gs;
  if (cells == 0) return RYOLO_OK;
  RY_CHECK_ARG(row0 + 3ll * gs * gs <= R, "decode_csl: rows exceed output");
  const int64_t threads = cells * 32;
  int64_t blocks = (threads + 255) / 256;
  const int64_t cap = (int64_t)ry_sm_count() * 8 * 4;             // 8 resident blocks per SM, ~4 cells per warp and more
  if (blocks > cap) blocks = cap;
  decode_csl_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      level, cells, 3, gs, nc, stride, anchors_wh[0], anchors_wh[1], anchors_wh[2], anchors_wh[3], anchors_wh[4],
      anchors_wh[5], out, row0, R);
  RY_CHECK_LAUNCH();
  return RYOLO_OK;
}

// level: [B,na,gs,gs,nc+6]; anchors_dev: DEVICE float[na,3] (w,h,rad) in grid units.
int ryolo_decode_kfiou(const float* level, int64_t B, int na, int gs, int nc, float stride,
                       const float* anchors_dev, float* out, int64_t row0, int64_t R, void* stream) {
  RY_CHECK_ARG(B >= 0 && gs > 0 && nc >= 1 && na >= 1, "decode_kfiou: bad shape");
  const int64_t cells = B * (int64_t)na * gs * gs;
  if (cells == 0) return RYOLO_OK;
  RY_CHECK_ARG(row0 + (int64_t)na * gs * gs <= R, "decode_kfiou: rows exceed output");
  decode_kfiou_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      level, cells, na, gs, nc, stride, anchors_dev, out, row0, R);
  RY_CHECK_LAUNCH();
  return RYOLO_OK;
}

}  // extern "C"
