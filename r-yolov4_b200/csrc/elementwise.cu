// HBM-bound companions of the conv stack on NHWC bf16 views (K4, K5 in SURVEY.md §2.1):
//   BatchNorm2d train-mode statistics / finalize / normalise+activation(+residual)   model/utils.py:16-23
//   MaxPool2d (SPP 5/9/13 s1, MaxConv 2x2 s2)                                         model/utils.py:152,231-233
//   nearest x2 upsample into a concat slice                                           model/neck.py:9,19
//   stem im2col (fp32 NCHW image -> bf16 NHWC 27(+37 zero)-channel rows)              model/backbone.py:7
//   weight packing OIHW fp32 -> [Cout][kh][kw][Cin] bf16
// Every thread moves 8 channels (one 128-bit word) so loads/stores are fully vectorised and coalesced.
#include "common.cuh"
#include "ryolo_b200.h"
#include <cuda_bf16.h>

namespace {

__device__ __forceinline__ float act_apply(float x, int act) {
  if (act == RYOLO_ACT_LEAKY) return x > 0.f ? x : 0.1f * x;
  if (act == RYOLO_ACT_MISH) {
    const float e = __expf(fminf(x, 20.f));          // x > 20: n/(n+2) rounds to 1
    const float n = e * (e + 2.f);
    return x * (n * ry_rcp_fma(n + 2.f));            // n + 2 in [2, 2.4e17]
  }
  if (act == RYOLO_ACT_SWISH) return x * ry_rcp_fma(1.f + __expf(fminf(-x, 60.f)));   // denominator in [1, 1.1e26]
  return x;
}

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const float2 t = __bfloat1622float2(b[j]);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}
// same values through integer ops on the packed words: the compiler keeps a prefetched uint4 as four 32-bit registers
// across a loop back-edge (with the __nv_bfloat162 accessors it splits them into 16-bit halves with two PRMTs per word)
__device__ __forceinline__ void unpack8u(const uint4& v, float (&f)[8]) {
  f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
  f[4] = __uint_as_float(v.z << 16); f[5] = __uint_as_float(v.z & 0xffff0000u);
  f[6] = __uint_as_float(v.w << 16); f[7] = __uint_as_float(v.w & 0xffff0000u);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 v;
  __nv_bfloat162* b = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int j = 0; j < 4; j++) b[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
  return v;
}

// ------------------------------------------------------------------------------------ BN statistics
// grid.x blocks stride over pixels; blockDim = (C/8 channel groups) x rows.  sum/sumsq must be zeroed.
__global__ void __launch_bounds__(256)
bn_stats_kernel(const __nv_bfloat16* __restrict__ x, long long pitch, long long P, int C, float* __restrict__ sum,
                float* __restrict__ sumsq) {
  extern __shared__ float red[];   // [rows][C] x 2
  const int groups = C >> 3;
  const int rows = blockDim.x / groups;
  const int g = threadIdx.x % groups, r = threadIdx.x / groups;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (r < rows) {
    for (long long pix = (long long)blockIdx.x * rows + r; pix < P; pix += (long long)gridDim.x * rows) {
      const uint4 v = *reinterpret_cast<const uint4*>(x + pix * pitch + 8 * g);
      float f[8];
      unpack8(v, f);
#pragma unroll
      for (int j = 0; j < 8; j++) { s[j] += f[j]; q[j] += f[j] * f[j]; }
    }
  }
  float* rs = red;
  float* rq = red + (size_t)rows * C;
  if (r < rows) {
#pragma unroll
    for (int j = 0; j < 8; j++) { rs[r * C + 8 * g + j] = s[j]; rq[r * C + 8 * g + j] = q[j]; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int rr = 0; rr < rows; rr++) { a += rs[rr * C + c]; b += rq[rr * C + c]; }
    atomicAdd(sum + c, a);
    atomicAdd(sumsq + c, b);
  }
}

// scale = gamma * rsqrt(var + eps), shift = beta - mean * scale; running stats EMA (unbiased variance)
__global__ void bn_finalize_kernel(const float* __restrict__ sum, const float* __restrict__ sumsq, double count, int C,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, float* __restrict__ running_mean, float* __restrict__ running_var,
                                   long long* __restrict__ num_batches, float* __restrict__ scale,
                                   float* __restrict__ shift, float* __restrict__ save_mean,
                                   float* __restrict__ save_invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && num_batches) *num_batches += 1;
  if (c >= C) return;
  const double mean = (double)sum[c] / count;
  double var = (double)sumsq[c] / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float sc = gamma[c] * invstd;
  scale[c] = sc;
  shift[c] = beta[c] - (float)mean * sc;
  if (save_mean) save_mean[c] = (float)mean;
  if (save_invstd) save_invstd[c] = invstd;
  if (running_mean) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// y = act(x*scale + shift) (+ x2*scale2 + shift2 before the activation) (+ residual after it)
// Thread layout: (C/8 channel groups) x rows; a thread keeps the scale/shift of its 8 channels in registers and
// walks pixels, four independent 128-bit loads in flight.
template <int ACT, bool HAS_X2, bool HAS_RES>
__global__ void __launch_bounds__(256)
scale_shift_act_kernel(const __nv_bfloat16* __restrict__ x, long long xp, const float* __restrict__ scale,
                       const float* __restrict__ shift, const __nv_bfloat16* __restrict__ x2, long long x2p,
                       const float* __restrict__ scale2, const float* __restrict__ shift2,
                       const __nv_bfloat16* __restrict__ res, long long rp, __nv_bfloat16* __restrict__ y, long long yp,
                       long long P, int C) {
  ry_pdl_wait();
  const int groups = C >> 3;
  const int rows = blockDim.x / groups;
  const int g = threadIdx.x % groups, r = threadIdx.x / groups;
  if (r >= rows) return;
  const int c = 8 * g;
  float sc[8], sh[8], sc2[8], sh2[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    sc[j] = scale[c + j];
    sh[j] = shift[c + j];
    if (HAS_X2) { sc2[j] = scale2[c + j]; sh2[j] = shift2[c + j]; }
  }
  const long long stride = (long long)gridDim.x * rows;
  constexpr int U = 4;
  for (long long pix0 = (long long)blockIdx.x * rows + r; pix0 < P; pix0 += U * stride) {
    uint4 vx[U], v2[U], vr[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const long long pix = pix0 + u * stride;
      if (pix < P) {
        vx[u] = *reinterpret_cast<const uint4*>(x + pix * xp + c);
        if (HAS_X2) v2[u] = *reinterpret_cast<const uint4*>(x2 + pix * x2p + c);
        if (HAS_RES) vr[u] = *reinterpret_cast<const uint4*>(res + pix * rp + c);
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const long long pix = pix0 + u * stride;
      if (pix < P) {
        float f[8];
        unpack8(vx[u], f);
#pragma unroll
        for (int j = 0; j < 8; j++) f[j] = f[j] * sc[j] + sh[j];
        if (HAS_X2) {
          float h[8];
          unpack8(v2[u], h);
#pragma unroll
          for (int j = 0; j < 8; j++) f[j] += h[j] * sc2[j] + sh2[j];
        }
#pragma unroll
        for (int j = 0; j < 8; j++) f[j] = act_apply(f[j], ACT);
        if (HAS_RES) {
          float h[8];
          unpack8(vr[u], h);
#pragma unroll
          for (int j = 0; j < 8; j++) f[j] += h[j];
        }
        *reinterpret_cast<uint4*>(y + pix * yp + c) = pack8(f);
      }
    }
  }
}

// Same pass, latency-proofed (knob ssa = 1, default): the grid is exactly the resident capacity (2 blocks of 256 threads
// per SM, grid-stride), a thread requests its next U rows BEFORE it processes the current U (register double buffer:
// every resident warp keeps U x (1-3 operands) 128-bit loads in flight through its compute phase), the row walk is
// pointer increments with a 32-bit trip count, and Mish takes its reciprocal from the SFU (10 instructions; with the
// FMA-pipe reciprocal the kernel above is issue-bound: 23 instructions per element against a budget of 22 at 4 B/element).
// The kernel above issues four loads, waits, computes ~250 instructions with nothing in flight, and runs 2-3 waves of
// short blocks (ncu: 3.8-3.9 TB/s, long-scoreboard + wait stalls).
__device__ __forceinline__ float ssa_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ssa_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <int ACT>
__device__ __forceinline__ float act_apply_sfu(float x) {
  if (ACT == RYOLO_ACT_LEAKY) return x > 0.f ? x : 0.1f * x;
  if (ACT == RYOLO_ACT_MISH) {           // x * n / (n + 2), n = e (e + 2); x > 20: the quotient rounds to 1
    const float e = ssa_ex2(fminf(x * 1.4426950408889634f, 28.853900817779268f));
    const float n = e * (e + 2.f);
    return x * (n * ssa_rcp(n + 2.f));
  }
  if (ACT == RYOLO_ACT_SWISH) return x * ssa_rcp(1.f + ssa_ex2(fminf(x * -1.4426950408889634f, 86.f)));
  return x;
}

template <int ACT, bool HAS_X2, bool HAS_RES>
__device__ __forceinline__ uint4 ssa_row(const uint4& vx, const uint4& v2, const uint4& vr, const float (&sc)[8],
                                         const float (&sh)[8], const float (&sc2)[8]) {
  float f[8];
  unpack8u(vx, f);
#pragma unroll
  for (int j = 0; j < 8; j++) f[j] = fmaf(f[j], sc[j], sh[j]);
  if (HAS_X2) {
    float h[8];
    unpack8u(v2, h);
#pragma unroll
    for (int j = 0; j < 8; j++) f[j] = fmaf(h[j], sc2[j], f[j]);
  }
#pragma unroll
  for (int j = 0; j < 8; j++) f[j] = act_apply_sfu<ACT>(f[j]);
  if (HAS_RES) {
    float h[8];
    unpack8u(vr, h);
#pragma unroll
    for (int j = 0; j < 8; j++) f[j] += h[j];
  }
  return pack8(f);
}

template <int ACT, bool HAS_X2, bool HAS_RES>
__device__ __forceinline__ void ssa_walk(const __nv_bfloat16* __restrict__ x, long long xp,
                                         const __nv_bfloat16* __restrict__ x2, long long x2p,
                                         const __nv_bfloat16* __restrict__ res, long long rp,
                                         __nv_bfloat16* __restrict__ y, long long yp, long long P, int c, int r, int rows,
                                         const float (&sc)[8], const float (&sh)[8], const float (&sc2)[8]) {
  constexpr int U = (HAS_X2 || HAS_RES) ? 2 : 4;
  const int stride = (int)gridDim.x * rows;
  const int pix0 = (int)blockIdx.x * rows + r;
  const int n = pix0 < P ? (int)((P - 1 - pix0) / stride) + 1 : 0;      // rows of this thread
  const long long sx = (long long)stride * xp, s2 = (long long)stride * x2p, sr = (long long)stride * rp,
                  sy = (long long)stride * yp;
  const __nv_bfloat16* px = x + c + (long long)pix0 * xp;
  const __nv_bfloat16* p2 = HAS_X2 ? x2 + c + (long long)pix0 * x2p : nullptr;
  const __nv_bfloat16* pr = HAS_RES ? res + c + (long long)pix0 * rp : nullptr;
  __nv_bfloat16* py = y + c + (long long)pix0 * yp;
  const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
  const int trips = n / U;
  if (trips) {
    uint4 vx[U], v2[U], vr[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      vx[u] = *reinterpret_cast<const uint4*>(px + u * sx);
      v2[u] = HAS_X2 ? *reinterpret_cast<const uint4*>(p2 + u * s2) : z4;
      vr[u] = HAS_RES ? *reinterpret_cast<const uint4*>(pr + u * sr) : z4;
    }
#pragma unroll 2
    for (int i = 1; i < trips; i++) {
      px += U * sx;
      if (HAS_X2) p2 += U * s2;
      if (HAS_RES) pr += U * sr;
      uint4 nx[U], n2[U], nr[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        nx[u] = *reinterpret_cast<const uint4*>(px + u * sx);
        n2[u] = HAS_X2 ? *reinterpret_cast<const uint4*>(p2 + u * s2) : z4;
        nr[u] = HAS_RES ? *reinterpret_cast<const uint4*>(pr + u * sr) : z4;
      }
#pragma unroll
      for (int u = 0; u < U; u++)
        *reinterpret_cast<uint4*>(py + u * sy) = ssa_row<ACT, HAS_X2, HAS_RES>(vx[u], v2[u], vr[u], sc, sh, sc2);
      py += U * sy;
#pragma unroll
      for (int u = 0; u < U; u++) { vx[u] = nx[u]; v2[u] = n2[u]; vr[u] = nr[u]; }
    }
#pragma unroll
    for (int u = 0; u < U; u++)
      *reinterpret_cast<uint4*>(py + u * sy) = ssa_row<ACT, HAS_X2, HAS_RES>(vx[u], v2[u], vr[u], sc, sh, sc2);
    px += U * sx;
    if (HAS_X2) p2 += U * s2;
    if (HAS_RES) pr += U * sr;
    py += U * sy;
  }
  for (int i = trips * U; i < n; i++) {                      // at most U - 1 left-over rows
    const uint4 a = *reinterpret_cast<const uint4*>(px);
    const uint4 b = HAS_X2 ? *reinterpret_cast<const uint4*>(p2) : z4;
    const uint4 d = HAS_RES ? *reinterpret_cast<const uint4*>(pr) : z4;
    *reinterpret_cast<uint4*>(py) = ssa_row<ACT, HAS_X2, HAS_RES>(a, b, d, sc, sh, sc2);
    px += sx;
    if (HAS_X2) p2 += s2;
    if (HAS_RES) pr += sr;
    py += sy;
  }
}

template <int ACT, bool HAS_X2, bool HAS_RES>
__global__ void __launch_bounds__(256, 2)
scale_shift_act_p_kernel(const __nv_bfloat16* __restrict__ x, long long xp, const float* __restrict__ scale,
                         const float* __restrict__ shift, const __nv_bfloat16* __restrict__ x2, long long x2p,
                         const float* __restrict__ scale2, const float* __restrict__ shift2,
                         const __nv_bfloat16* __restrict__ res, long long rp, __nv_bfloat16* __restrict__ y, long long yp,
                         long long P, int C) {
  ry_pdl_wait();
  const int groups = C >> 3;
  const int rows = blockDim.x / groups;
  const int g = threadIdx.x % groups, r = threadIdx.x / groups;
  if (r >= rows) return;
  const int c = 8 * g;
  float sc[8], sh[8], sc2[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    sc[j] = scale[c + j];
    sh[j] = shift[c + j];
    sc2[j] = 0.f;
    if (HAS_X2) { sc2[j] = scale2[c + j]; sh[j] += shift2[c + j]; }
  }
  ssa_walk<ACT, HAS_X2, HAS_RES>(x, xp, x2, x2p, res, rp, y, yp, P, c, r, rows, sc, sh, sc2);
}

// The same pass with the train-mode BatchNorm FINALIZE folded in (ryolo_scale_shift_act_bn): the producing conv only
// accumulated its per-channel fixed-point totals (ryolo_bn_fuse with counter == NULL); every thread here turns the
// totals of its 8 channels into scale / shift (the same double-precision arithmetic as the conv kernel's last-CTA pass,
// bit for bit), and block 0 publishes scale / shift / mean / invstd for the backward pass and moves the running
// statistics.  Removes the fence + arrival counter + last-CTA pass from every conv launch's tail.
struct SsaBn {
  const unsigned long long* acc;      // [2][C] fixed-point sum | sum of squares
  const float* gamma; const float* beta;
  float* running_mean; float* running_var; long long* num_batches;
  float eps, momentum;
  double k_sum, k_sq, k_unbias;
  float* scale; float* shift; float* save_mean; float* save_invstd;
};

template <int ACT, bool HAS_RES>
__global__ void __launch_bounds__(256, 2)
scale_shift_act_bn_kernel(const __nv_bfloat16* __restrict__ x, long long xp, const SsaBn bn,
                          const __nv_bfloat16* __restrict__ res, long long rp, __nv_bfloat16* __restrict__ y, long long yp,
                          long long P, int C) {
  ry_pdl_wait();
  extern __shared__ float s_fin[];               // [2][C]: scale | shift of this launch, computed once per block
  const int groups = C >> 3;
  const int rows = blockDim.x / groups;
  const int g = threadIdx.x % groups, r = threadIdx.x / groups;
  if (blockIdx.x == 0 && threadIdx.x == 0 && bn.num_batches) *bn.num_batches += 1;
  // One channel per thread (the first version had every thread finalize its own 8 channels: 2048 double-precision
  // finalizations per block instead of C, ~10 us per launch).  The cancellation-prone part (E[x^2] - mean^2 on exact
  // integer totals) is double, the reciprocal square root rsqrt.approx + one Newton step in fp32.
  for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
    const long long is = (long long)__ldcg(bn.acc + ch), iq = (long long)__ldcg(bn.acc + C + ch);
    const double mean = (double)is * bn.k_sum;                    // k_sum = 1 / (2^24 count)
    double var = fma((double)iq, bn.k_sq, -mean * mean);          // k_sq  = 1 / (2^20 count)
    if (var < 0.0) var = 0.0;
    const float vf = (float)(var + (double)bn.eps);
    float invstd = rsqrtf(vf);
    invstd = invstd * fmaf(-0.5f * vf * invstd, invstd, 1.5f);
    const float s = bn.gamma[ch] * invstd;
    const float t = bn.beta[ch] - (float)mean * s;
    s_fin[ch] = s;
    s_fin[C + ch] = t;
    if (blockIdx.x == 0) {                       // publish for the backward pass / eval mode
      bn.scale[ch] = s;
      bn.shift[ch] = t;
      if (bn.save_mean) bn.save_mean[ch] = (float)mean;
      if (bn.save_invstd) bn.save_invstd[ch] = invstd;
      if (bn.running_mean) {
        const double unbiased = var * bn.k_unbias;                // count / (count - 1), or 1
        bn.running_mean[ch] = (1.f - bn.momentum) * bn.running_mean[ch] + bn.momentum * (float)mean;
        bn.running_var[ch] = (1.f - bn.momentum) * bn.running_var[ch] + bn.momentum * (float)unbiased;
      }
    }
  }
  __syncthreads();
  if (r >= rows) return;
  const int c = 8 * g;
  float sc[8], sh[8], sc2[8];
#pragma unroll
  for (int j = 0; j < 8; j++) { sc[j] = s_fin[c + j]; sh[j] = s_fin[C + c + j]; sc2[j] = 0.f; }
  ssa_walk<ACT, false, HAS_RES>(x, xp, nullptr, 0, res, rp, y, yp, P, c, r, rows, sc, sh, sc2);
}

typedef void (*SsaFn)(const __nv_bfloat16*, long long, const float*, const float*, const __nv_bfloat16*, long long,
                      const float*, const float*, const __nv_bfloat16*, long long, __nv_bfloat16*, long long, long long,
                      int);

template <int ACT>
SsaFn ssa_pick(bool x2, bool res, bool pipelined) {
  if (pipelined) {
    if (x2) return res ? scale_shift_act_p_kernel<ACT, true, true> : scale_shift_act_p_kernel<ACT, true, false>;
    return res ? scale_shift_act_p_kernel<ACT, false, true> : scale_shift_act_p_kernel<ACT, false, false>;
  }
  if (x2) return res ? scale_shift_act_kernel<ACT, true, true> : scale_shift_act_kernel<ACT, true, false>;
  return res ? scale_shift_act_kernel<ACT, false, true> : scale_shift_act_kernel<ACT, false, false>;
}

// ------------------------------------------------------------------------------------ pooling / resize / copy
__global__ void __launch_bounds__(256)
maxpool_kernel(const __nv_bfloat16* __restrict__ x, long long xp, int N, int H, int W, int C, int k, int stride,
               int pad, __nv_bfloat16* __restrict__ y, long long yp, int Ho, int Wo) {
  const int groups = C >> 3;
  const long long total = (long long)N * Ho * Wo * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % groups) * 8;
    long long pix = i / groups;
    const int wo = (int)(pix % Wo), ho = (int)((pix / Wo) % Ho), n = (int)(pix / ((long long)Wo * Ho));
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; j++) m[j] = -INFINITY;
    for (int dh = 0; dh < k; dh++) {
      const int hi = ho * stride + dh - pad;
      if (hi < 0 || hi >= H) continue;
      for (int dw = 0; dw < k; dw++) {
        const int wi = wo * stride + dw - pad;
        if (wi < 0 || wi >= W) continue;
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(x + (((long long)n * H + hi) * W + wi) * xp + c), f);
#pragma unroll
        for (int j = 0; j < 8; j++) m[j] = fmaxf(m[j], f[j]);
      }
    }
    *reinterpret_cast<uint4*>(y + pix * yp + c) = pack8(m);
  }
}

// Stride-1 "same" max pool on a small map (SPP: k = 5 / 9 / 13 on 25x25): one block per (image, 8 channels), the map
// lives in shared memory and the window maximum is separable (row maxima, then a column of row maxima): 2k instead
// of k*k comparisons per output.
__device__ __forceinline__ uint4 max8_bf16(const uint4& a, const uint4& b) {
  uint4 r;
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
  __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int j = 0; j < 4; j++) pr[j] = __hmax2(pa[j], pb[j]);
  return r;
}

__global__ void __launch_bounds__(256)
maxpool_same_small_kernel(const __nv_bfloat16* __restrict__ x, long long xp, int H, int W, int C, int k,
                          __nv_bfloat16* __restrict__ y, long long yp) {
  extern __shared__ uint4 sm_pool[];           // [H*W] map | [H*W] row maxima
  uint4* sx = sm_pool;
  uint4* sr = sm_pool + H * W;
  const int groups = C >> 3;
  const int n = blockIdx.x / groups, c = (blockIdx.x % groups) * 8;
  const int HW = H * W, p = k >> 1;
  const __nv_bfloat16* xb = x + (long long)n * HW * xp + c;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) sx[i] = *reinterpret_cast<const uint4*>(xb + (long long)i * xp);
  __syncthreads();
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const int h = i / W, w = i - h * W;
    const int w0 = max(w - p, 0), w1 = min(w + p, W - 1);
    uint4 m = sx[h * W + w0];
    for (int ww = w0 + 1; ww <= w1; ww++) m = max8_bf16(m, sx[h * W + ww]);
    sr[i] = m;
  }
  __syncthreads();
  __nv_bfloat16* yb = y + (long long)n * HW * yp + c;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const int h = i / W, w = i - h * W;
    const int h0 = max(h - p, 0), h1 = min(h + p, H - 1);
    uint4 m = sr[h0 * W + w];
    for (int hh = h0 + 1; hh <= h1; hh++) m = max8_bf16(m, sr[hh * W + w]);
    *reinterpret_cast<uint4*>(yb + (long long)i * yp) = m;
  }
}

// nearest-neighbour x`f` upsample (f = 1 is a plain strided copy between views)
__global__ void __launch_bounds__(256)
resize_copy_kernel(const __nv_bfloat16* __restrict__ x, long long xp, int N, int H, int W, int C, int f,
                   __nv_bfloat16* __restrict__ y, long long yp) {
  const int groups = C >> 3, Ho = H * f, Wo = W * f;
  const long long total = (long long)N * Ho * Wo * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % groups) * 8;
    long long pix = i / groups;
    const int wo = (int)(pix % Wo), ho = (int)((pix / Wo) % Ho), n = (int)(pix / ((long long)Wo * Ho));
    const uint4 v = *reinterpret_cast<const uint4*>(x + (((long long)n * H + ho / f) * W + wo / f) * xp + c);
    *reinterpret_cast<uint4*>(y + pix * yp + c) = v;
  }
}

// ------------------------------------------------------------------------------------ stem im2col
// img fp32 NCHW [N,3,H,W] -> bf16 [N,Ho,Wo,Kpad]: channel (kh*k+kw)*3 + ci holds img[ci, ho*s+kh-pad, wo*s+kw-pad]
// (zero padded), channels 3*k*k..Kpad-1 are zero.  The 3-channel stem (3x3/s1 of yolov4/v7, 6x6/s2 of yolov5) then
// runs as a K=Kpad 1x1 conv on the tensor cores.  One thread per (pixel, 8 channels).
__global__ void __launch_bounds__(256)
stem_im2col_kernel(const float* __restrict__ img, int N, int H, int W, int k, int s, int pad, int Kpad, int Ho, int Wo,
                   __nv_bfloat16* __restrict__ y) {
  const int groups = Kpad >> 3;
  const long long total = (long long)N * Ho * Wo * groups;
  const int kreal = 3 * k * k;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    const long long pix = i / groups;
    const int wo = (int)(pix % Wo), ho = (int)((pix / Wo) % Ho), n = (int)(pix / ((long long)Wo * Ho));
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int kk = 8 * g + j;
      float v = 0.f;
      if (kk < kreal) {
        const int tap = kk / 3, ci = kk - 3 * tap;
        const int kh = tap / k, kw = tap - kh * k;
        const int hi = ho * s + kh - pad, wi = wo * s + kw - pad;
        if (hi >= 0 && hi < H && wi >= 0 && wi < W) v = __ldg(img + (((long long)n * 3 + ci) * H + hi) * W + wi);
      }
      f[j] = v;
    }
    *reinterpret_cast<uint4*>(y + pix * Kpad + 8 * g) = pack8(f);
  }
}

// 3x3 / stride-1 stem (yolov4, yolov7), Kpad = 32: one thread per pixel gathers its 27 inputs (coalesced along W for
// every (tap, channel), 9x reuse out of L1) and writes one 64-byte row.  The generic kernel above spends its time in
// per-element integer divisions.
__global__ void __launch_bounds__(256)
stem_im2col_k3_kernel(const float* __restrict__ img, int N, int H, int W, __nv_bfloat16* __restrict__ y) {
  const long long total = (long long)N * H * W;
  const long long plane = (long long)H * W;
  for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < total;
       pix += (long long)gridDim.x * blockDim.x) {
    const int wo = (int)(pix % W), ho = (int)((pix / W) % H);
    const long long n = pix / plane;
    const float* base = img + n * 3 * plane;
    float f[32];
#pragma unroll
    for (int kh = 0; kh < 3; kh++) {
      const int hi = ho + kh - 1;
      const bool hok = hi >= 0 && hi < H;
#pragma unroll
      for (int kw = 0; kw < 3; kw++) {
        const int wi = wo + kw - 1;
        const bool ok = hok && wi >= 0 && wi < W;
#pragma unroll
        for (int ci = 0; ci < 3; ci++)
          f[(kh * 3 + kw) * 3 + ci] = ok ? __ldg(base + ci * plane + (long long)hi * W + wi) : 0.f;
      }
    }
#pragma unroll
    for (int j = 27; j < 32; j++) f[j] = 0.f;
    uint4* o = reinterpret_cast<uint4*>(y + pix * 32);
#pragma unroll
    for (int g = 0; g < 4; g++) {
      float t[8];
#pragma unroll
      for (int j = 0; j < 8; j++) t[j] = f[8 * g + j];
      o[g] = pack8(t);
    }
  }
}

// ------------------------------------------------------------------------------------ weight packing
// OIHW fp32 -> [Cout][kh][kw][Cin] bf16 (stem == 0).  stem == 2 (transpose, for dgrad): -> [Cin][kh][kw][Cout].
// stem >= 8: the 3-channel stem, [Cout,3,k,k] -> [Cout][Kpad = stem] in the im2col channel order (zero padded).
__global__ void pack_weights_kernel(const float* __restrict__ w, int Cout, int Cin, int k, int stem,
                                    __nv_bfloat16* __restrict__ out) {
  if (stem == 2) {
    const long long total = (long long)Cin * k * k * Cout;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
      const int co = (int)(i % Cout);
      const int tap = (int)((i / Cout) % (k * k));
      const int ci = (int)(i / ((long long)Cout * k * k));
      out[i] = __float2bfloat16_rn(w[((long long)co * Cin + ci) * k * k + tap]);
    }
    return;
  }
  const int Kp = stem >= 8 ? stem : k * k * Cin;
  const long long total = (long long)Cout * Kp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i / Kp), kk = (int)(i % Kp);
    float v = 0.f;
    if (stem >= 8) {
      if (kk < 3 * k * k) {
        const int tap = kk / 3, ci = kk % 3;
        v = w[((long long)co * 3 + ci) * k * k + tap];
      }
    } else {
      const int tap = kk / Cin, ci = kk % Cin;
      v = w[((long long)co * Cin + ci) * k * k + tap];
    }
    out[i] = __float2bfloat16_rn(v);
  }
}

// All conv weights of a model in ONE launch: table[i] describes tensor i.  Thread t owns elements [8t, 8t+8) of the
// launch's flat element space (every tensor's element count is a multiple of 8) and writes ONE 16-byte piece of each
// destination layout: 8 consecutive Cin of dst[co][tap][.] and 8 consecutive Cout of dst_t[ci][tap][.].  Lanes step
// through (tap, ci) fastest, so a warp's strided fp32 reads cover whole sectors between them.
__device__ __forceinline__ uint4 pack8_bf16(const float (&f)[8]) {
  uint4 v;
  __nv_bfloat162* b = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int j = 0; j < 4; j++) b[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
  return v;
}

__global__ void __launch_bounds__(256)
pack_weights_multi_kernel(const ryolo_pack_entry* __restrict__ table, int n, long long total) {
  const long long chunks = total >> 3;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < chunks;
       t += (long long)gridDim.x * blockDim.x) {
    const long long i = t << 3;
    int lo = 0, hi = n - 1;                       // last entry whose first element <= i
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (table[mid].first <= i) lo = mid; else hi = mid - 1;
    }
    const ryolo_pack_entry e = table[lo];
    const long long r0 = i - e.first;             // first of this thread's 8 elements (OIHW order for the fallback)
    const int kk = e.k * e.k;
    __nv_bfloat16* dst = (__nv_bfloat16*)e.dst;
    __nv_bfloat16* dst_t = (__nv_bfloat16*)e.dst_t;
    if (e.stem || (e.Cin & 7)) {                  // 3-channel stem: element-wise into the im2col order
      for (int j = 0; j < 8; j++) {
        const long long r = r0 + j;               // OIHW index: ((co*Cin + ci)*k*k + tap)
        const int tap = (int)(r % kk);
        const int ci = (int)((r / kk) % e.Cin);
        const int co = (int)(r / ((long long)kk * e.Cin));
        const __nv_bfloat16 v = __float2bfloat16_rn(e.src[r]);
        if (e.stem) dst[(long long)co * e.stem + tap * 3 + ci] = v;       // [Cout][Kpad] (pad pre-zeroed)
        else {
          dst[((long long)co * kk + tap) * e.Cin + ci] = v;
          if (dst_t) dst_t[((long long)ci * kk + tap) * e.Cout + co] = v;
        }
      }
      continue;
    }
    const long long q0 = r0 >> 3;                 // chunk index inside the tensor
    float f[8];
    {   // dst = [Cout][kh][kw][Cin]: chunk -> (co, ci chunk, tap), tap fastest
      long long q = q0;
      const int tap = (int)(q % kk); q /= kk;
      const int cch = e.Cin >> 3;
      const int cc = (int)(q % cch);
      const int co = (int)(q / cch);
      const float* sp = e.src + ((long long)co * e.Cin + 8 * cc) * kk + tap;
#pragma unroll
      for (int j = 0; j < 8; j++) f[j] = sp[(long long)j * kk];
      *reinterpret_cast<uint4*>(dst + ((long long)co * kk + tap) * e.Cin + 8 * cc) = pack8_bf16(f);
    }
    if (dst_t) {
      if ((e.Cout & 7) == 0) {   // dst_t = [Cin][kh][kw][Cout]: chunk -> (co chunk, ci, tap), tap fastest
        long long q = q0;
        const int tap = (int)(q % kk); q /= kk;
        const int ci = (int)(q % e.Cin);
        const int coc = (int)(q / e.Cin);
        const float* sp = e.src + ((long long)(8 * coc) * e.Cin + ci) * kk + tap;
        const long long cs = (long long)e.Cin * kk;
#pragma unroll
        for (int j = 0; j < 8; j++) f[j] = sp[j * cs];
        *reinterpret_cast<uint4*>(dst_t + ((long long)ci * kk + tap) * e.Cout + 8 * coc) = pack8_bf16(f);
      } else {
        for (int j = 0; j < 8; j++) {
          const long long r = r0 + j;
          const int tap = (int)(r % kk);
          const int ci = (int)((r / kk) % e.Cin);
          const int co = (int)(r / ((long long)kk * e.Cin));
          dst_t[((long long)ci * kk + tap) * e.Cout + co] = __float2bfloat16_rn(e.src[r]);
        }
      }
    }
  }
}

// Reverse of pack_weights_multi for gradients: grad_oihw[first + r] += dwk[K-major index of r]  (dst = OIHW fp32
// gradient, src = K-major fp32 scratch written by ryolo_conv2d_wgrad; stem: src is [Cout][Kpad] im2col order).
__global__ void __launch_bounds__(256)
unpack_wgrad_multi_kernel(const ryolo_pack_entry* __restrict__ table, int n, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (table[mid].first <= i) lo = mid; else hi = mid - 1;
    }
    const ryolo_pack_entry e = table[lo];
    const long long r = i - e.first;
    const int kk = e.k * e.k;
    const int tap = (int)(r % kk);
    const int ci = (int)((r / kk) % e.Cin);
    const int co = (int)(r / ((long long)kk * e.Cin));
    const float* src = e.src;
    const float v = e.stem ? src[(long long)co * e.stem + tap * 3 + ci] : src[((long long)co * kk + tap) * e.Cin + ci];
    ((float*)e.dst)[r] += v;
  }
}

// The same fold, tiled, and it CLEARS the scratch it has folded (the next step's wgrad atomics start from zero without a
// separate 0.6 GB fill): blockIdx.y = table entry, a warp owns (one output channel, 32 input channels).  It reads the
// k*k K-major rows of its 32 channels (k*k coalesced 128-byte requests), turns them around in a [32][k*k] shared tile
// (odd k*k: conflict free) and adds them to the 32*k*k CONTIGUOUS floats of the OIHW gradient.  The element-per-thread
// kernel above reads with a stride of Cin floats: 32 sectors per request, 0.50 ms per yolov4 step at 1.5 TB/s.
__global__ void __launch_bounds__(256)
unpack_wgrad_tiled_kernel(const ryolo_pack_entry* __restrict__ table, int n) {
  __shared__ float tile[8][32 * 9 + 1];
  const ryolo_pack_entry e = table[blockIdx.y];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kk = e.k * e.k;
  float* dst = (float*)e.dst;
  if (e.stem || kk > 9) {                        // 3-channel stem (im2col order) and anything unusual: element-wise
    const long long cnt = (long long)e.Cout * e.Cin * kk;
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < cnt; r += (long long)gridDim.x * blockDim.x) {
      const int tap = (int)(r % kk);
      const int ci = (int)((r / kk) % e.Cin);
      const int co = (int)(r / ((long long)kk * e.Cin));
      float* sp = const_cast<float*>(e.src) + (e.stem ? (long long)co * e.stem + tap * 3 + ci : ((long long)co * kk + tap) * e.Cin + ci);
      dst[r] += *sp;
      *sp = 0.f;
    }
    return;
  }
  const int chunks = (e.Cin + 31) >> 5;
  const long long items = (long long)e.Cout * chunks;
  float* t = tile[warp];
  for (long long it = (long long)blockIdx.x * 8 + warp; it < items; it += (long long)gridDim.x * 8) {
    const int co = (int)(it / chunks), c0 = (int)(it - (long long)co * chunks) * 32;
    const int nv = min(32, e.Cin - c0);
    const float* sp = e.src + (long long)co * kk * e.Cin + c0 + lane;
    if (lane < nv) {
      float* spw = const_cast<float*>(sp);          // fold AND clear: the scratch is zero again for the next step's atomics
      float v[9];                                   // all loads first: a store to the same pointer between them would
#pragma unroll                                      // serialise the k*k round trips (measured: 0.39 -> 3.6 ms)
      for (int tap = 0; tap < 9; tap++) v[tap] = tap < kk ? sp[(long long)tap * e.Cin] : 0.f;
#pragma unroll
      for (int tap = 0; tap < 9; tap++)
        if (tap < kk) { t[lane * kk + tap] = v[tap]; spw[(long long)tap * e.Cin] = 0.f; }
    }
    __syncwarp();
    float* dp = dst + ((long long)co * e.Cin + c0) * kk;
    for (int j = lane; j < nv * kk; j += 32) dp[j] += t[j];
    __syncwarp();
  }
}

inline int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = 148ll * 32;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" {

int ryolo_bn_stats(const void* x, long long pitch, long long P, int C, float* sum, float* sumsq, void* stream) {
  RY_CHECK_ARG(C % 8 == 0 && C >= 8 && C <= 2048 && pitch % 8 == 0, "bn_stats: C must be a multiple of 8 in [8, 2048]");
  if (P == 0) return RYOLO_OK;
  const int groups = C / 8;
  const int threads = groups >= 256 ? groups : 256;       // C > 2048 excluded above -> threads <= 256
  const int rows = threads / groups;
  const size_t smem = (size_t)2 * rows * C * sizeof(float);
  long long want = (P + rows - 1) / rows;
  int blocks = (int)(want > 148 * 8 ? 148 * 8 : want);
  bn_stats_kernel<<<blocks, threads, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, pitch, P, C, sum, sumsq);
  RY_CHECK_LAUNCH();
  return RYOLO_OK;
}

int ryolo_bn_finalize(const float* sum, const float* sumsq, double count, int C, const float* gamma, const float* beta,
                      float eps, float momentum, float* running_mean, float* running_var, long long* num_batches,
                      float* scale, float* shift, float* save_mean, float* save_invstd, void* stream) {
  RY_CHECK_ARG(C > 0 && count > 0, "bn_finalize: bad shape");
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sum, sumsq, count, C, gamma, beta, eps, momentum,
                                                                        running_mean, running_var, num_batches, scale,
                                                                        shift, save_mean, save_invstd);
  RY_CHECK_LAUNCH();
  return RYOLO_OK;
}

int ryolo_scale_shift_act(const void* x, long long xp, const float* scale, const float* shift, const void* x2,
                          long long x2p, const float* scale2, const float* shift2, int act, const void* residual,
                          long long rp, void* y, long long yp, long long P, int C, void* stream) {
  RY_CHECK_ARG(C % 8 == 0 && C >= 8 && C <= 2048 && xp % 8 == 0 && yp % 8 == 0,
               "scale_shift_act: channels must be a multiple of 8 in [8, 2048]");
  RY_CHECK_ARG(scale && shift && (!x2 || (scale2 && shift2)), "scale_shift_act: missing scale/shift");
  RY_CHECK_ARG(P < (1ll << 31) - (1 << 20), "scale_shift_act: more than 2^31 pixels");
  if (P == 0) return RYOLO_OK;
  const int groups = C / 8;
  const int threads = groups >= 256 ? groups : 256;
  const int rows = threads / groups;
  const bool pipelined = threads == 256 && ryolo_knob(RYOLO_KNOB_SSA) != 0;
  SsaFn fn;
  switch (act) {
    case RYOLO_ACT_LEAKY: fn = ssa_pick<RYOLO_ACT_LEAKY>(x2 != nullptr, residual != nullptr, pipelined); break;
    case RYOLO_ACT_MISH: fn = ssa_pick<RYOLO_ACT_MISH>(x2 != nullptr, residual != nullptr, pipelined); break;
    case RYOLO_ACT_SWISH: fn = ssa_pick<RYOLO_ACT_SWISH>(x2 != nullptr, residual != nullptr, pipelined); break;
    default: fn = ssa_pick<RYOLO_ACT_LINEAR>(x2 != nullptr, residual != nullptr, pipelined); break;
  }
  const int per_trip = pipelined ? ((x2 || residual) ? 2 : 4) : 4;
  long long want = (P + (long long)rows * per_trip - 1) / ((long long)rows * per_trip);
  const long long cap = pipelined ? 2ll * ry_sm_count() : 148 * 16;
  const int blocks = (int)(want > cap ? cap : (want < 1 ? 1 : want));
  ry_launch(fn, dim3(blocks), dim3(threads), 0, (cudaStream_t)stream, (const __nv_bfloat16*)x, xp, scale, shift,
            (const __nv_bfloat16*)x2, x2p, scale2, shift2, (const __nv_bfloat16*)residual, rp, (__nv_bfloat16*)y, yp, P, C);
  RY_CHECK_LAUNCH();
  return RYOLO_OK;
}

// scale_shift_act with the BatchNorm finalize of a deferred-finalize conv folded in (see scale_shift_act_bn_kernel).
// bn: HOST pointer; bn->partial = the accumulators the conv filled (NOT cleared here: the caller zeroes its arena once per
// forward pass), bn->counter ignored; count = N*Ho*Wo of the conv output.
int ryolo_scale_shift_act_bn(const void* x, long long xp, const ryolo_bn_fuse* bn, double count, int act,
                             const void* residual, long long rp, void* y, long long yp, long long P, int C,
                             void* stream) {
  RY_CHECK_ARG(C % 8 == 0 && C >= 8 && C <= 2048 && xp % 8 == 0 && yp % 8 == 0,
               "scale_shift_act_bn: channels must be a multiple of 8 in [8, 2048]");
  RY_CHECK_ARG(bn && bn->partial && bn->gamma && bn->beta && bn->scale && bn->shift && count > 0,
               "scale_shift_act_bn: incomplete ryolo_bn_fuse");
  RY_CHECK_ARG(P < (1ll << 31) - (1 << 20), "scale_shift_act_bn: more than 2^31 pixels");
  if (P == 0) return RYOLO_OK;
  SsaBn b;
  b.acc = reinterpret_cast<const unsigned long long*>(bn->partial);
  b.gamma = bn->gamma; b.beta = bn->beta;
  b.running_mean = bn->running_mean; b.running_var = bn->running_var; b.num_batches = bn->num_batches;
  b.eps = bn->eps; b.momentum = bn->momentum;
  b.k_sum = 1.0 / ((double)RY_BN_SUM_SCALE * count);
  b.k_sq = 1.0 / ((double)RY_BN_SQ_SCALE * count);
  b.k_unbias = count > 1.0 ? count / (count - 1.0) : 1.0;
  b.scale = bn->scale; b.shift = bn->shift; b.save_mean = bn->save_mean; b.save_invstd = bn->save_invstd;
  const int groups = C / 8;
  RY_CHECK_ARG(groups <= 256, "scale_shift_act_bn: too many channels");
  const int rows = 256 / groups;
  const int per_trip = residual ? 2 : 4;
  long long want = (P + (long long)rows * per_trip - 1) / ((long long)rows * per_trip);
  const long long cap = 2ll * ry_sm_count();
  const int blocks = (int)(want > cap ? cap : (want < 1 ? 1 : want));
  cudaStream_t st = (cudaStream_t)stream;
  const __nv_bfloat16* xx = (const __nv_bfloat16*)x;
  const __nv_bfloat16* rr = (const __nv_bfloat16*)residual;
  __nv_bfloat16* yy = (__nv_bfloat16*)y;
  const size_t fin_smem = (size_t)2 * C * sizeof(float);
#define RY_SSABN(ACT)                                                                                            \
  if (residual) ry_launch(scale_shift_act_bn_kernel<ACT, true>, dim3(blocks), dim3(256), fin_smem, st, xx, xp, b, rr, rp, yy, yp, P, C); \
  else ry_launch(scale_shift_act_bn_kernel<ACT, false>, dim3(blocks), dim3(256), fin_smem, st, xx, xp, b, rr, rp, yy, yp, P, C);
  switch (act) {
    case RYOLO_ACT_LEAKY: RY_SSABN(RYOLO_ACT_LEAKY) break;
    case RYOLO_ACT_MISH: RY_SSABN(RYOLO_ACT_MISH) break;
    case RYOLO_ACT_SWISH: RY_SSABN(RYOLO_ACT_SWISH) break;
    default: RY_SSABN(RYOLO_ACT_LINEAR) break;
  }
#undef RY_SSABN
  RY_CHECK_LAUNCH();
  return RYOLO_OK;
}

int ryolo_maxpool(const void* x, long long xp, int N, int H, int W, int C, int k, int stride, int pad, void* y,
                  long long yp, void* stream) {
  RY_CHECK_ARG(C % 8 == 0 && k >= 1 && stride >= 1, "maxpool: bad arguments");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const long long total = (long long)N * Ho * Wo * (C / 8);
  if (total == 0) return RYOLO_OK;
  if (stride == 1 && (k & 1) && pad == k / 2 && H * W <= 1024 && (long long)N * (C / 8) < (1ll << 31)) {
    maxpool_same_small_kernel<<<(unsigned)(N * (C / 8)), 256, (size_t)2 * H * W * sizeof(uint4), (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, xp, H, W, C, k, (__nv_bfloat16*)y, yp);
    RY_CHECK_LAUNCH();
    return RYOLO_OK;
  }
  maxpool_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, xp, N, H, W, C, k,
                                                                         stride, pad, (__nv_bfloat16*)y, yp, Ho, Wo);
  RY_CHECK_LAUNCH();
  return RYOLO_OK;
}

int ryolo_resize_copy(const void* x, long long xp, int N, int H, int W, int C, int factor, void* y, long long yp,
                      void* stream) {
  RY_CHECK_ARG(C % 8 == 0 && factor >= 1, "resize_copy: bad arguments");
  const long long total = (long long)N * H * factor * W * factor * (C / 8);
  if (total == 0) return RYOLO_OK;
  resize_copy_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, xp, N, H, W, C,
                                                                             factor, (__nv_bfloat16*)y, yp);
  RY_CHECK_LAUNCH();
  return RYOLO_OK;
}

int ryolo_stem_im2col(const float* img, int N, int H, int W, int k, int stride, int Kpad, void* y, void* stream) {
  RY_CHECK_ARG((k == 3 || k == 6) && stride >= 1 && Kpad % 8 == 0 && Kpad >= 3 * k * k, "stem_im2col: bad arguments");
  const int pad = (k - 1) / 2;
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const long long total = (long long)N * Ho * Wo * (Kpad / 8);
  if (total == 0) return RYOLO_OK;
  if (k == 3 && stride == 1 && Kpad == 32)
    stem_im2col_k3_kernel<<<grid_for((long long)N * H * W, 256), 256, 0, (cudaStream_t)stream>>>(img, N, H, W,
                                                                                                 (__nv_bfloat16*)y);
  else
    stem_im2col_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(img, N, H, W, k, stride, pad, Kpad, Ho,
                                                                              Wo, (__nv_bfloat16*)y);
  RY_CHECK_LAUNCH();
  return RYOLO_OK;
}

int ryolo_pack_weights_multi(const ryolo_pack_entry* table_dev, int n, long long total, void* stream) {
  RY_CHECK_ARG(n > 0 && total > 0 && (total & 7) == 0, "pack_weights_multi: empty table or element count not a multiple of 8");
  pack_weights_multi_kernel<<<grid_for(total >> 3, 256), 256, 0, (cudaStream_t)stream>>>(table_dev, n, total);
  RY_CHECK_LAUNCH();
  return RYOLO_OK;
}

int ryolo_unpack_wgrad_multi(const ryolo_pack_entry* table_dev, int n, long long total, void* stream) {
  RY_CHECK_ARG(n > 0 && total > 0, "unpack_wgrad_multi: empty table");
  if (ryolo_knob(RYOLO_KNOB_SSA) != 0)            // the tiled fold shares the "rebuilt HBM passes" switch
    unpack_wgrad_tiled_kernel<<<dim3(48, (unsigned)n), 256, 0, (cudaStream_t)stream>>>(table_dev, n);
  else
    unpack_wgrad_multi_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(table_dev, n, total);
  RY_CHECK_LAUNCH();
  return RYOLO_OK;
}

int ryolo_pack_weights(const float* w, int Cout, int Cin, int k, int stem, void* out, void* stream) {
  RY_CHECK_ARG(Cout > 0 && Cin > 0 && k > 0, "pack_weights: bad shape");
  RY_CHECK_ARG(stem == 0 || stem == 2 || (stem >= 8 && Cin == 3 && stem >= 3 * k * k),
               "pack_weights: the stem layout is for 3-channel convs with Kpad >= 3*k*k");
  const long long total = (long long)Cout * (stem >= 8 ? stem : k * k * Cin);
  pack_weights_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(w, Cout, Cin, k, stem,
                                                                             (__nv_bfloat16*)out);
  RY_CHECK_LAUNCH();
  return RYOLO_OK;
}

}  // extern "C"
