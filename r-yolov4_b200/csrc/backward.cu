// HBM-bound backward companions of the conv stack on NHWC bf16 views (the BatchNorm / activation / pooling /
// upsample / concat halves of autograd's backward at train.py:198):
//   bn_act_bwd_reduce / bn_act_bwd_apply   d(act(BN_train(raw))) -> d raw, d gamma, d beta      model/utils.py:16-23
//   add_into                               gradient accumulation across consumers (residuals, fan-out, concat)
//   maxpool_bwd                            nn.MaxPool2d backward (first maximum in window scan order wins)
//   upsample2x_bwd                         nearest x2 upsample backward (sum over the 2x2 block)
//   head_grad_pack                         fp32 [B,na,gs,gs,ch] head gradient -> bf16 NHWC [B,gs,gs,Cpad] + bias grad
#include "common.cuh"
#include "ryolo_b200.h"
#include <cuda_bf16.h>

namespace {

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

// derivative of the activation at pre-activation value y
template <int ACT>
__device__ __forceinline__ float act_grad(float y) {
  if (ACT == RYOLO_ACT_LEAKY) return y > 0.f ? 1.f : 0.1f;
  if (ACT == RYOLO_ACT_MISH) {           // f = y*t, t = tanh(softplus(y)) = n/(n+2), n = e^y(e^y+2)
    if (y > 20.f) return 1.f;
    const float e = __expf(y);
    const float n = e * (e + 2.f);
    const float d1 = n + 2.f, d2 = 1.f + e;
    const float r = __fdividef(1.f, d1 * d2);          // one reciprocal serves tanh(softplus) and sigmoid
    const float t = n * d2 * r;
    const float sg = e * d1 * r;
    return t + y * (1.f - t * t) * sg;
  }
  if (ACT == RYOLO_ACT_SWISH) {
    const float sg = __fdividef(1.f, 1.f + __expf(-y));
    return sg * (1.f + y * (1.f - sg));
  }
  return 1.f;
}

// Pass 1: per-channel  s1 = sum dY,  s2 = sum dY * xhat   with dY = dOut * act'(raw*scale+shift),
// xhat = (raw - mean) * invstd.  sums = [s1 | s2] (fp32[2C], zeroed by the caller).  dY overwrites dOut (the
// gradient of this layer's output is dead after this pass), so pass 2 needs no transcendental.
template <int ACT>
__global__ void __launch_bounds__(256)
bn_act_bwd_reduce_kernel(__nv_bfloat16* __restrict__ dout, long long dp, const __nv_bfloat16* __restrict__ raw,
                         long long rp, const float* __restrict__ scale, const float* __restrict__ shift,
                         const float* __restrict__ mean, const float* __restrict__ invstd, long long P, int C,
                         float* __restrict__ sums) {
  ry_pdl_wait();
  extern __shared__ float red[];   // [rows][C] x 2
  const int groups = C >> 3;
  const int rows = blockDim.x / groups;
  const int g = threadIdx.x % groups, r = threadIdx.x / groups;
  const int c = 8 * g;
  float s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (r < rows) {
    float sc[8], sh[8], mu[8], is[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { sc[j] = scale[c + j]; sh[j] = shift[c + j]; mu[j] = mean[c + j]; is[j] = invstd[c + j]; }
    const long long stride = (long long)gridDim.x * rows;
    for (long long pix0 = (long long)blockIdx.x * rows + r; pix0 < P; pix0 += 2 * stride) {
      uint4 vd[2], vr[2];
#pragma unroll
      for (int u = 0; u < 2; u++) {
        const long long pix = pix0 + u * stride;
        if (pix < P) {
          vd[u] = *reinterpret_cast<const uint4*>(dout + pix * dp + c);
          vr[u] = *reinterpret_cast<const uint4*>(raw + pix * rp + c);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; u++) {
        const long long pix = pix0 + u * stride;
        if (pix < P) {
          float d[8], x[8];
          unpack8(vd[u], d);
          unpack8(vr[u], x);
#pragma unroll
          for (int j = 0; j < 8; j++) d[j] *= act_grad<ACT>(x[j] * sc[j] + sh[j]);
          const uint4 pk = pack8(d);
          if (ACT != RYOLO_ACT_LINEAR) *reinterpret_cast<uint4*>(dout + pix * dp + c) = pk;
          unpack8(pk, d);                         // statistics of dY as stored
#pragma unroll
          for (int j = 0; j < 8; j++) {
            s1[j] += d[j];
            s2[j] += d[j] * (x[j] - mu[j]) * is[j];
          }
        }
      }
    }
  }
  float* r1 = red;
  float* r2 = red + (size_t)rows * C;
  if (r < rows) {
#pragma unroll
    for (int j = 0; j < 8; j++) { r1[r * C + c + j] = s1[j]; r2[r * C + c + j] = s2[j]; }
  }
  __syncthreads();
  for (int cc = threadIdx.x; cc < C; cc += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int rr = 0; rr < rows; rr++) { a += r1[rr * C + cc]; b += r2[rr * C + cc]; }
    atomicAdd(sums + cc, a);
    atomicAdd(sums + C + cc, b);
  }
}

// Same pass, leaner: mean / invstd are not live in the loop (it accumulates sum dY and sum dY*raw; each thread turns its
// own short partial into sum dY*xhat = invstd * (sum dY*raw - mean * sum dY) afterwards, so no long-range cancellation),
// and mish' uses  t + 4 y e (e+1) r^2  with r = 1/(e^2+2e+2), t = 1-2r  (one exp, one reciprocal, ~12 flops).  Three
// blocks per SM instead of two: the pass is latency-bound (ncu: 24 % warps active, stalls = wait + long scoreboard).
template <int ACT>
__device__ __forceinline__ float act_grad2(float y) {
  if (ACT == RYOLO_ACT_MISH) {
    const float e = __expf(fminf(y, 20.f));
    const float r = ry_rcp_fma(fmaf(e, e + 2.f, 2.f));       // denominator in [2, 2.4e17]
    const float t = fmaf(-2.f, r, 1.f);
    return fmaf(4.f * y, e * (e + 1.f) * (r * r), t);
  }
  return act_grad<ACT>(y);
}

template <int ACT>
__global__ void __launch_bounds__(256, 3)
bn_act_bwd_reduce2_kernel(__nv_bfloat16* __restrict__ dout, long long dp, const __nv_bfloat16* __restrict__ raw,
                          long long rp, const float* __restrict__ scale, const float* __restrict__ shift,
                          const float* __restrict__ mean, const float* __restrict__ invstd, long long P, int C,
                          float* __restrict__ sums) {
  ry_pdl_wait();
  extern __shared__ float red[];   // [rows][C] x 2
  const int groups = C >> 3;
  const int rows = blockDim.x / groups;
  const int g = threadIdx.x % groups, r = threadIdx.x / groups;
  const int c = 8 * g;
  float s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (r < rows) {
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { sc[j] = scale[c + j]; sh[j] = shift[c + j]; }
    const long long stride = (long long)gridDim.x * rows;
    for (long long pix0 = (long long)blockIdx.x * rows + r; pix0 < P; pix0 += 2 * stride) {
      uint4 vd[2], vr[2];
#pragma unroll
      for (int u = 0; u < 2; u++) {
        const long long pix = pix0 + u * stride;
        if (pix < P) {
          vd[u] = *reinterpret_cast<const uint4*>(dout + pix * dp + c);
          vr[u] = *reinterpret_cast<const uint4*>(raw + pix * rp + c);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; u++) {
        const long long pix = pix0 + u * stride;
        if (pix < P) {
          float d[8], x[8];
          unpack8(vd[u], d);
          unpack8(vr[u], x);
          if (ACT != RYOLO_ACT_LINEAR) {
#pragma unroll
            for (int j = 0; j < 8; j++) d[j] *= act_grad2<ACT>(fmaf(x[j], sc[j], sh[j]));
            const uint4 pk = pack8(d);
            *reinterpret_cast<uint4*>(dout + pix * dp + c) = pk;
            unpack8(pk, d);                       // statistics of dY as stored
          }
#pragma unroll
          for (int j = 0; j < 8; j++) {
            s1[j] += d[j];
            s2[j] = fmaf(d[j], x[j], s2[j]);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; j++) s2[j] = invstd[c + j] * (s2[j] - mean[c + j] * s1[j]);
  }
  float* r1 = red;
  float* r2 = red + (size_t)rows * C;
  if (r < rows) {
#pragma unroll
    for (int j = 0; j < 8; j++) { r1[r * C + c + j] = s1[j]; r2[r * C + c + j] = s2[j]; }
  }
  __syncthreads();
  for (int cc = threadIdx.x; cc < C; cc += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int rr = 0; rr < rows; rr++) { a += r1[rr * C + cc]; b += r2[rr * C + cc]; }
    atomicAdd(sums + cc, a);
    atomicAdd(sums + C + cc, b);
  }
}

// Pass 2: d raw = scale * (dY - s1/P - xhat * s2/P) with dY read back from pass 1;  block 0 also emits
// d gamma += s2, d beta += s1.
__global__ void __launch_bounds__(256)
bn_act_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dout, long long dp, const __nv_bfloat16* __restrict__ raw,
                        long long rp, const float* __restrict__ scale, const float* __restrict__ mean,
                        const float* __restrict__ invstd, const float* __restrict__ sums, long long P, int C,
                        __nv_bfloat16* __restrict__ draw, long long op, float* __restrict__ dgamma,
                        float* __restrict__ dbeta) {
  ry_pdl_wait();
  const int groups = C >> 3;
  const int rows = blockDim.x / groups;
  const int g = threadIdx.x % groups, r = threadIdx.x / groups;
  if (blockIdx.x == 0) {
    for (int cc = threadIdx.x; cc < C; cc += blockDim.x) {
      if (dbeta) dbeta[cc] += sums[cc];          // accumulate like autograd's .grad (train.py:198-202)
      if (dgamma) dgamma[cc] += sums[C + cc];
    }
  }
  if (r >= rows) return;
  const int c = 8 * g;
  const float invP = 1.f / (float)P;
  float sc[8], mu[8], is[8], m1[8], m2[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    sc[j] = scale[c + j]; mu[j] = mean[c + j]; is[j] = invstd[c + j];
    m1[j] = sums[c + j] * invP; m2[j] = sums[C + c + j] * invP;
  }
  const long long stride = (long long)gridDim.x * rows;
  for (long long pix0 = (long long)blockIdx.x * rows + r; pix0 < P; pix0 += 2 * stride) {
    uint4 vd[2], vr[2];
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const long long pix = pix0 + u * stride;
      if (pix < P) {
        vd[u] = *reinterpret_cast<const uint4*>(dout + pix * dp + c);
        vr[u] = *reinterpret_cast<const uint4*>(raw + pix * rp + c);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const long long pix = pix0 + u * stride;
      if (pix < P) {
        float d[8], x[8], o[8];
        unpack8(vd[u], d);
        unpack8(vr[u], x);
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const float xh = (x[j] - mu[j]) * is[j];
          o[j] = sc[j] * (d[j] - m1[j] - xh * m2[j]);
        }
        *reinterpret_cast<uint4*>(draw + pix * op + c) = pack8(o);
      }
    }
  }
}

// Variant 2 of the two passes (knob bn_bwd = 2): the reduce pass does not store dY (4 B/element of traffic instead of 6);
// the apply pass recomputes dY = dOut * act'(.) from dOut and raw (6 B/element as before, plus ~15 flops: the
// activation derivative costs one exp and an FMA-pipe reciprocal).  dY never rounds to bf16 and dOut stays intact.
template <int ACT>
__global__ void __launch_bounds__(256, 3)
bn_act_bwd_reduce3_kernel(const __nv_bfloat16* __restrict__ dout, long long dp, const __nv_bfloat16* __restrict__ raw,
                          long long rp, const float* __restrict__ scale, const float* __restrict__ shift,
                          const float* __restrict__ mean, const float* __restrict__ invstd, long long P, int C,
                          float* __restrict__ sums) {
  ry_pdl_wait();
  extern __shared__ float red[];   // [rows][C] x 2
  const int groups = C >> 3;
  const int rows = blockDim.x / groups;
  const int g = threadIdx.x % groups, r = threadIdx.x / groups;
  const int c = 8 * g;
  float s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (r < rows) {
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { sc[j] = scale[c + j]; sh[j] = shift[c + j]; }
    const long long stride = (long long)gridDim.x * rows;
    for (long long pix0 = (long long)blockIdx.x * rows + r; pix0 < P; pix0 += 2 * stride) {
      uint4 vd[2], vr[2];
#pragma unroll
      for (int u = 0; u < 2; u++) {
        const long long pix = pix0 + u * stride;
        if (pix < P) {
          vd[u] = *reinterpret_cast<const uint4*>(dout + pix * dp + c);
          vr[u] = *reinterpret_cast<const uint4*>(raw + pix * rp + c);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; u++) {
        const long long pix = pix0 + u * stride;
        if (pix < P) {
          float d[8], x[8];
          unpack8(vd[u], d);
          unpack8(vr[u], x);
#pragma unroll
          for (int j = 0; j < 8; j++) {
            if (ACT != RYOLO_ACT_LINEAR) d[j] *= act_grad2<ACT>(fmaf(x[j], sc[j], sh[j]));
            s1[j] += d[j];
            s2[j] = fmaf(d[j], x[j], s2[j]);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; j++) s2[j] = invstd[c + j] * (s2[j] - mean[c + j] * s1[j]);
  }
  float* r1 = red;
  float* r2 = red + (size_t)rows * C;
  if (r < rows) {
#pragma unroll
    for (int j = 0; j < 8; j++) { r1[r * C + c + j] = s1[j]; r2[r * C + c + j] = s2[j]; }
  }
  __syncthreads();
  for (int cc = threadIdx.x; cc < C; cc += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int rr = 0; rr < rows; rr++) { a += r1[rr * C + cc]; b += r2[rr * C + cc]; }
    atomicAdd(sums + cc, a);
    atomicAdd(sums + C + cc, b);
  }
}

template <int ACT>
__global__ void __launch_bounds__(256, 3)
bn_act_bwd_apply2_kernel(const __nv_bfloat16* __restrict__ dout, long long dp, const __nv_bfloat16* __restrict__ raw,
                         long long rp, const float* __restrict__ scale, const float* __restrict__ shift,
                         const float* __restrict__ mean, const float* __restrict__ invstd,
                         const float* __restrict__ sums, long long P, int C, __nv_bfloat16* __restrict__ draw,
                         long long op, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  ry_pdl_wait();
  const int groups = C >> 3;
  const int rows = blockDim.x / groups;
  const int g = threadIdx.x % groups, r = threadIdx.x / groups;
  if (blockIdx.x == 0) {
    for (int cc = threadIdx.x; cc < C; cc += blockDim.x) {
      if (dbeta) dbeta[cc] += sums[cc];
      if (dgamma) dgamma[cc] += sums[C + cc];
    }
  }
  if (r >= rows) return;
  const int c = 8 * g;
  const float invP = 1.f / (float)P;
  // o = sc*(d - m1 - (x - mu)*is*m2) = sc*d - x*k2 - k1   with k2 = sc*is*m2, k1 = sc*(m1 - mu*is*m2)
  float sc[8], sh[8], k1[8], k2[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    sc[j] = scale[c + j]; sh[j] = shift[c + j];
    const float is = invstd[c + j], mu = mean[c + j];
    const float m1 = sums[c + j] * invP, m2 = sums[C + c + j] * invP;
    k2[j] = sc[j] * is * m2;
    k1[j] = sc[j] * (m1 - mu * is * m2);
  }
  const long long stride = (long long)gridDim.x * rows;
  for (long long pix0 = (long long)blockIdx.x * rows + r; pix0 < P; pix0 += 2 * stride) {
    uint4 vd[2], vr[2];
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const long long pix = pix0 + u * stride;
      if (pix < P) {
        vd[u] = *reinterpret_cast<const uint4*>(dout + pix * dp + c);
        vr[u] = *reinterpret_cast<const uint4*>(raw + pix * rp + c);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const long long pix = pix0 + u * stride;
      if (pix < P) {
        float d[8], x[8], o[8];
        unpack8(vd[u], d);
        unpack8(vr[u], x);
#pragma unroll
        for (int j = 0; j < 8; j++) {
          if (ACT != RYOLO_ACT_LINEAR) d[j] *= act_grad2<ACT>(fmaf(x[j], sc[j], sh[j]));
          o[j] = fmaf(sc[j], d[j], -fmaf(x[j], k2[j], k1[j]));
        }
        *reinterpret_cast<uint4*>(draw + pix * op + c) = pack8(o);
      }
    }
  }
}

// Variant 3 of the two passes (knob bn_bwd = 3, default): same traffic as variant 2 (4 + 6 B/element), rebuilt around the
// two things ncu showed the variant-2 kernels waiting on (profiles/r02_elementwise_notes.txt):
//   * memory latency: a thread's next two rows are requested BEFORE the current two are processed (register double
//     buffer), so every resident warp keeps 64-128 bytes in flight through its ~300-instruction compute phase; the grid
//     is exactly the resident capacity (2 blocks/SM x SMs, grid-stride) instead of 2.7 waves of short blocks;
//   * issue slots: Mish' costs 13 instructions (ex2 and rcp on the SFU pipe, which has the headroom here: 2 x 8 pipe
//     cycles against ~20 issue slots per warp row) instead of ~24 with the FMA-pipe reciprocal; out-of-range rows are
//     loaded as zeros so the compute phase carries no predicates.
__device__ __forceinline__ float ry_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ry_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// d act / d z at pre-activation z.  Mish: f = z*t, t = tanh(softplus(z)) = 1 - 2r, r = 1/((e+1)^2 + 1), e = e^z;
// f' = t + 4 z e (e+1) r^2.  z > 20 is clamped inside the exponential only (r -> 4e-18, f' -> 1).
template <int ACT>
__device__ __forceinline__ float act_grad3(float z) {
  if (ACT == RYOLO_ACT_MISH) {
    const float e = ry_ex2(fminf(z * 1.4426950408889634f, 28.853900817779268f));
    const float p = e + 1.f;
    const float r = ry_rcp(fmaf(p, p, 1.f));
    const float w = (e * p) * r;
    return fmaf((z * w) * r, 4.f, fmaf(-2.f, r, 1.f));
  }
  if (ACT == RYOLO_ACT_SWISH) {           // sg (1 + z (1 - sg)), sg = 1/(1 + e^-z)
    const float sg = ry_rcp(1.f + ry_ex2(fminf(z * -1.4426950408889634f, 86.f)));
    return sg * fmaf(z, 1.f - sg, 1.f);
  }
  if (ACT == RYOLO_ACT_LEAKY) return z > 0.f ? 1.f : 0.1f;
  return 1.f;
}

// Mish' with the pre-activation already in log2 units (zl = z * log2 e, the BatchNorm scale/shift carry the factor):
// one instruction fewer than act_grad3 in the reduce pass, which is the issue-bound one (4 B/element).
__device__ __forceinline__ float mish_grad_l2(float zl) {
  const float e = ry_ex2(fminf(zl, 28.853900817779268f));
  const float p = e + 1.f;
  const float r = ry_rcp(fmaf(p, p, 1.f));
  const float w = (e * p) * r;
  return fmaf((zl * w) * r, 4.f * 0.6931471805599453f, fmaf(-2.f, r, 1.f));
}

// Rows of this thread: pix0, pix0 + stride, ... < P, walked as pairs; `n` = how many.
__device__ __forceinline__ int rows_of(long long P, int pix0, int stride) {
  return pix0 < P ? (int)((P - 1 - pix0) / stride) + 1 : 0;
}
__device__ __forceinline__ uint4 ldv(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }

template <int ACT>
__device__ __forceinline__ void reduce_row(const uint4& vd, const uint4& vr, const float (&sc)[8], const float (&sh)[8],
                                           float (&s1)[8], float (&s2)[8]) {
  float d[8], x[8];
  unpack8u(vd, d);
  unpack8u(vr, x);
#pragma unroll
  for (int j = 0; j < 8; j++) {
    if (ACT == RYOLO_ACT_MISH) d[j] *= mish_grad_l2(fmaf(x[j], sc[j], sh[j]));      // sc, sh pre-multiplied by log2 e
    else if (ACT != RYOLO_ACT_LINEAR) d[j] *= act_grad3<ACT>(fmaf(x[j], sc[j], sh[j]));
    s1[j] += d[j];
    s2[j] = fmaf(d[j], x[j], s2[j]);
  }
}

// Pass 1 for one block: per-channel sums of dY and dY*xhat over the rows this block walks, added to sums[2C].
// The block-level combine needs 8*C bytes of shared memory (C <= 768) or none (direct global atomics), so that it fits
// beside a resident wgrad CTA's boxes.
template <int ACT>
__device__ __forceinline__ void bn_bwd_reduce_phase(const __nv_bfloat16* __restrict__ dout, long long dp,
                                                    const __nv_bfloat16* __restrict__ raw, long long rp,
                                                    const float* __restrict__ scale, const float* __restrict__ shift,
                                                    const float* __restrict__ mean, const float* __restrict__ invstd,
                                                    long long P, int C, float* __restrict__ sums, float* red) {
  const int groups = C >> 3;
  const int rows = blockDim.x / groups;
  const int g = threadIdx.x % groups, r = threadIdx.x / groups;
  const int c = 8 * g;
  float s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (r < rows) {
    float sc[8], sh[8];
    const float k = ACT == RYOLO_ACT_MISH ? 1.4426950408889634f : 1.f;
#pragma unroll
    for (int j = 0; j < 8; j++) { sc[j] = scale[c + j] * k; sh[j] = shift[c + j] * k; }
    const int stride = (int)gridDim.x * rows;
    const int pix0 = (int)blockIdx.x * rows + r;
    const int n = rows_of(P, pix0, stride);
    const long long sd = (long long)stride * dp, sr = (long long)stride * rp;      // one row step, in elements
    const __nv_bfloat16* pd = dout + c + (long long)pix0 * dp;
    const __nv_bfloat16* pr = raw + c + (long long)pix0 * rp;
    const int pairs = n >> 1;
    if (pairs) {
      uint4 d0 = ldv(pd), d1 = ldv(pd + sd), r0 = ldv(pr), r1 = ldv(pr + sr);
#pragma unroll 2
      for (int i = 1; i < pairs; i++) {
        pd += 2 * sd;
        pr += 2 * sr;
        const uint4 e0 = ldv(pd), e1 = ldv(pd + sd), q0 = ldv(pr), q1 = ldv(pr + sr);    // next pair: in flight while
        reduce_row<ACT>(d0, r0, sc, sh, s1, s2);                                          // this one is processed
        reduce_row<ACT>(d1, r1, sc, sh, s1, s2);
        d0 = e0; d1 = e1; r0 = q0; r1 = q1;
      }
      reduce_row<ACT>(d0, r0, sc, sh, s1, s2);
      reduce_row<ACT>(d1, r1, sc, sh, s1, s2);
      pd += 2 * sd;
      pr += 2 * sr;
    }
    if (n & 1) reduce_row<ACT>(ldv(pd), ldv(pr), sc, sh, s1, s2);
#pragma unroll
    for (int j = 0; j < 8; j++) s2[j] = invstd[c + j] * (s2[j] - mean[c + j] * s1[j]);
  }
  // block-level combine: lanes holding the same channel group first add up by shuffle (groups a power of two <= 32:
  // every thread of the warp is active), then one lane per (warp, group) adds into the block's 2*C floats
  const bool use_smem = C <= 768;
  if (use_smem) {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
  }
  int leader = r < rows;
  if (groups <= 32 && (groups & (groups - 1)) == 0) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      for (int o = groups; o < 32; o <<= 1) {
        s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
        s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o);
      }
    }
    leader = (threadIdx.x & 31) < groups;
  }
  if (leader) {
    float* dst = use_smem ? red : sums;
#pragma unroll
    for (int j = 0; j < 8; j++) { atomicAdd(dst + c + j, s1[j]); atomicAdd(dst + C + c + j, s2[j]); }
  }
  if (use_smem) {
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(sums + i, red[i]);
  }
}

// MAXT only sets the register budget (the block is always 256 threads): 256 -> 128 registers, 304 -> 104, which lets
// two blocks share an SM with one 192-thread x 56-register wgrad CTA.
template <int ACT, int MAXT>
__global__ void __launch_bounds__(MAXT, 2)
bn_act_bwd_reduce4_kernel(const __nv_bfloat16* __restrict__ dout, long long dp, const __nv_bfloat16* __restrict__ raw,
                          long long rp, const float* __restrict__ scale, const float* __restrict__ shift,
                          const float* __restrict__ mean, const float* __restrict__ invstd, long long P, int C,
                          float* __restrict__ sums) {
  ry_pdl_wait();
  extern __shared__ float red[];   // [2][C] when C <= 768
  bn_bwd_reduce_phase<ACT>(dout, dp, raw, rp, scale, shift, mean, invstd, P, C, sums, red);
}

template <int ACT>
__device__ __forceinline__ uint4 apply_row(const uint4& vd, const uint4& vr, const float (&sc)[8], const float (&sh)[8],
                                           const float (&nk1)[8], const float (&nk2)[8]) {
  float d[8], x[8], o[8];
  unpack8u(vd, d);
  unpack8u(vr, x);
#pragma unroll
  for (int j = 0; j < 8; j++) {
    if (ACT != RYOLO_ACT_LINEAR) d[j] *= act_grad3<ACT>(fmaf(x[j], sc[j], sh[j]));
    o[j] = fmaf(sc[j], d[j], fmaf(x[j], nk2[j], nk1[j]));
  }
  return pack8(o);
}

// Pass 2 for one block: d raw = scale * (dY - s1/P - xhat * s2/P), dY recomputed from d out and raw; block 0 also emits
// d gamma += s2, d beta += s1.  `sums` is read through L2 (it may have been written by this very launch, see the fused
// kernel below).
template <int ACT>
__device__ __forceinline__ void bn_bwd_apply_phase(const __nv_bfloat16* __restrict__ dout, long long dp,
                                                   const __nv_bfloat16* __restrict__ raw, long long rp,
                                                   const float* __restrict__ scale, const float* __restrict__ shift,
                                                   const float* __restrict__ mean, const float* __restrict__ invstd,
                                                   const float* sums, long long P, int C,
                                                   __nv_bfloat16* __restrict__ draw, long long op,
                                                   float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                   __nv_bfloat16* __restrict__ dres = nullptr, long long rsp = 0) {
  const int groups = C >> 3;
  const int rows = blockDim.x / groups;
  const int g = threadIdx.x % groups, r = threadIdx.x / groups;
  if (blockIdx.x == 0) {
    for (int cc = threadIdx.x; cc < C; cc += blockDim.x) {
      if (dbeta) dbeta[cc] += __ldcg(sums + cc);
      if (dgamma) dgamma[cc] += __ldcg(sums + C + cc);
    }
  }
  if (r >= rows) return;
  const int c = 8 * g;
  const float invP = 1.f / (float)P;
  // o = sc*(d - m1 - (x - mu)*is*m2) = sc*d + (x*nk2 + nk1)   with nk2 = -sc*is*m2, nk1 = -sc*(m1 - mu*is*m2)
  float sc[8], sh[8], nk1[8], nk2[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    sc[j] = scale[c + j]; sh[j] = shift[c + j];
    const float is = invstd[c + j], mu = mean[c + j];
    const float m1 = __ldcg(sums + c + j) * invP, m2 = __ldcg(sums + C + c + j) * invP;
    nk2[j] = -(sc[j] * is * m2);
    nk1[j] = -(sc[j] * (m1 - mu * is * m2));
  }
  const int stride = (int)gridDim.x * rows;
  const int pix0 = (int)blockIdx.x * rows + r;
  const int n = rows_of(P, pix0, stride);
  const long long sd = (long long)stride * dp, sr = (long long)stride * rp, so = (long long)stride * op;
  const __nv_bfloat16* pd = dout + c + (long long)pix0 * dp;
  const __nv_bfloat16* pr = raw + c + (long long)pix0 * rp;
  __nv_bfloat16* po = draw + c + (long long)pix0 * op;
  // optional copy of d out into the gradient of a residual operand (out = residual + act(bn(raw)): the identity path
  // gets d out as is); the words are already in registers, so this replaces a separate read + write pass
  const long long sq = (long long)stride * rsp;
  __nv_bfloat16* pq = dres ? dres + c + (long long)pix0 * rsp : nullptr;
  const int pairs = n >> 1;
  if (pairs) {
    uint4 d0 = ldv(pd), d1 = ldv(pd + sd), r0 = ldv(pr), r1 = ldv(pr + sr);
#pragma unroll 2
    for (int i = 1; i < pairs; i++) {
      pd += 2 * sd;
      pr += 2 * sr;
      const uint4 e0 = ldv(pd), e1 = ldv(pd + sd), q0 = ldv(pr), q1 = ldv(pr + sr);
      if (pq) { *reinterpret_cast<uint4*>(pq) = d0; *reinterpret_cast<uint4*>(pq + sq) = d1; pq += 2 * sq; }
      *reinterpret_cast<uint4*>(po) = apply_row<ACT>(d0, r0, sc, sh, nk1, nk2);
      *reinterpret_cast<uint4*>(po + so) = apply_row<ACT>(d1, r1, sc, sh, nk1, nk2);
      po += 2 * so;
      d0 = e0; d1 = e1; r0 = q0; r1 = q1;
    }
    if (pq) { *reinterpret_cast<uint4*>(pq) = d0; *reinterpret_cast<uint4*>(pq + sq) = d1; pq += 2 * sq; }
    *reinterpret_cast<uint4*>(po) = apply_row<ACT>(d0, r0, sc, sh, nk1, nk2);
    *reinterpret_cast<uint4*>(po + so) = apply_row<ACT>(d1, r1, sc, sh, nk1, nk2);
    pd += 2 * sd;
    pr += 2 * sr;
    po += 2 * so;
  }
  if (n & 1) {
    const uint4 d0 = ldv(pd);
    if (pq) *reinterpret_cast<uint4*>(pq) = d0;
    *reinterpret_cast<uint4*>(po) = apply_row<ACT>(d0, ldv(pr), sc, sh, nk1, nk2);
  }
}

template <int ACT, int MAXT>
__global__ void __launch_bounds__(MAXT, 2)
bn_act_bwd_apply3_kernel(const __nv_bfloat16* __restrict__ dout, long long dp, const __nv_bfloat16* __restrict__ raw,
                         long long rp, const float* __restrict__ scale, const float* __restrict__ shift,
                         const float* __restrict__ mean, const float* __restrict__ invstd,
                         const float* __restrict__ sums, long long P, int C, __nv_bfloat16* __restrict__ draw,
                         long long op, float* __restrict__ dgamma, float* __restrict__ dbeta,
                         __nv_bfloat16* __restrict__ dres, long long rsp) {
  ry_pdl_wait();
  bn_bwd_apply_phase<ACT>(dout, dp, raw, rp, scale, shift, mean, invstd, sums, P, C, draw, op, dgamma, dbeta, dres, rsp);
}

// (Experiment, knob bn_fuse, off by default: measured neutral to slightly slower on the step.)
// Both passes in ONE launch for layers whose d out + raw fit in L2 (the 25x25 / 50x50 maps: 56 of the 110 layers of
// yolov4, 2.4 ms per step as launch pairs of 14-27 us that move 20-80 MB each): the grid is at most the resident
// capacity, so the blocks can meet at a grid-wide barrier between the passes; the second pass then re-reads its
// operands from L2 instead of HBM and the second launch (fixed cost, ramp, tail) disappears.
// Barrier: arrival counter + generation word; a block reads the generation BEFORE it arrives, so the last arriver's bump
// cannot be missed.  Blocks that are not resident yet (an SM still held by a side-stream wgrad CTA, which never waits for
// this kernel) arrive late, not never; a bounded spin turns an impossible wait into a trap instead of a hang.
__device__ unsigned int g_bn_bar_count = 0;
__device__ volatile unsigned int g_bn_bar_gen = 0;

__device__ __forceinline__ void bn_grid_barrier() {
  __threadfence();                               // this thread's atomics into sums[] are visible device-wide
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int gen = g_bn_bar_gen;
    __threadfence();
    if (atomicAdd(&g_bn_bar_count, 1u) == gridDim.x - 1) {
      g_bn_bar_count = 0;
      __threadfence();
      g_bn_bar_gen = gen + 1;
    } else {
      long long spins = 0;
      while (g_bn_bar_gen == gen) {
        __nanosleep(100);
        if (++spins > (1ll << 25)) __trap();     // ~3 s: co-residency assumption violated
      }
    }
    __threadfence();
  }
  __syncthreads();
}

template <int ACT>
__global__ void __launch_bounds__(256, 2)
bn_act_bwd_fused_kernel(const __nv_bfloat16* __restrict__ dout, long long dp, const __nv_bfloat16* __restrict__ raw,
                        long long rp, const float* __restrict__ scale, const float* __restrict__ shift,
                        const float* __restrict__ mean, const float* __restrict__ invstd, float* __restrict__ sums,
                        long long P, int C, __nv_bfloat16* __restrict__ draw, long long op,
                        float* __restrict__ dgamma, float* __restrict__ dbeta) {
  ry_pdl_wait();
  extern __shared__ float red[];   // [2][C] when C <= 768
  bn_bwd_reduce_phase<ACT>(dout, dp, raw, rp, scale, shift, mean, invstd, P, C, sums, red);
  bn_grid_barrier();
  bn_bwd_apply_phase<ACT>(dout, dp, raw, rp, scale, shift, mean, invstd, sums, P, C, draw, op, dgamma, dbeta);
}

// ds = dout * act'(x1*s1+b1 + x2*s2+b2)   (two-branch pre-activation, RepConv)
template <int ACT>
__global__ void __launch_bounds__(256)
act_bwd2_kernel(const __nv_bfloat16* __restrict__ dout, long long dp, const __nv_bfloat16* __restrict__ x1,
                long long p1, const float* __restrict__ s1, const float* __restrict__ b1,
                const __nv_bfloat16* __restrict__ x2, long long p2, const float* __restrict__ s2,
                const float* __restrict__ b2, __nv_bfloat16* __restrict__ ds, long long op, long long P, int C) {
  const int groups = C >> 3;
  const long long total = P * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / groups;
    const int c = (int)(i - pix * groups) * 8;
    float d[8], a[8], b[8], o[8];
    unpack8(*reinterpret_cast<const uint4*>(dout + pix * dp + c), d);
    unpack8(*reinterpret_cast<const uint4*>(x1 + pix * p1 + c), a);
    unpack8(*reinterpret_cast<const uint4*>(x2 + pix * p2 + c), b);
#pragma unroll
    for (int j = 0; j < 8; j++)
      o[j] = d[j] * act_grad<ACT>(a[j] * s1[c + j] + b1[c + j] + b[j] * s2[c + j] + b2[c + j]);
    *reinterpret_cast<uint4*>(ds + pix * op + c) = pack8(o);
  }
}

// dst (+)= src over P pixels x C channels
__global__ void __launch_bounds__(256)
add_into_kernel(__nv_bfloat16* __restrict__ dst, long long dpitch, const __nv_bfloat16* __restrict__ src, long long sp,
                long long P, int C, int accumulate) {
  const int groups = C >> 3;
  const long long total = P * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / groups;
    const int c = (int)(i - pix * groups) * 8;
    uint4 v = *reinterpret_cast<const uint4*>(src + pix * sp + c);
    if (accumulate) {
      float a[8], b[8];
      unpack8(v, a);
      unpack8(*reinterpret_cast<const uint4*>(dst + pix * dpitch + c), b);
#pragma unroll
      for (int j = 0; j < 8; j++) a[j] += b[j];
      v = pack8(a);
    }
    *reinterpret_cast<uint4*>(dst + pix * dpitch + c) = v;
  }
}

// MaxPool2d backward: every output routes its gradient to the first maximum of its window (row-major scan,
// strict >).  Windows overlap (stride < k), so contributions are summed with fp32 atomics in a dense scratch
// [N,H,W,C] (a 13x13 window can send 169 gradients to one element: bf16 accumulation would lose them).
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ x, long long xp, const __nv_bfloat16* __restrict__ dy,
                   long long dyp, int N, int H, int W, int C, int k, int stride, int pad, int Ho, int Wo,
                   float* __restrict__ acc) {
  const int groups = C >> 3;
  const long long total = (long long)N * Ho * Wo * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % groups) * 8;
    long long pix = i / groups;
    const int wo = (int)(pix % Wo), ho = (int)((pix / Wo) % Ho), n = (int)(pix / ((long long)Wo * Ho));
    float m[8];
    int am[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { m[j] = -INFINITY; am[j] = -1; }
    for (int dh = 0; dh < k; dh++) {
      const int hi = ho * stride + dh - pad;
      if (hi < 0 || hi >= H) continue;
      for (int dw = 0; dw < k; dw++) {
        const int wi = wo * stride + dw - pad;
        if (wi < 0 || wi >= W) continue;
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(x + (((long long)n * H + hi) * W + wi) * xp + c), f);
#pragma unroll
        for (int j = 0; j < 8; j++)
          if (f[j] > m[j] || am[j] < 0) { m[j] = f[j]; am[j] = hi * W + wi; }
      }
    }
    float g[8];
    unpack8(*reinterpret_cast<const uint4*>(dy + pix * dyp + c), g);
#pragma unroll
    for (int j = 0; j < 8; j++)
      if (am[j] >= 0) atomicAdd(acc + ((long long)n * H * W + am[j]) * C + c + j, g[j]);
  }
}

// Non-overlapping pool (k == stride, no padding, H and W multiples of k: yolov7's MaxConv 2x2/s2): every input element
// belongs to exactly one window, so its gradient is written directly (the window's first maximum gets dy, the others 0):
// no scratch, no atomics, no second pass.
__global__ void __launch_bounds__(256)
maxpool_tiled_bwd_kernel(const __nv_bfloat16* __restrict__ x, long long xp, const __nv_bfloat16* __restrict__ dy,
                         long long dyp, int N, int H, int W, int C, int k, __nv_bfloat16* __restrict__ dx, long long dxp,
                         int accumulate) {
  const int groups = C >> 3, Ho = H / k, Wo = W / k;
  const long long total = (long long)N * Ho * Wo * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % groups) * 8;
    const long long pix = i / groups;
    const int wo = (int)(pix % Wo), ho = (int)((pix / Wo) % Ho), n = (int)(pix / ((long long)Wo * Ho));
    float m[8], g[8];
    int am[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { m[j] = -INFINITY; am[j] = -1; }
    for (int dh = 0; dh < k; dh++) {
      for (int dw = 0; dw < k; dw++) {
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(x + (((long long)n * H + ho * k + dh) * W + wo * k + dw) * xp + c), f);
#pragma unroll
        for (int j = 0; j < 8; j++)
          if (f[j] > m[j] || am[j] < 0) { m[j] = f[j]; am[j] = dh * k + dw; }
      }
    }
    unpack8(*reinterpret_cast<const uint4*>(dy + pix * dyp + c), g);
    for (int dh = 0; dh < k; dh++) {
      for (int dw = 0; dw < k; dw++) {
        __nv_bfloat16* o = dx + (((long long)n * H + ho * k + dh) * W + wo * k + dw) * dxp + c;
        float a[8];
#pragma unroll
        for (int j = 0; j < 8; j++) a[j] = am[j] == dh * k + dw ? g[j] : 0.f;
        if (accumulate) {
          float b[8];
          unpack8(*reinterpret_cast<const uint4*>(o), b);
#pragma unroll
          for (int j = 0; j < 8; j++) a[j] += b[j];
        }
        *reinterpret_cast<uint4*>(o) = pack8(a);
      }
    }
  }
}

// Same routing for a stride-1 "same" pool on a small map (SPP), one block per (image, 8 channels): the map, its row
// maxima WITH the column of their first maximum, and an fp32 gradient tile live in shared memory.  The window maximum is
// separable; PyTorch's winner (first maximum in row-major scan order) is the first row whose row-window maximum equals
// the window maximum, and inside that row the first column attaining it: both "first" rules are a strict > while
// scanning forwards.  A thread owns 4 channels of a pixel (128-bit shared loads, four independent compare chains):
// 2(2p+1) vector loads per output instead of ~4(2p+1) scalar ones with two data-dependent while loops.
// Up to three pools of the SAME map (SPP: 5 / 9 / 13) in one launch: the map is read once, every pool routes its
// gradient into the same fp32 tile, and dx is written (or read-added) once instead of three times.
struct SppGrads {
  const __nv_bfloat16* dy[3];
  int k[3];
  int n;
};

__global__ void __launch_bounds__(256)
maxpool_same_small_bwd_kernel(const __nv_bfloat16* __restrict__ x, long long xp, SppGrads dys, long long dyp, int H, int W,
                              int C, __nv_bfloat16* __restrict__ dx, long long dxp, int accumulate) {
  extern __shared__ float sm_pb[];             // [H*W][8] map | [H*W][8] row maxima | [H*W][8] gradient | [H*W][8] u8 columns
  const int HW = H * W;
  float* sx = sm_pb;
  float* sr = sm_pb + 8 * HW;
  float* sg = sm_pb + 16 * HW;
  unsigned char* sa = reinterpret_cast<unsigned char*>(sm_pb + 24 * HW);
  const int groups = C >> 3;
  const int n = blockIdx.x / groups, c = (blockIdx.x % groups) * 8;
  const __nv_bfloat16* xb = x + (long long)n * HW * xp + c;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(xb + (long long)i * xp), f);
    *reinterpret_cast<float4*>(sx + i * 8) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(sx + i * 8 + 4) = make_float4(f[4], f[5], f[6], f[7]);
    *reinterpret_cast<float4*>(sg + i * 8) = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(sg + i * 8 + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int pi = 0; pi < dys.n; pi++) {
  const int p = dys.k[pi] >> 1;
  __syncthreads();
  // pass 1: per (pixel, 4 channels) the maximum of the row window and the first column that attains it
  for (int i = threadIdx.x; i < HW * 2; i += blockDim.x) {
    const int pix = i >> 1, q = (i & 1) * 4;
    const int h = pix / W, w = pix - h * W;
    const int w0 = max(w - p, 0), w1 = min(w + p, W - 1);
    float4 m = *reinterpret_cast<const float4*>(sx + (h * W + w0) * 8 + q);
    int ax = w0, ay = w0, az = w0, aw = w0;
    for (int ww = w0 + 1; ww <= w1; ww++) {
      const float4 v = *reinterpret_cast<const float4*>(sx + (h * W + ww) * 8 + q);
      if (v.x > m.x) { m.x = v.x; ax = ww; }
      if (v.y > m.y) { m.y = v.y; ay = ww; }
      if (v.z > m.z) { m.z = v.z; az = ww; }
      if (v.w > m.w) { m.w = v.w; aw = ww; }
    }
    *reinterpret_cast<float4*>(sr + pix * 8 + q) = m;
    *reinterpret_cast<uchar4*>(sa + pix * 8 + q) = make_uchar4((unsigned char)ax, (unsigned char)ay, (unsigned char)az,
                                                               (unsigned char)aw);
  }
  __syncthreads();
  // pass 2: scan the rows of the window; the first row with the largest row maximum holds the winner
  const __nv_bfloat16* gb = dys.dy[pi] + (long long)n * HW * dyp + c;
  for (int i = threadIdx.x; i < HW * 2; i += blockDim.x) {
    const int pix = i >> 1, q = (i & 1) * 4;
    const int h = pix / W, w = pix - h * W;
    const int h0 = max(h - p, 0), h1 = min(h + p, H - 1);
    float4 m = *reinterpret_cast<const float4*>(sr + (h0 * W + w) * 8 + q);
    int rx = h0, ry = h0, rz = h0, rw = h0;
    for (int hh = h0 + 1; hh <= h1; hh++) {
      const float4 v = *reinterpret_cast<const float4*>(sr + (hh * W + w) * 8 + q);
      if (v.x > m.x) { m.x = v.x; rx = hh; }
      if (v.y > m.y) { m.y = v.y; ry = hh; }
      if (v.z > m.z) { m.z = v.z; rz = hh; }
      if (v.w > m.w) { m.w = v.w; rw = hh; }
    }
    const uint2 graw = *reinterpret_cast<const uint2*>(gb + (long long)pix * dyp + q);
    const float g0 = __uint_as_float(graw.x << 16), g1 = __uint_as_float(graw.x & 0xffff0000u);
    const float g2 = __uint_as_float(graw.y << 16), g3 = __uint_as_float(graw.y & 0xffff0000u);
    atomicAdd(&sg[(rx * W + sa[(rx * W + w) * 8 + q + 0]) * 8 + q + 0], g0);
    atomicAdd(&sg[(ry * W + sa[(ry * W + w) * 8 + q + 1]) * 8 + q + 1], g1);
    atomicAdd(&sg[(rz * W + sa[(rz * W + w) * 8 + q + 2]) * 8 + q + 2], g2);
    atomicAdd(&sg[(rw * W + sa[(rw * W + w) * 8 + q + 3]) * 8 + q + 3], g3);
  }
  __syncthreads();
  }
  __nv_bfloat16* db = dx + (long long)n * HW * dxp + c;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; j++) a[j] = sg[i * 8 + j];
    if (accumulate) {
      float b[8];
      unpack8(*reinterpret_cast<const uint4*>(db + (long long)i * dxp), b);
#pragma unroll
      for (int j = 0; j < 8; j++) a[j] += b[j];
    }
    *reinterpret_cast<uint4*>(db + (long long)i * dxp) = pack8(a);
  }
}

// dst (bf16 view) (+)= acc (dense fp32 [P, C])
__global__ void __launch_bounds__(256)
add_f32_into_kernel(__nv_bfloat16* __restrict__ dst, long long dpitch, const float* __restrict__ acc, long long P, int C,
                    int accumulate) {
  const int groups = C >> 3;
  const long long total = P * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / groups;
    const int c = (int)(i - pix * groups) * 8;
    const float4 a0 = *reinterpret_cast<const float4*>(acc + pix * C + c);
    const float4 a1 = *reinterpret_cast<const float4*>(acc + pix * C + c + 4);
    float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    if (accumulate) {
      float b[8];
      unpack8(*reinterpret_cast<const uint4*>(dst + pix * dpitch + c), b);
#pragma unroll
      for (int j = 0; j < 8; j++) a[j] += b[j];
    }
    *reinterpret_cast<uint4*>(dst + pix * dpitch + c) = pack8(a);
  }
}

// dx (+)= sum of the 2x2 block of dy  (nearest x2 upsample backward)
__global__ void __launch_bounds__(256)
upsample2x_bwd_kernel(const __nv_bfloat16* __restrict__ dy, long long dyp, int N, int H, int W, int C,
                      __nv_bfloat16* __restrict__ dx, long long dxp, int accumulate) {
  const int groups = C >> 3;
  const long long total = (long long)N * H * W * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % groups) * 8;
    long long pix = i / groups;
    const int w = (int)(pix % W), h = (int)((pix / W) % H), n = (int)(pix / ((long long)W * H));
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int dh = 0; dh < 2; dh++)
#pragma unroll
      for (int dw = 0; dw < 2; dw++) {
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(dy + (((long long)n * 2 * H + 2 * h + dh) * 2 * W + 2 * w + dw) * dyp + c), f);
#pragma unroll
        for (int j = 0; j < 8; j++) a[j] += f[j];
      }
    if (accumulate) {
      float b[8];
      unpack8(*reinterpret_cast<const uint4*>(dx + pix * dxp + c), b);
#pragma unroll
      for (int j = 0; j < 8; j++) a[j] += b[j];
    }
    *reinterpret_cast<uint4*>(dx + pix * dxp + c) = pack8(a);
  }
}

// glev fp32 [B, na, H, W, ch] -> out bf16 [B, H, W, Cpad] (channel c = a*ch + k, zero for c >= na*ch), multiplied by
// mul[c] when given (ImplicitM); dbias[c] += sum over pixels of the packed value.  A thread owns 8 consecutive output
// channels (one 16-byte store per pixel) and walks pixels with a block-wide stride, like the other HBM-bound kernels;
// its 8 source offsets and bias partials stay in registers.
__global__ void __launch_bounds__(256)
head_grad_pack_kernel(const float* __restrict__ glev, int B, int na, int H, int W, int ch, int Cpad,
                      const float* __restrict__ mul, __nv_bfloat16* __restrict__ out, float* __restrict__ dbias,
                      const float* __restrict__ yhead, float* __restrict__ dsum, float* __restrict__ dmul) {
  extern __shared__ float sb[];    // [Cpad] block-level bias partials | [Cpad] sum of glev * yhead (ImplicitM gradient)
  float* sy = sb + Cpad;
  const int C = na * ch;
  for (int c = threadIdx.x; c < 2 * Cpad; c += blockDim.x) sb[c] = 0.f;
  __syncthreads();
  const int chunks = Cpad >> 3;
  const int rows = blockDim.x / chunks;
  const int g = threadIdx.x % chunks, r = threadIdx.x / chunks;
  const long long npix = (long long)B * H * W;
  const long long plane = (long long)H * W * ch;
  if (r < rows) {
    long long off[8];
    float m[8], acc[8], accy[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int c = 8 * g + j;
      const int a = c < C ? c / ch : 0, k = c < C ? c - a * ch : 0;
      off[j] = c < C ? (long long)a * plane + k : -1;
      m[j] = (mul && c < C) ? mul[c] : 1.f;
      acc[j] = 0.f;
      accy[j] = 0.f;
    }
    const long long hw = (long long)H * W;
    for (long long pix = (long long)blockIdx.x * rows + r; pix < npix; pix += (long long)gridDim.x * rows) {
      const long long b = pix / hw, p = pix - b * hw;
      const float* src = glev + b * na * plane + p * ch;
      float gv[8], v[8];
#pragma unroll
      for (int j = 0; j < 8; j++) gv[j] = off[j] >= 0 ? __ldg(src + off[j]) : 0.f;
      if (yhead) {                                  // yolov7: d ImplicitM_c = sum glev * pre_c, pre = yhead / mul
        const float* ys = yhead + b * na * plane + p * ch;
#pragma unroll
        for (int j = 0; j < 8; j++)
          if (off[j] >= 0) accy[j] = fmaf(gv[j], __ldg(ys + off[j]), accy[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; j++) { v[j] = gv[j] * m[j]; acc[j] += v[j]; }
      *reinterpret_cast<uint4*>(out + pix * Cpad + 8 * g) = pack8(v);
    }
#pragma unroll
    for (int j = 0; j < 8; j++)
      if (off[j] >= 0) {
        atomicAdd(&sb[8 * g + j], acc[j]);
        if (yhead) atomicAdd(&sy[8 * g + j], accy[j]);
      }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    if (dbias) atomicAdd(dbias + c, sb[c]);
    if (dsum) atomicAdd(dsum + c, sb[c]);
    if (yhead && dmul) atomicAdd(dmul + c, sy[c] / mul[c]);
  }
}

__global__ void __launch_bounds__(256)
sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, long long n, float lr,
           float momentum, float wd, int nesterov, int first) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i] + wd * p[i];
    float b = first ? gi : momentum * buf[i] + gi;
    buf[i] = b;
    p[i] -= lr * (nesterov ? gi + momentum * b : b);
  }
}

// Same update with the learning rate read from device memory: a step captured in a CUDA graph keeps following the
// warm-up / one-cycle schedule (train.py:189-193,220) by rewriting one float between replays.
__global__ void __launch_bounds__(256)
sgd_lrdev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, long long n,
                 const float* __restrict__ lr_dev, float momentum, float wd, int nesterov) {
  const float lr = *lr_dev;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i] + wd * p[i];
    float b = momentum * buf[i] + gi;
    buf[i] = b;
    p[i] -= lr * (nesterov ? gi + momentum * b : b);
  }
}

// torch.optim.Adam (train.py:154): m = b1*m + (1-b1)*g; v = b2*v + (1-b2)*g*g; p -= lr * (m/(1-b1^t)) / (sqrt(v/(1-b2^t)) + eps)
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            long long n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] + wd * p[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= (lr / bc1) * mi / (sqrtf(vi) / sqrtf(bc2) + eps);
  }
}

inline int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = 148ll * 32;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" {

// Backward of act(BatchNorm2d_train(raw)) (+ gradient already flowing to a residual is the caller's business).
// NOTE: dout is overwritten with dY = dout * act'(.) (it is dead afterwards).
//   dout, raw: bf16 NHWC views over P pixels x C channels; scale/shift/mean/invstd: fp32[C] saved by the forward pass
//   sums: fp32[2C] scratch, zeroed;  draw: bf16 view;  dgamma / dbeta: fp32[C], accumulated into (nullable)
int ryolo_bn_act_bwd(void* dout, long long dp, const void* raw, long long rp, const float* scale,
                     const float* shift, const float* mean, const float* invstd, int act, long long P, int C,
                     float* sums, void* draw, long long op, float* dgamma, float* dbeta, void* dres, long long resp,
                     void* stream) {
  RY_CHECK_ARG(C % 8 == 0 && C >= 8 && C <= 2048 && dp % 8 == 0 && rp % 8 == 0 && op % 8 == 0,
               "bn_act_bwd: channels must be a multiple of 8 in [8, 2048]");
  RY_CHECK_ARG(!dres || (resp % 8 == 0 && (((uintptr_t)dres) & 15) == 0), "bn_act_bwd: bad residual-gradient view");
  RY_CHECK_ARG(P < (1ll << 31) - (1 << 20), "bn_act_bwd: more than 2^31 pixels");
  if (P == 0) return RYOLO_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int groups = C / 8;
  const int threads = groups >= 256 ? groups : 256;
  const int rows = threads / groups;
  const size_t smem = (size_t)2 * rows * C * sizeof(float);
  long long want = (P + 2ll * rows - 1) / (2ll * rows);
  const int blocks = (int)(want > 148 * 8 ? 148 * 8 : (want < 1 ? 1 : want));
  __nv_bfloat16* d = (__nv_bfloat16*)dout;
  const __nv_bfloat16* r = (const __nv_bfloat16*)raw;
  __nv_bfloat16* o = (__nv_bfloat16*)draw;
  const int variant = threads == 256 ? ryolo_knob(RYOLO_KNOB_BN_BWD) : 0;
  // variant 3: persistent grid = resident capacity (2 blocks of 256 threads per SM), two rows per trip
  long long want3 = (P + 2ll * rows - 1) / (2ll * rows);
  const int blocks3 = (int)(want3 > 2ll * ry_sm_count() ? 2ll * ry_sm_count() : (want3 < 1 ? 1 : want3));
  const size_t smem3 = C <= 768 ? (size_t)2 * C * sizeof(float) : 0;
  const bool lean = ryolo_knob(RYOLO_KNOB_EW_REGS) != 0;
  // Blocks of these kernels are meant to share SMs with a resident wgrad CTA, which runs under the maximum shared-memory
  // carveout; an SM serves one carveout at a time, so the co-tenants must ask for the same one.
  static bool carved = false;
  if (!carved && ryolo_knob(RYOLO_KNOB_EW_REGS) >= 1 && ryolo_knob(RYOLO_KNOB_EW_REGS) != 2) {
#define RY_CARVE(ACT)                                                                                                   \
    cudaFuncSetAttribute(bn_act_bwd_reduce4_kernel<ACT, 304>, cudaFuncAttributePreferredSharedMemoryCarveout,            \
                         cudaSharedmemCarveoutMaxShared);                                                                \
    cudaFuncSetAttribute(bn_act_bwd_apply3_kernel<ACT, 304>, cudaFuncAttributePreferredSharedMemoryCarveout,             \
                         cudaSharedmemCarveoutMaxShared);
    RY_CARVE(RYOLO_ACT_LINEAR) RY_CARVE(RYOLO_ACT_LEAKY) RY_CARVE(RYOLO_ACT_MISH) RY_CARVE(RYOLO_ACT_SWISH)
#undef RY_CARVE
    carved = true;
  }
  // one launch for both passes when the operands fit in L2 and the grid is resident (knob bn_fuse, elements <= 24 M)
  const bool fuse = variant == 3 && !lean && ryolo_knob(RYOLO_KNOB_BN_FUSE) != 0 && P * C <= 24ll * 1000 * 1000;
  // the copy of d out into a residual operand's gradient rides in variant 3's apply pass; the other variants (A/B
  // switches) do it as the separate pass it used to be, BEFORE d out can be overwritten
  __nv_bfloat16* q = (variant == 3 && !fuse) ? (__nv_bfloat16*)dres : nullptr;
  if (dres && !q)
    add_into_kernel<<<grid_for(P * (C / 8), 256), 256, 0, st>>>((__nv_bfloat16*)dres, resp, (const __nv_bfloat16*)dout, dp, P,
                                                                C, 0);
#define RY_BWD(ACT)                                                                                                  \
  if (fuse) {                                                                                                        \
    ry_launch(bn_act_bwd_fused_kernel<ACT>, dim3(blocks3), dim3(threads), smem3, st, (const __nv_bfloat16*)d, dp, r, rp, \
              scale, shift, mean, invstd, sums, P, C, o, op, dgamma, dbeta);                                         \
  } else if (variant == 3 && lean) {                                                                                        \
    ry_launch(bn_act_bwd_reduce4_kernel<ACT, 304>, dim3(blocks3), dim3(threads), smem3, st, (const __nv_bfloat16*)d, dp, \
              r, rp, scale, shift, mean, invstd, P, C, sums);                                                        \
    ry_launch(bn_act_bwd_apply3_kernel<ACT, 304>, dim3(blocks3), dim3(threads), 0, st, (const __nv_bfloat16*)d, dp, r, \
              rp, scale, shift, mean, invstd, (const float*)sums, P, C, o, op, dgamma, dbeta, q, resp);              \
  } else if (variant == 3) {                                                                                         \
    ry_launch(bn_act_bwd_reduce4_kernel<ACT, 256>, dim3(blocks3), dim3(threads), smem3, st, (const __nv_bfloat16*)d, dp, \
              r, rp, scale, shift, mean, invstd, P, C, sums);                                                        \
    ry_launch(bn_act_bwd_apply3_kernel<ACT, 256>, dim3(blocks3), dim3(threads), 0, st, (const __nv_bfloat16*)d, dp, r, \
              rp, scale, shift, mean, invstd, (const float*)sums, P, C, o, op, dgamma, dbeta, q, resp);              \
  } else if (variant == 2) {                                                                                                \
    ry_launch(bn_act_bwd_reduce3_kernel<ACT>, dim3(blocks), dim3(threads), smem, st, (const __nv_bfloat16*)d, dp, r, rp, \
              scale, shift, mean, invstd, P, C, sums);                                                               \
    ry_launch(bn_act_bwd_apply2_kernel<ACT>, dim3(blocks * 2), dim3(threads), 0, st, (const __nv_bfloat16*)d, dp, r, rp, \
              scale, shift, mean, invstd, (const float*)sums, P, C, o, op, dgamma, dbeta);                           \
  } else {                                                                                                           \
    if (variant == 1)                                                                                                \
      ry_launch(bn_act_bwd_reduce2_kernel<ACT>, dim3(blocks), dim3(threads), smem, st, d, dp, r, rp, scale, shift,   \
                mean, invstd, P, C, sums);                                                                           \
    else                                                                                                             \
      ry_launch(bn_act_bwd_reduce_kernel<ACT>, dim3(blocks), dim3(threads), smem, st, d, dp, r, rp, scale, shift,    \
                mean, invstd, P, C, sums);                                                                           \
    ry_launch(bn_act_bwd_apply_kernel, dim3(blocks * 2), dim3(threads), 0, st, (const __nv_bfloat16*)d, dp, r, rp,   \
              scale, mean, invstd, (const float*)sums, P, C, o, op, dgamma, dbeta);                                  \
  }
  switch (act) {
    case RYOLO_ACT_LEAKY: RY_BWD(RYOLO_ACT_LEAKY) break;
    case RYOLO_ACT_MISH: RY_BWD(RYOLO_ACT_MISH) break;
    case RYOLO_ACT_SWISH: RY_BWD(RYOLO_ACT_SWISH) break;
    default: RY_BWD(RYOLO_ACT_LINEAR) break;
  }
#undef RY_BWD
  RY_CHECK_LAUNCH();
  return RYOLO_OK;
}

int ryolo_act_bwd2(const void* dout, long long dp, const void* x1, long long p1, const float* s1, const float* b1,
                   const void* x2, long long p2, const float* s2, const float* b2, int act, void* ds, long long op,
                   long long P, int C, void* stream) {
  RY_CHECK_ARG(C % 8 == 0, "act_bwd2: channels must be a multiple of 8");
  if (P == 0) return RYOLO_OK;
  const int g = grid_for(P * (C / 8), 256);
  cudaStream_t st = (cudaStream_t)stream;
#define RY_A2(ACT)                                                                                               \
  act_bwd2_kernel<ACT><<<g, 256, 0, st>>>((const __nv_bfloat16*)dout, dp, (const __nv_bfloat16*)x1, p1, s1, b1,  \
                                          (const __nv_bfloat16*)x2, p2, s2, b2, (__nv_bfloat16*)ds, op, P, C);
  switch (act) {
    case RYOLO_ACT_LEAKY: RY_A2(RYOLO_ACT_LEAKY) break;
    case RYOLO_ACT_MISH: RY_A2(RYOLO_ACT_MISH) break;
    case RYOLO_ACT_SWISH: RY_A2(RYOLO_ACT_SWISH) break;
    default: RY_A2(RYOLO_ACT_LINEAR) break;
  }
#undef RY_A2
  RY_CHECK_LAUNCH();
  return RYOLO_OK;
}

int ryolo_add_into(void* dst, long long dpitch, const void* src, long long sp, long long P, int C, int accumulate,
                   void* stream) {
  RY_CHECK_ARG(C % 8 == 0 && dpitch % 8 == 0 && sp % 8 == 0, "add_into: channels must be multiples of 8");
  if (P == 0) return RYOLO_OK;
  add_into_kernel<<<grid_for(P * (C / 8), 256), 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)dst, dpitch,
                                                                                (const __nv_bfloat16*)src, sp, P, C,
                                                                                accumulate);
  RY_CHECK_LAUNCH();
  return RYOLO_OK;
}

// SPP backward (model/utils.py:231-241): the three stride-1 "same" pools (odd k0 / k1 / k2) of ONE small map in one launch.
//   dx (+)= route(dy0, k0) + route(dy1, k1) + route(dy2, k2);  the dy views share one channel pitch (slices of the concat buffer)
// Returns RYOLO_ERR_INVALID when the map does not fit the shared-memory kernel (callers then use ryolo_maxpool_bwd per pool).
int ryolo_spp_bwd(const void* x, long long xp, const void* dy0, const void* dy1, const void* dy2, long long dyp, int N, int H,
                  int W, int C, int k0, int k1, int k2, void* dx, long long dxp, int accumulate, void* stream) {
  RY_CHECK_ARG(C % 8 == 0 && (k0 & 1) && (k1 & 1) && (k2 & 1) && k0 >= 1 && k1 >= 1 && k2 >= 1, "spp_bwd: bad arguments");
  RY_CHECK_ARG(H * W <= 1024 && W <= 256 && (long long)N * (C / 8) < (1ll << 31), "spp_bwd: map too large for the fused kernel");
  if ((long long)N * H * W == 0) return RYOLO_OK;
  const size_t smem = (size_t)26 * H * W * sizeof(float);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(maxpool_same_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         104 * 1024);
    if (e != cudaSuccess) { ryolo_set_error(cudaGetErrorString(e)); return RYOLO_ERR_CUDA; }
    configured = true;
  }
  SppGrads g{};
  g.dy[0] = (const __nv_bfloat16*)dy0; g.dy[1] = (const __nv_bfloat16*)dy1; g.dy[2] = (const __nv_bfloat16*)dy2;
  g.k[0] = k0; g.k[1] = k1; g.k[2] = k2; g.n = 3;
  maxpool_same_small_bwd_kernel<<<(unsigned)(N * (C / 8)), 256, smem, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, xp, g, dyp, H, W, C, (__nv_bfloat16*)dx, dxp, accumulate);
  RY_CHECK_LAUNCH();
  return RYOLO_OK;
}

// scratch: fp32 [N*H*W*C] (16-byte aligned); it is zeroed, filled with the routed gradients, then dx (+)= scratch.
int ryolo_maxpool_bwd(const void* x, long long xp, const void* dy, long long dyp, int N, int H, int W, int C, int k,
                      int stride, int pad, void* dx, long long dxp, int accumulate, float* scratch, void* stream) {
  RY_CHECK_ARG(C % 8 == 0 && k >= 1 && stride >= 1 && scratch, "maxpool_bwd: bad arguments");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const long long total = (long long)N * Ho * Wo * (C / 8);
  const long long P = (long long)N * H * W;
  if (P == 0) return RYOLO_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (stride == 1 && (k & 1) && pad == k / 2 && H * W <= 1024 && W <= 256 && (long long)N * (C / 8) < (1ll << 31)) {
    const size_t smem = (size_t)26 * H * W * sizeof(float);          // three fp32 [HW][8] tiles + [HW][8] bytes: <= 104 KB
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(maxpool_same_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           104 * 1024);
      if (e != cudaSuccess) { ryolo_set_error(cudaGetErrorString(e)); return RYOLO_ERR_CUDA; }
      configured = true;
    }
    SppGrads g{};
    g.dy[0] = (const __nv_bfloat16*)dy; g.k[0] = k; g.n = 1;
    maxpool_same_small_bwd_kernel<<<(unsigned)(N * (C / 8)), 256, smem, st>>>(
        (const __nv_bfloat16*)x, xp, g, dyp, H, W, C, (__nv_bfloat16*)dx, dxp, accumulate);
    RY_CHECK_LAUNCH();
    return RYOLO_OK;
  }
  if (stride == k && pad == 0 && H % k == 0 && W % k == 0) {
    maxpool_tiled_bwd_kernel<<<grid_for(total, 256), 256, 0, st>>>((const __nv_bfloat16*)x, xp, (const __nv_bfloat16*)dy,
                                                                  dyp, N, H, W, C, k, (__nv_bfloat16*)dx, dxp, accumulate);
    RY_CHECK_LAUNCH();
    return RYOLO_OK;
  }
  cudaMemsetAsync(scratch, 0, (size_t)P * C * sizeof(float), st);
  if (total > 0)
    maxpool_bwd_kernel<<<grid_for(total, 256), 256, 0, st>>>((const __nv_bfloat16*)x, xp, (const __nv_bfloat16*)dy, dyp, N,
                                                            H, W, C, k, stride, pad, Ho, Wo, scratch);
  add_f32_into_kernel<<<grid_for(P * (C / 8), 256), 256, 0, st>>>((__nv_bfloat16*)dx, dxp, scratch, P, C, accumulate);
  RY_CHECK_LAUNCH();
  return RYOLO_OK;
}

int ryolo_upsample2x_bwd(const void* dy, long long dyp, int N, int H, int W, int C, void* dx, long long dxp,
                         int accumulate, void* stream) {
  RY_CHECK_ARG(C % 8 == 0, "upsample2x_bwd: channels must be a multiple of 8");
  const long long total = (long long)N * H * W * (C / 8);
  if (total == 0) return RYOLO_OK;
  upsample2x_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dy, dyp, N, H, W,
                                                                                C, (__nv_bfloat16*)dx, dxp, accumulate);
  RY_CHECK_LAUNCH();
  return RYOLO_OK;
}

int ryolo_sgd_step(float* param, const float* grad, float* buf, long long n, float lr, float momentum,
                   float weight_decay, int nesterov, int first, void* stream) {
  if (n <= 0) return RYOLO_OK;
  sgd_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(param, grad, buf, n, lr, momentum, weight_decay,
                                                                nesterov, first);
  RY_CHECK_LAUNCH();
  return RYOLO_OK;
}

int ryolo_sgd_step_lrdev(float* param, const float* grad, float* buf, long long n, const float* lr_dev, float momentum,
                         float weight_decay, int nesterov, void* stream) {
  RY_CHECK_ARG(lr_dev != nullptr, "sgd_step_lrdev: lr_dev must point at a device float");
  if (n <= 0) return RYOLO_OK;
  sgd_lrdev_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(param, grad, buf, n, lr_dev, momentum, weight_decay,
                                                                      nesterov);
  RY_CHECK_LAUNCH();
  return RYOLO_OK;
}

int ryolo_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                    float beta1, float beta2, float eps, float weight_decay, int step, void* stream) {
  RY_CHECK_ARG(step >= 1, "adam_step: step counts from 1");
  if (n <= 0) return RYOLO_OK;
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
  adam_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2,
                                                                 eps, weight_decay, bc1, bc2);
  RY_CHECK_LAUNCH();
  return RYOLO_OK;
}

int ryolo_head_grad_pack(const float* glev, int B, int na, int H, int W, int ch, int Cpad, const float* mul, void* out,
                         float* dbias, const float* yhead, float* dsum, float* dmul, void* stream) {
  RY_CHECK_ARG(Cpad % 8 == 0 && Cpad >= na * ch && Cpad <= 2048, "head_grad_pack: bad padded channel count");
  RY_CHECK_ARG(!yhead || (mul && dmul), "head_grad_pack: the ImplicitM gradient needs mul and dmul");
  RY_CHECK_ARG((((uintptr_t)out) & 15) == 0, "head_grad_pack: output must be 16-byte aligned");
  const long long npix = (long long)B * H * W;
  if (npix == 0) return RYOLO_OK;
  const int rows = 256 / (Cpad / 8);
  long long want = (npix + 4ll * rows - 1) / (4ll * rows);
  const int blocks = (int)(want > 148 * 8 ? 148 * 8 : (want < 1 ? 1 : want));
  head_grad_pack_kernel<<<blocks, 256, (size_t)2 * Cpad * sizeof(float), (cudaStream_t)stream>>>(
      glev, B, na, H, W, ch, Cpad, mul, (__nv_bfloat16*)out, dbias, yhead, dsum, dmul);
  RY_CHECK_LAUNCH();
  return RYOLO_OK;
}

}  // extern "C"
