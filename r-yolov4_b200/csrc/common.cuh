// Shared helpers for the ryolo_b200 CUDA translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define RYOLO_OK 0
#define RYOLO_ERR_INVALID 1
#define RYOLO_ERR_CUDA 2
#define RYOLO_ERR_WORKSPACE 3

extern "C" void ryolo_set_error(const char* msg);

// process-wide tuning knobs (lib.cu); ids index ryolo_tune's keys
enum { RYOLO_KNOB_HALO = 0, RYOLO_KNOB_DBG, RYOLO_KNOB_WG_SPLIT, RYOLO_KNOB_WG_DBG, RYOLO_KNOB_EPI_TMA,
       RYOLO_KNOB_EPI_MAXBN, RYOLO_KNOB_WG_TAPGRP, RYOLO_KNOB_BN_BWD, RYOLO_KNOB_WG_TRANS, RYOLO_KNOB_SW64, RYOLO_KNOB_NACC, RYOLO_KNOB_PDL, RYOLO_KNOB_SSA, RYOLO_KNOB_WG_BOXES, RYOLO_KNOB_EW_REGS, RYOLO_KNOB_NMS_BAND, RYOLO_KNOB_BN_FUSE, RYOLO_KNOB_KGRP, RYOLO_KNOB_PAIR, RYOLO_KNOB_WRES, RYOLO_KNOB_WG_X32, RYOLO_KNOB_COUNT };
extern "C" int ryolo_knob(int id);
extern "C" int ry_sm_count(void);      // multiprocessors of the current device (cached)

#define RY_CHECK_ARG(cond, msg)   \
  do {                            \
    if (!(cond)) {                \
      ryolo_set_error(msg);       \
      return RYOLO_ERR_INVALID;   \
    }                             \
  } while (0)

#define RY_CHECK_LAUNCH()                            \
  do {                                               \
    cudaError_t e__ = cudaGetLastError();            \
    if (e__ != cudaSuccess) {                        \
      ryolo_set_error(cudaGetErrorString(e__));      \
      return RYOLO_ERR_CUDA;                         \
    }                                                \
  } while (0)

// Programmatic dependent launch: with the attribute set, a kernel may be scheduled while the previous kernel of its
// stream is still draining (as soon as every CTA of that kernel has executed griddepcontrol.launch_dependents or
// exited); it runs its prologue (barrier init, TMEM allocation, descriptor prefetch) and must execute
// griddepcontrol.wait before it reads or writes global memory.  Both instructions are no-ops in a normal launch.
#ifdef __CUDACC__
__device__ __forceinline__ void ry_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void ry_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t ry_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                    Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = ryolo_knob(RYOLO_KNOB_PDL) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}
#endif

// fused BatchNorm statistics: fixed-point scales of the cross-CTA 64-bit accumulators (|sum| < 2^39, sum of squares < 2^43)
#define RY_BN_SUM_SCALE 16777216.f   /* 2^24 */
#define RY_BN_SQ_SCALE 1048576.f     /* 2^20 */

static inline size_t ry_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__device__ __forceinline__ float ry_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double ry_warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 1/d for d in [1, 1e30] on the FMA pipe (magic-constant seed + 2 Newton steps, relative error < 3e-4: below bf16's
// 2^-9).  The HBM-bound bf16 activation kernels spend two MUFU ops per element on Mish (ex2 + rcp) at 16 MUFU/clk/SM;
// moving the reciprocal to the 128-lane FMA pipe halves their SFU time.
__device__ __forceinline__ float ry_rcp_fma(float d) {
  float r = __int_as_float(0x7EF311C7 - __float_as_int(d));
  r = r * fmaf(-d, r, 2.f);
  r = r * fmaf(-d, r, 2.f);
  return r;
}

// order-preserving map float -> uint32 (ascending)
__device__ __forceinline__ uint32_t ry_float_order(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ry_order_float(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(u);
}

// streaming 128-bit load that does not pollute L1
__device__ __forceinline__ float4 ry_ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
