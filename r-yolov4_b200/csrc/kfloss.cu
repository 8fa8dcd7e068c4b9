// Standalone KFLoss (BASELINE.json configs[3]) — split from loss.cu so that it is NOT compiled with --fmad=false:
// nothing here is compared bit for bit (the assignment code in loss.cu is), the closed form is ~250 instructions per
// pair and at 64 B/pair the kernel sits on the instruction-issue roof, where every un-fused multiply-add costs a slot.
#define RY_KF_FAST_MATH 1
#include "common.cuh"
#include "loss_math.cuh"
#include "ryolo_b200.h"

using namespace ryolo;

namespace {

// ------------------------------------------------------------------------------------------ standalone KFLoss
// KFLoss.forward (lib/loss.py:100-150) on N (pred, target) pairs, value + KFIoU + gradient in one pass.
// 256 pairs per block: the [256,5] slabs are moved with coalesced 128-bit accesses through shared memory
// (44 B/pair forward, 64 B/pair with the gradient).
template <bool GRAD>
__global__ void __launch_bounds__(256)
kfloss_pairs_kernel(const float* __restrict__ pred, const float* __restrict__ target, long long N,
                    float* __restrict__ kfiou, float* __restrict__ grad, double* __restrict__ sums) {
  // Persistent blocks walk 256-pair slabs.  The next slab's 2 x 320 float4 are fetched into registers before the
  // current slab's closed form runs (HBM latency hides behind ~250 instructions per pair), and a block issues its two
  // fp64 atomics once, not once per slab (50k blocks x 2 same-address atomics used to serialise in one L2 slice).
  __shared__ __align__(16) float sp[256 * 5], st[256 * 5];
  __shared__ double wpart[2][8];
  const int tid = threadIdx.x;
  const long long nslabs = (N + 255) / 256;
  const float invN = 1.f / (float)N;
  double a_xy = 0.0, a_kf = 0.0;
  float4 rp[2], rt[2];
  auto fetch = [&](long long slab) {
    const long long base = slab * 256;
    const int n = (int)min((long long)256, N - base);
    const int nvec = (n * 5) >> 2;
    const float4* p4 = reinterpret_cast<const float4*>(pred + base * 5);
    const float4* t4 = reinterpret_cast<const float4*>(target + base * 5);
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const int i = tid + u * 256;
      if (i < nvec) { rp[u] = ry_ld_stream(p4 + i); rt[u] = ry_ld_stream(t4 + i); }
    }
  };
  long long slab = blockIdx.x;
  if (slab < nslabs) fetch(slab);
  for (; slab < nslabs; slab += gridDim.x) {
    const long long base = slab * 256;
    const int n = (int)min((long long)256, N - base);
    const int nvec = (n * 5) >> 2;
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const int i = tid + u * 256;
      if (i < nvec) { reinterpret_cast<float4*>(sp)[i] = rp[u]; reinterpret_cast<float4*>(st)[i] = rt[u]; }
    }
    for (int i = nvec * 4 + tid; i < n * 5; i += 256) { sp[i] = pred[base * 5 + i]; st[i] = target[base * 5 + i]; }
    __syncthreads();
    if (slab + gridDim.x < nslabs) fetch(slab + gridDim.x);
    Box5 g = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (tid < n) {
      const Box5 pb = {sp[tid * 5], sp[tid * 5 + 1], sp[tid * 5 + 2], sp[tid * 5 + 3], sp[tid * 5 + 4]};
      const Box5 tb = {st[tid * 5], st[tid * 5 + 1], st[tid * 5 + 2], st[tid * 5 + 3], st[tid * 5 + 4]};
      float xl, kl, k;
      kf_fwd_bwd(pb, tb, &xl, &kl, &k, GRAD ? &g : nullptr);
      kfiou[base + tid] = k;
      a_xy += (double)xl;
      a_kf += (double)kl;
    }
    __syncthreads();                         // everyone has read its pair: the slabs can be overwritten
    if (GRAD) {
      if (tid < n) {
        sp[tid * 5] = g.x * invN; sp[tid * 5 + 1] = g.y * invN; sp[tid * 5 + 2] = g.w * invN;
        sp[tid * 5 + 3] = g.h * invN; sp[tid * 5 + 4] = g.r * invN;
      }
      __syncthreads();
      float4* g4 = reinterpret_cast<float4*>(grad + base * 5);
      for (int i = tid; i < nvec; i += 256) g4[i] = reinterpret_cast<const float4*>(sp)[i];
      for (int i = nvec * 4 + tid; i < n * 5; i += 256) grad[base * 5 + i] = sp[i];
      __syncthreads();
    }
  }
  a_xy = ry_warp_sum_d(a_xy);
  a_kf = ry_warp_sum_d(a_kf);
  if ((tid & 31) == 0) { wpart[0][tid >> 5] = a_xy; wpart[1][tid >> 5] = a_kf; }
  __syncthreads();
  if (tid == 0) {
    double x = 0.0, k = 0.0;
    for (int w = 0; w < 8; w++) { x += wpart[0][w]; k += wpart[1][w]; }
    atomicAdd(&sums[0], x);
    atomicAdd(&sums[1], k);
  }
}

__global__ void kfloss_finalize_kernel(const double* __restrict__ sums, long long N, float* __restrict__ loss) {
  loss[0] = N > 0 ? (float)((sums[0] + sums[1]) / (double)N) : 0.f;   // mean(xy_loss) + mean(kf_loss), see oracle
}

}  // namespace

extern "C" {

// KFLoss()(pred, target) -> (loss, KFIoU)  (lib/loss.py:81-150); grad = d loss / d pred or NULL.
// pred/target/grad: device fp32 [N,5] (x, y, w, h, theta rad), 16-byte aligned; kfiou [N]; loss [1];
// workspace: 16 bytes (two doubles).
int ryolo_kfloss(const float* pred, const float* target, int64_t N, float* kfiou, float* grad, float* loss,
                 void* workspace, size_t ws_bytes, void* stream) {
  RY_CHECK_ARG(N >= 0 && ws_bytes >= 16, "kfloss: bad arguments");
  RY_CHECK_ARG((((uintptr_t)pred) & 15) == 0 && (((uintptr_t)target) & 15) == 0 && (((uintptr_t)grad) & 15) == 0,
               "kfloss: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  double* sums = (double*)workspace;
  cudaMemsetAsync(sums, 0, 16, st);
  if (N > 0) {
    long long nslabs = (N + 255) / 256;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    static int occ[2] = {0, 0};                       // resident blocks per SM: one full wave of persistent blocks
    if (!occ[0]) {
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[0], kfloss_pairs_kernel<false>, 256, 0);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[1], kfloss_pairs_kernel<true>, 256, 0);
      if (occ[0] < 1) occ[0] = 1;
      if (occ[1] < 1) occ[1] = 1;
    }
    const long long wave = (long long)sms * occ[grad ? 1 : 0];
    const unsigned blocks = (unsigned)(nslabs < wave ? nslabs : wave);
    if (grad) kfloss_pairs_kernel<true><<<blocks, 256, 0, st>>>(pred, target, N, kfiou, grad, sums);
    else kfloss_pairs_kernel<false><<<blocks, 256, 0, st>>>(pred, target, N, kfiou, grad, sums);
  }
  kfloss_finalize_kernel<<<1, 1, 0, st>>>(sums, N, loss);
  RY_CHECK_LAUNCH();
  return RYOLO_OK;
}

}  // extern "C"
