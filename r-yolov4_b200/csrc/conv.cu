// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, operands
// staged by TMA) — K1 in SURVEY.md §2.1.  Replaces nn.Conv2d(+BatchNorm2d eval-folded +activation) of
// the reference's Conv block (model/utils.py:6-32) for NHWC bf16 activations.
//
//   GEMM view      D[M = pixels, N = Cout] = sum over taps (kh,kw) and Cin chunks of  A_tap[M, 64] * W[N, 64]^T
//   M tile         128 rows = a TH x TW spatial patch of ONE image, so that the A operand of tap (kh,kw)
//                  is a single 4-D TMA box of the NHWC input at coordinates shifted by (kh-pad, kw-pad):
//                  zero padding falls out of TMA's out-of-bounds fill, stride-2 convs use the tensor map's
//                  element strides.  No im2col buffer ever exists.
//   operands       bf16, K-major, 128-byte swizzle (TMA writes it, the UMMA descriptor reads it); 64-byte swizzle and
//                  32-element K blocks when Cin == 32 (no overhanging TMA boxes)
//   pipeline       persistent CTAs (one per SM); warp 0: TMA producer | warp 1: TMEM alloc + MMA issue (warp-uniform
//                  loop, one elected lane) into 2-8 rotating TMEM accumulators | warps 2-5 / 6-9: two epilogue groups
//                  taking alternate tiles (tcgen05.ld -> scale/shift/activation/residual/BN statistics -> global)
//                  overlapping the following tiles' main loops; mbarrier smem ring; programmatic dependent launch
//   ring variants  (compile-time, chosen per layer in launch()): 1-3 K blocks per stage (narrow tiles) | cluster pairs:
//                  two CTAs on neighbouring m-tiles share every weight tile through TMA multicast (deep N = 256
//                  layers) | resident weights: the whole weight matrix loaded once per CTA, activation-only ring
//                  (one n-tile, <= 96 KB of weights, many tiles per CTA)
//   epilogue       mode 0: bf16 NHWC into a channel slice of a (concat) buffer, staged in smem and stored (or, for
//                  dgrad's accumulation, reduce-added) by TMA; per-channel scale/shift = folded BatchNorm (eval) or
//                  identity (train: raw conv output + fused BatchNorm statistics and finalize);
//                  mode 1: fp32 head written directly in the reference's [B, na, gs, gs, ch] layout
//                  (the permute+contiguous of model/yololayer.py:25,76 is fused away).
#include "common.cuh"
#include "ryolo_b200.h"
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

namespace {

constexpr int kBM = 128;        // UMMA M (rows of the patch tile, TH*TW <= 128)
constexpr int kBK = 64;         // bf16 elements per k-block = one 128-byte swizzle row
constexpr int kThreads = 320;   // 10 warps: TMA, MMA, 2 x 4 epilogue
constexpr int kMaxStages = 12;
constexpr int kMaxAcc = 8;
// fused BatchNorm statistics: fixed-point scales of the cross-CTA accumulators (|sum| < 2^39, sum of squares < 2^43)
constexpr float kSumScale = RY_BN_SUM_SCALE;   // 2^24
constexpr float kSqScale = RY_BN_SQ_SCALE;     // 2^20

// ------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// smem -> global tile store / reduction through TMA (bulk async-group completion)
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// Cluster-pair variants (PAIR kernels): one TMA load lands in the same smem offset of every CTA in `mask` and completes
// bytes on each CTA's own barrier at the same offset; one commit arrives on the same barrier offset of every CTA in `mask`.
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void group_bar(uint32_t id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
  return (uint32_t)v;
}
// One lane of a converged warp (the same one every time for the full mask).  Single-thread work (TMA / MMA issue) is
// written as `warp-uniform loop + if (elect_one())` rather than `if (lane == 0) { loop }`: under a lane test the
// compiler treats the whole loop as divergent code and wraps every uniform-datapath instruction (UTMALDG, UTCHMMA,
// UTCBAR) in an elect/branch retry sequence with R2UR copies; ~90 SASS instructions of dependent uniform ALU work per
// 4 MMAs made the issue thread the bottleneck of every tile narrower than N = 256 (profiles/r01_conv_issue_loop.txt).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) /* LBO (unused for swizzled K-major) */ |
         ((uint64_t)(1024 >> 4) << 32) /* SBO */ | ((uint64_t)1 << 46) /* descriptor version (sm_100) */ |
         ((uint64_t)2 << 61) /* SWIZZLE_128B */;
}
// K-major, SWIZZLE_64B: rows of 64 bytes (32 bf16), 8-row groups 512 bytes apart.  Used when Cin == 32: a 64-channel box
// over a 32-channel tensor overhangs it, and TMA serves overhanging rows on a ~2x slower path (tools/oob_probe.py).
__device__ __forceinline__ uint64_t umma_desc_k_sw64(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)4 << 61) /* SWIZZLE_64B */;
}
// Same layout, but the 8-row groups are `sbo_bytes` apart and the start may sit on any 128-byte row of a swizzle atom
// (shifted view into a halo tile): base_offset carries the row phase of the start address.
__device__ __forceinline__ uint64_t umma_desc_k_sw128_view(uint32_t saddr, uint32_t sbo_bytes, int use_base_offset) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) |
               ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
  if (use_base_offset) d |= (uint64_t)((saddr >> 7) & 7u) << 49;
  return d;
}
// instruction descriptor: D fp32, A/B bf16, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}

__device__ __forceinline__ float act_apply(float x, int act) {
  if (act == RYOLO_ACT_LEAKY) return x > 0.f ? x : 0.1f * x;
  if (act == RYOLO_ACT_MISH) {           // x * tanh(softplus(x)) = x * n / (n + 2), n = e^x (e^x + 2)
    if (x > 20.f) return x;
    const float e = __expf(x);
    const float n = e * (e + 2.f);
    return x * __fdividef(n, n + 2.f);
  }
  if (act == RYOLO_ACT_SWISH) return x * __fdividef(1.f, 1.f + __expf(-x));
  return x;
}

struct ConvKernelParams {
  int N, Ho, Wo, Cout, Cin;
  int TH, TW, tiles_h, tiles_w, n_tiles;
  int BN, stages;
  uint32_t tmem_cols;
  int nacc;                        // TMEM accumulators in rotation (2..8, even): hides the MMA <-> epilogue hand-back latency
  int ksize, stride, pad, kb_per_tap;
  int ntaps, Ktap;                 // taps of this launch; K elements per tap in the weight matrix
  int kbk;                         // K elements per K block: 64 (SWIZZLE_128B rows) or 32 (SWIZZLE_64B rows, Cin == 32)
  int kgrp;                        // K blocks per ring stage (one full / empty barrier hand-shake and one commit per stage)
  int pair;                        // 1: clusters of two CTAs share every weight tile (each loads half, multicast to both)
  int wres;                        // 1: the whole weight matrix is loaded once per CTA and stays in shared memory (WRES kernels)
  signed char tap_dh[9], tap_dw[9]; // input offset of each tap (rows / cols, input-lattice units)
  unsigned char tap_k[9];          // weight K-block index of each tap
  int dbg;                         // RYOLO_DBG timing experiments (wrong results): 1 no stores, 2 no BN statistics, 4 no MMAs, 8 no A loads, 16 no cross-CTA statistics tail, 32 no weight loads
  int halo;                        // 0 | 1 | 2: 3x3 stride-1 taps read shifted views of ONE (TH+2)x(TW+2) halo box (2: base_offset set)
  int a_slots; uint32_t a_slot_bytes;
  int epi_tma;                     // bf16 epilogue stages 64-channel slabs in smem and stores them with TMA (2: reduce-add)
  int out_s, out_oh, out_ow, OutH, OutW;   // output lattice: pixel (ho*out_s + out_oh, wo*out_s + out_ow) of an OutH x OutW map
  int mode, act;
  void* out;
  long long out_cpitch;
  const float* scale;
  const float* shift;
  const __nv_bfloat16* residual;
  long long res_cpitch;
  int head_na, head_ch;
  ryolo_bn_fuse bn;        // bn.partial != nullptr: fused train-mode BatchNorm statistics + finalize (EPI_RAW only)
  double bn_count;         // N*Ho*Wo
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Persistent, warp-specialised implicit-GEMM conv.  One CTA per SM walks tiles  t = blockIdx.x + i*gridDim.x
// (n-tile fastest, so CTAs running side by side share the activation patch in L2).  The TMA producer runs
// ahead across tile boundaries through a ring of `stages` smem slots; the MMA warp alternates between two
// TMEM accumulators so that the epilogue of tile i overlaps the main loop of tile i+1.
//
// EPI selects the epilogue at compile time so that its inner loops carry no per-element branches:
//   EPI_RAW     bf16 NHWC store of the bare accumulator (train mode: BatchNorm statistics come next)
//   EPI_AFFINE  y = ACT(acc*scale + shift) [+ residual] -> bf16 NHWC   (eval-folded BN, scale/shift staged in smem)
//   EPI_HEAD    y = acc*scale + shift -> fp32 [B, na, gs, gs, ch]
enum { EPI_RAW = 0, EPI_AFFINE = 1, EPI_HEAD = 2 };

template <int EPI, int ACT, bool DBG, int KG, bool PAIR, bool WRES>
__global__ void __launch_bounds__(kThreads, 1)
conv_fwd_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmO, const ConvKernelParams p) {
  // The timing-experiment switches only exist in the DBG instantiations: the production kernels' producer / issuer loops
  // are serial single-warp instruction chains in which every instruction costs ~5 cycles per K stage.
  const int dbg = DBG ? p.dbg : 0;
  const int BN = p.BN, STAGES = p.stages;
  const uint32_t kABytes = kBM * (uint32_t)p.kbk * 2, kBBytes = (uint32_t)BN * (uint32_t)p.kbk * 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B needs 1024-byte alignment
  const uint32_t sA = smem_base;
  constexpr int G = KG;                                              // K blocks per ring stage (compile time: the issue loops below
                                                                     // are single-warp serial chains, every instruction in them counts)
  const uint32_t stA = (uint32_t)G * kABytes, stB = (uint32_t)G * kBBytes;
  const uint32_t sB = smem_base + (p.halo ? (uint32_t)p.a_slots * p.a_slot_bytes : (uint32_t)STAGES * stA);
  // WRES: the layer's whole weight matrix (one n-tile, <= 96 KB) stays resident after sB for the life of the CTA, K block
  // (tap, cb) at sB + (tap*kb_per_tap + cb)*kBBytes; the ring then carries activation boxes only.
  const uint32_t sStage = sB + (WRES ? (uint32_t)(p.ntaps * p.kb_per_tap) * kBBytes
                                     : (uint32_t)STAGES * stB);   // epi_tma: one 128-row x 64-channel bf16 slab per epilogue group
  __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 2 * kMaxAcc + 9];
  __shared__ uint32_t tmem_slot;
  const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[kMaxStages]),
                 bar_acc_full = smem_u32(&bars[2 * kMaxStages]), bar_acc_empty = smem_u32(&bars[2 * kMaxStages + kMaxAcc]),
                 bar_afull = smem_u32(&bars[2 * kMaxStages + 2 * kMaxAcc]),
                 bar_aempty = smem_u32(&bars[2 * kMaxStages + 2 * kMaxAcc + 4]),
                 bar_w = smem_u32(&bars[2 * kMaxStages + 2 * kMaxAcc + 8]);      // WRES: resident weights have landed
  // warp index through a shuffle: the compiler then knows it is warp-uniform, keeps the role loops' state in uniform
  // registers and drops the per-stage R2UR / constant-bank reloads from the single-warp issue chains
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int KB = p.ntaps * p.kb_per_tap;
  const int total_tiles = p.N * p.tiles_h * p.tiles_w * p.n_tiles;
  // Tile walk.  Plain: CTA b takes tiles b, b + grid, ... (n-tile fastest).  PAIR (clusters of two CTAs): the pair takes
  // (n-tile, m-tile pair) items and rank r works on m-tile 2*mp + r, so both CTAs need the SAME weight tile at the same
  // time: each loads half of it and multicasts it to both (the L2 -> SM weight traffic of the pair halves).  An odd tile
  // count leaves rank 1 a ghost tile past the last image: TMA zero-fills its loads and clips its stores.
  const int rank = PAIR ? (int)cluster_ctarank() : 0;
  const int tile0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tstep = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int ntile = PAIR ? p.n_tiles * ((p.N * p.tiles_h * p.tiles_w + 1) >> 1) : total_tiles;
  __shared__ __align__(16) float s_aff[2 * 256];           // scale | shift of this CTA's n-tile (EPI_AFFINE / EPI_HEAD)
  __shared__ float s_tr[8 * 32 * 17];                      // per-warp 32x16-word transpose tile (BN statistics, head stores)
  __shared__ int s_last;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if (p.epi_tma) prefetch_tmap(&tmO);
    for (int s = 0; s < STAGES; s++) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, PAIR ? 2 : 1);    // PAIR: a slot is free once BOTH CTAs' MMAs have read it (the peer writes half of it)
    }
    for (int b = 0; b < p.nacc; b++) {
      mbar_init(bar_acc_full + 8 * b, 1);
      mbar_init(bar_acc_empty + 8 * b, 4);     // one arrival per epilogue warp
    }
    for (int b = 0; b < 4; b++) {
      mbar_init(bar_afull + 8 * b, 1);
      mbar_init(bar_aempty + 8 * b, 1);
    }
    mbar_init(bar_w, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_slot), p.tmem_cols);
    tmem_relinquish();
  }
  if (threadIdx.x == 0) ry_pdl_trigger();       // every CTA is resident from the start: the next kernel may queue up now
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();                 // the peer's barriers exist before anything arrives on them
  tc_fence_after();
  ry_pdl_wait();                                // nothing above touches global memory
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0 && !p.halo) {
    // ================================ TMA producer (warp-uniform loop, one elected lane issues) ===============
    const uint32_t a_bytes = (uint32_t)(p.TH * p.TW) * (uint32_t)p.kbk * 2;
    const uint32_t tx_bytes = (((dbg & 8) ? 0u : a_bytes) + ((WRES || (dbg & 32)) ? 0u : kBBytes)) * (uint32_t)G;
    // A stage carries G consecutive K blocks behind ONE full / empty barrier pair: the issue loops pay their fixed cost
    // (barrier wait, fence, elect, commit: ~0.2 us whatever the stage carries) once per G blocks.
    const int KS = KB / G;
    // every CTA walks the K blocks in its own rotation: otherwise all 148 CTAs request the SAME weight tile from the same
    // L2 lines at the same time (the sum is order-independent up to fp32 rounding, and the tile -> CTA map is static, so
    // results stay reproducible)
    const int krot = (int)(((unsigned)tile0 * 5u) % (unsigned)KS) * G;
    const uint32_t b_half = PAIR ? (uint32_t)rank * (kBBytes >> 1) : 0u;      // PAIR: this CTA loads rows [rank*BN/2, +BN/2) of the weight tile
    if (WRES && tile0 < ntile) {
      if (elect_one()) {
        mbar_expect_tx(bar_w, (uint32_t)KB * kBBytes);
        int kb = 0;
        for (int tp = 0; tp < p.ntaps; tp++)
          for (int c = 0; c < p.kb_per_tap; c++, kb++)
            tma_load_2d(sB + (uint32_t)kb * kBBytes, &tmB, bar_w, (int)p.tap_k[tp] * p.Ktap + c * p.kbk, 0);
      }
      __syncwarp();
    }
    int s = 0;
    uint32_t phase = 0;
    for (int t = tile0; t < ntile; t += tstep) {
      const int nt = t % p.n_tiles;
      int mt = t / p.n_tiles;
      if (PAIR) mt = 2 * mt + rank;
      const int pw = mt % p.tiles_w; mt /= p.tiles_w;
      const int ph = mt % p.tiles_h;
      const int img = mt / p.tiles_h;
      const int hs = ph * p.TH * p.stride, ws = pw * p.TW * p.stride, n0 = nt * BN;
      int tap = krot / p.kb_per_tap, cb = krot - tap * p.kb_per_tap;
      for (int ks = 0; ks < KS; ks++) {
        mbar_wait(bar_empty + 8 * s, phase ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(bar_full + 8 * s, tx_bytes);
          int tp = tap, c = cb;
          uint32_t da = sA + s * stA, db = sB + s * stB;
#pragma unroll
          for (int g = 0; g < G; g++) {
            if (!(dbg & 8))
              tma_load_4d(da, &tmA, bar_full + 8 * s, c * p.kbk, ws + p.tap_dw[tp], hs + p.tap_dh[tp], img);
            if (!WRES) {                        // (resident weights were loaded once, above)
              if (PAIR)
                tma_load_2d_mc(db + b_half, &tmB, bar_full + 8 * s, (int)p.tap_k[tp] * p.Ktap + c * p.kbk,
                               n0 + rank * (BN >> 1), (uint16_t)3);
              else if (!(dbg & 32))             // dbg 32 (timing experiment, wrong results): no weight loads
                tma_load_2d(db, &tmB, bar_full + 8 * s, (int)p.tap_k[tp] * p.Ktap + c * p.kbk, n0);
            }
            da += kABytes; db += kBBytes;
            if (++c == p.kb_per_tap) { c = 0; if (++tp == p.ntaps) tp = 0; }
          }
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; phase ^= 1u; }
#pragma unroll
        for (int g = 0; g < G; g++)
          if (++cb == p.kb_per_tap) { cb = 0; if (++tap == p.ntaps) tap = 0; }
      }
    }
  } else if (warp == 0) {
    // ================================ TMA producer, halo mode (one lane) =====================================
    if (lane == 0) {
      const uint32_t a_bytes = (uint32_t)(p.TH * p.TW) * (uint32_t)p.kbk * 2;
      const int krot = (int)((blockIdx.x * 5u) % (unsigned)KB);
      int s = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int nt = t % p.n_tiles;
        int mt = t / p.n_tiles;
        const int pw = mt % p.tiles_w; mt /= p.tiles_w;
        const int ph = mt % p.tiles_h;
        const int img = mt / p.tiles_h;
        const int hs = ph * p.TH * p.stride, ws = pw * p.TW * p.stride, n0 = nt * BN;
        if (p.halo) {
          // one (TH+2)x(TW+2) halo box per 64-channel chunk serves all 9 taps; weights stream through the B ring
          const uint32_t halo_bytes = (uint32_t)((p.TH + 2) * (p.TW + 2)) * kBK * 2;
          for (int cb = 0; cb < p.kb_per_tap; cb++) {
            mbar_wait(bar_aempty + 8 * as, aphase ^ 1u);
            mbar_expect_tx(bar_afull + 8 * as, halo_bytes);
            tma_load_4d(sA + as * p.a_slot_bytes, &tmA, bar_afull + 8 * as, cb * kBK, ws - 1, hs - 1, img);
            if (++as == p.a_slots) { as = 0; aphase ^= 1u; }
            for (int tap0 = 0; tap0 < p.ntaps; tap0++) {
              int tap = tap0 + (int)(blockIdx.x % 9u);
              if (tap >= p.ntaps) tap -= p.ntaps;
              mbar_wait(bar_empty + 8 * s, phase ^ 1u);
              mbar_expect_tx(bar_full + 8 * s, kBBytes);
              tma_load_2d(sB + s * kBBytes, &tmB, bar_full + 8 * s, (int)p.tap_k[tap] * p.Ktap + cb * kBK, n0);
              if (++s == STAGES) { s = 0; phase ^= 1u; }
            }
          }
          continue;
        }
        for (int kb0 = 0; kb0 < KB; kb0++) {
          // every CTA walks the K blocks in its own rotation: otherwise all 148 CTAs request the SAME weight tile
          // from the same L2 lines at the same time (the sum is order-independent up to fp32 rounding, and the
          // tile -> CTA map is static, so results stay reproducible)
          int kb = kb0 + krot;
          if (kb >= KB) kb -= KB;
          const int tap = kb / p.kb_per_tap, cb = kb - tap * p.kb_per_tap;
          mbar_wait(bar_empty + 8 * s, phase ^ 1u);
          mbar_expect_tx(bar_full + 8 * s, ((dbg & 8) ? 0u : a_bytes) + kBBytes);
          if (!(dbg & 8))
            tma_load_4d(sA + s * kABytes, &tmA, bar_full + 8 * s, cb * p.kbk, ws + p.tap_dw[tap], hs + p.tap_dh[tap], img);
          tma_load_2d(sB + s * kBBytes, &tmB, bar_full + 8 * s, (int)p.tap_k[tap] * p.Ktap + cb * p.kbk, n0);
          if (++s == STAGES) { s = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1 && !p.halo) {
    // ================================ MMA issuer (warp-uniform loop, one elected lane issues) =================
    const uint32_t idesc = umma_idesc_bf16(BN);
    const bool sw64 = p.kbk == 32;
    // descriptors differ between stages only in their 14-bit start-address field (bytes >> 4, no carry out of it)
    const uint64_t adesc0 = sw64 ? umma_desc_k_sw64(sA) : umma_desc_k_sw128(sA);
    const uint64_t bdesc0 = sw64 ? umma_desc_k_sw64(sB) : umma_desc_k_sw128(sB);
    const uint32_t a_step = stA >> 4, b_step = stB >> 4, a_blk = kABytes >> 4, b_blk = kBBytes >> 4;
    const int ksteps = p.kbk >> 4;
    const int KS = KB / G;
    int s = 0;
    uint32_t phase = 0, it = 0;
    // WRES: the weights are addressed by K block, walked in the producer's rotation (groups never straddle the wrap:
    // G divides KB and the rotation is a multiple of G)
    int kbw = WRES ? (int)(((unsigned)tile0 * 5u) % (unsigned)KS) * G : 0;
    if (WRES && tile0 < ntile) mbar_wait(bar_w, 0);
    for (int t = tile0; t < ntile; t += tstep, it++) {
      const uint32_t buf = it % (uint32_t)p.nacc, aphase = (it / (uint32_t)p.nacc) & 1u;
      mbar_wait(bar_acc_empty + 8 * buf, aphase ^ 1u);       // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * (uint32_t)BN;
      for (int ks = 0; ks < KS; ks++) {
        mbar_wait(bar_full + 8 * s, phase);
        tc_fence_after();
        if (elect_one()) {
          uint64_t ad = adesc0 + (uint64_t)((uint32_t)s * a_step),
                   bd = bdesc0 + (uint64_t)(WRES ? (uint32_t)kbw * b_blk : (uint32_t)s * b_step);
          if (!(dbg & 4)) {
            umma_bf16(d_tmem, ad, bd, idesc, ks ? 1u : 0u);
            umma_bf16(d_tmem, ad + 2, bd + 2, idesc, 1u);
            if (ksteps == 4) {
              umma_bf16(d_tmem, ad + 4, bd + 4, idesc, 1u);
              umma_bf16(d_tmem, ad + 6, bd + 6, idesc, 1u);
            }
#pragma unroll
            for (int g = 1; g < G; g++) {
              ad += a_blk; bd += b_blk;
              umma_bf16(d_tmem, ad, bd, idesc, 1u);
              umma_bf16(d_tmem, ad + 2, bd + 2, idesc, 1u);
              if (ksteps == 4) {
                umma_bf16(d_tmem, ad + 4, bd + 4, idesc, 1u);
                umma_bf16(d_tmem, ad + 6, bd + 6, idesc, 1u);
              }
            }
          }
          if (PAIR) umma_commit_mc(bar_empty + 8 * s, (uint16_t)3);   // one arrival on this slot's barrier in both CTAs
          else umma_commit(bar_empty + 8 * s);     // frees the smem slot once these MMAs have read it
          if (ks == KS - 1) umma_commit(bar_acc_full + 8 * buf);  // accumulator complete
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; phase ^= 1u; }
        if (WRES) { kbw += G; if (kbw == KB) kbw = 0; }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer, halo mode (one lane) =======================================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(BN);
      int s = 0, hslot = 0;
      uint32_t phase = 0, hphase = 0, it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, it++) {
        const uint32_t buf = it % (uint32_t)p.nacc, aphase = (it / (uint32_t)p.nacc) & 1u;
        mbar_wait(bar_acc_empty + 8 * buf, aphase ^ 1u);       // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * (uint32_t)BN;
        if (p.halo) {
          const uint32_t sbo = (uint32_t)(p.TW + 2) * 128u;      // 8-pixel patch rows sit one halo row apart
          for (int cb = 0; cb < p.kb_per_tap; cb++) {
            mbar_wait(bar_afull + 8 * hslot, hphase);
            tc_fence_after();
            const uint32_t ah = sA + hslot * p.a_slot_bytes;
            for (int tap0 = 0; tap0 < p.ntaps; tap0++) {
              int tap = tap0 + (int)(blockIdx.x % 9u);
              if (tap >= p.ntaps) tap -= p.ntaps;
              mbar_wait(bar_full + 8 * s, phase);
              tc_fence_after();
              const uint32_t a0 = ah + (uint32_t)((p.tap_dh[tap] + 1) * (p.TW + 2) + (p.tap_dw[tap] + 1)) * 128u;
              const uint32_t b0 = sB + s * kBBytes;
#pragma unroll
              for (int k = 0; k < kBK / 16; k++) {
                umma_bf16(d_tmem, umma_desc_k_sw128_view(a0 + k * 32, sbo, p.halo == 2),
                          umma_desc_k_sw128(b0 + k * 32), idesc, (cb | tap0 | k) ? 1u : 0u);
              }
              umma_commit(bar_empty + 8 * s);
              if (++s == STAGES) { s = 0; phase ^= 1u; }
            }
            umma_commit(bar_aempty + 8 * hslot);   // halo slot free once all nine taps have read it
            if (++hslot == p.a_slots) { hslot = 0; hphase ^= 1u; }
          }
          umma_commit(bar_acc_full + 8 * buf);
          continue;
        }
        for (int kb = 0; kb < KB; kb++) {
          mbar_wait(bar_full + 8 * s, phase);
          tc_fence_after();
          const uint32_t a0 = sA + s * kABytes, b0 = sB + s * kBBytes;
          if (!(dbg & 4)) {
            if (p.kbk == 32) {
#pragma unroll
              for (int k = 0; k < 2; k++)
                umma_bf16(d_tmem, umma_desc_k_sw64(a0 + k * 32), umma_desc_k_sw64(b0 + k * 32), idesc, (kb | k) ? 1u : 0u);
            } else {
#pragma unroll
              for (int k = 0; k < kBK / 16; k++) {
                umma_bf16(d_tmem, umma_desc_k_sw128(a0 + k * 32), umma_desc_k_sw128(b0 + k * 32), idesc,
                          (kb | k) ? 1u : 0u);
              }
            }
          }
          umma_commit(bar_empty + 8 * s);     // frees the smem slot once these MMAs have read it
          if (++s == STAGES) { s = 0; phase ^= 1u; }
        }
        umma_commit(bar_acc_full + 8 * buf);  // accumulator complete
      }
    }
  } else {
    // ================================ epilogue (4 warps, one TMEM sub-partition each) =========
    const int sub = warp & 3;               // TMEM lanes [32*sub, 32*sub+32) are accessible to this warp
    const int r = sub * 32 + lane;          // accumulator row = pixel of the patch
    const int hl = r / p.TW, wl = r - hl * p.TW;
    const int et = threadIdx.x - 64;        // 0..255
    const uint32_t eg = (uint32_t)(warp - 2) >> 2;   // epilogue group 0|1 owns TMEM accumulator 0|1 (tiles it%2 == eg)
    if (EPI != EPI_RAW) {
      // every tile of this CTA has the same n-tile when n_tiles divides gridDim.x (host guarantees it)
      const int n0c = (tile0 % p.n_tiles) * BN;
      for (int i = et; i < BN; i += 256) {
        const int c = n0c + i;
        s_aff[i] = (p.scale && c < p.Cout) ? p.scale[c] : 1.f;
        s_aff[256 + i] = (p.shift && c < p.Cout) ? p.shift[c] : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    const bool do_stats = (EPI == EPI_RAW) && (p.bn.partial != nullptr) && !(dbg & 2);
    const bool tma = (EPI != EPI_HEAD) && p.epi_tma != 0;
    const bool leader = ((warp - 2) & 3) == 0 && lane == 0;   // issues this epilogue group's TMA stores
    const uint32_t stg = sStage + eg * 16384u;
    float st_s[8], st_q[8];                  // per-lane channel partial sums, one slot per 32-channel chunk
#pragma unroll
    for (int i = 0; i < 8; i++) { st_s[i] = 0.f; st_q[i] = 0.f; }
    float* tw = s_tr + (warp - 2) * (32 * 17);
    uint32_t it = eg;
    for (int t = tile0 + (int)eg * tstep; t < ntile; t += 2 * tstep, it += 2) {
      const int nt = t % p.n_tiles;
      int mt = t / p.n_tiles;
      if (PAIR) mt = 2 * mt + rank;
      const int pw = mt % p.tiles_w; mt /= p.tiles_w;
      const int ph = mt % p.tiles_h;
      const int img = mt / p.tiles_h;
      const int ho = ph * p.TH + hl, wo = pw * p.TW + wl, n0 = nt * BN;
      const bool row_ok = (r < p.TH * p.TW) && (ho < p.Ho) && (wo < p.Wo) && (!PAIR || img < p.N) && !(dbg & 1);
      const long long pix = ((long long)img * p.OutH + ho * p.out_s + p.out_oh) * p.OutW + wo * p.out_s + p.out_ow;
      const uint32_t buf = it % (uint32_t)p.nacc, aphase = (it / (uint32_t)p.nacc) & 1u;   // nacc is even: tiles of one parity
      mbar_wait(bar_acc_full + 8 * buf, aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + buf * (uint32_t)BN + ((uint32_t)(sub * 32) << 16);
      const int nvalid = min(BN, p.Cout - n0);       // channels of this tile that exist
      // head layout: row base (anchor 0) of this lane's pixel, broadcast by shuffle when storing
      const long long head_row = (((long long)img * p.head_na * p.Ho + ho) * p.Wo + wo) * p.head_ch;
      const unsigned ok_mask = __ballot_sync(0xffffffffu, row_ok);
#pragma unroll
      for (int ci = 0; ci < 8; ci++) {
        const int c0 = ci * 32;
        if (c0 < BN) {
          uint32_t v[32];
          tmem_ld32(t_row + (uint32_t)c0, v);
          tmem_ld_wait();
          if (c0 + 32 >= BN) {               // last chunk is in registers: hand the accumulator back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acc_empty + 8 * buf);
          }
          if (c0 < nvalid) {
            if (EPI == EPI_HEAD) {
              // out[((img*na + a)*Ho + ho)*Wo + wo][k], channel c = a*ch + k.  Transpose through smem so that
              // the 32 lanes of a warp write 32 consecutive channels of one pixel (128-byte coalesced stores).
              float* o = (float*)p.out;
#pragma unroll
              for (int h = 0; h < 2; h++) {      // 16 channels per pass: two pixels per store instruction
#pragma unroll
                for (int j = 0; j < 16; j++)
                  tw[lane * 17 + j] = __uint_as_float(v[16 * h + j]) * s_aff[c0 + 16 * h + j] +
                                      s_aff[256 + c0 + 16 * h + j];
                __syncwarp();
                const int cl = c0 + 16 * h + (lane & 15);
                const int c = n0 + cl;
                const int a = c / p.head_ch, k = c - a * p.head_ch;
                const long long coff = (long long)a * p.Ho * p.Wo * p.head_ch + k;
                const bool c_ok = cl < nvalid;
#pragma unroll 4
                for (int q = 0; q < 16; q++) {
                  const int rr = 2 * q + (lane >> 4);
                  const long long rb = __shfl_sync(0xffffffffu, head_row, rr);
                  if (((ok_mask >> rr) & 1u) && c_ok) o[rb + coff] = tw[rr * 17 + (lane & 15)];
                }
                __syncwarp();
              }
            } else {
              __nv_bfloat16* o = (__nv_bfloat16*)p.out + pix * p.out_cpitch + n0 + c0;
              const __nv_bfloat16* res = (EPI == EPI_AFFINE && p.residual)
                                             ? p.residual + pix * p.res_cpitch + n0 + c0 : nullptr;
              const int ng = min(4, (nvalid - c0) >> 3);   // Cout is a multiple of 8
              // epi_tma: the tile leaves through smem, one [128 rows x 64 channels] SWIZZLE_128B slab per TMA store
              // (full 128-byte lines per request instead of 32 scattered 16-byte pieces per warp instruction)
              const bool slab_first = tma && !(ci & 1);
              const bool slab_last = tma && ((ci & 1) || c0 + 32 >= min(BN, nvalid));
              if (slab_first) {
                if (leader) bulk_wait_read0();           // the previous store has finished reading the slab
                group_bar(2u + eg);
              }
#pragma unroll
              for (int g = 0; g < 4; g++) {
                float f[8];
#pragma unroll
                for (int j = 0; j < 8; j++) f[j] = __uint_as_float(v[8 * g + j]);
                if (EPI == EPI_AFFINE) {
                  const float4 s0 = *reinterpret_cast<const float4*>(&s_aff[c0 + 8 * g]);
                  const float4 s1 = *reinterpret_cast<const float4*>(&s_aff[c0 + 8 * g + 4]);
                  const float4 b0 = *reinterpret_cast<const float4*>(&s_aff[256 + c0 + 8 * g]);
                  const float4 b1 = *reinterpret_cast<const float4*>(&s_aff[256 + c0 + 8 * g + 4]);
                  f[0] = act_apply(f[0] * s0.x + b0.x, ACT); f[1] = act_apply(f[1] * s0.y + b0.y, ACT);
                  f[2] = act_apply(f[2] * s0.z + b0.z, ACT); f[3] = act_apply(f[3] * s0.w + b0.w, ACT);
                  f[4] = act_apply(f[4] * s1.x + b1.x, ACT); f[5] = act_apply(f[5] * s1.y + b1.y, ACT);
                  f[6] = act_apply(f[6] * s1.z + b1.z, ACT); f[7] = act_apply(f[7] * s1.w + b1.w, ACT);
                  if (res && row_ok && g < ng) {
                    const uint4 rv = *reinterpret_cast<const uint4*>(res + 8 * g);
                    const __nv_bfloat162* rb = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                      const float2 tt = __bfloat1622float2(rb[j]);
                      f[2 * j] += tt.x;
                      f[2 * j + 1] += tt.y;
                    }
                  }
                }
                uint4 pk;
                __nv_bfloat162* pb = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
                for (int j = 0; j < 4; j++) pb[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                const bool live = row_ok && g < ng;
                if (tma) {
                  // dead rows / channels are stored as zeros: TMA clips them, the statistics read them
                  const uint32_t piece = (uint32_t)(((ci & 1) * 4 + g) ^ (r & 7));
                  sts128(stg + (uint32_t)r * 128u + piece * 16u, live ? pk : make_uint4(0u, 0u, 0u, 0u));
                } else {
                  if (live) *reinterpret_cast<uint4*>(o + 8 * g) = pk;
                  if (do_stats) {              // statistics of the values as stored (bf16-rounded, packed pairs)
                    uint32_t* tw32 = reinterpret_cast<uint32_t*>(tw) + lane * 17 + 4 * g;
                    tw32[0] = live ? pk.x : 0u;
                    tw32[1] = live ? pk.y : 0u;
                    tw32[2] = live ? pk.z : 0u;
                    tw32[3] = live ? pk.w : 0u;
                  }
                }
              }
              if (do_stats) {
                __syncwarp();
                float a = 0.f, b = 0.f;
                if (tma) {
                  // channel c0 + lane of this warp's own 32 rows, read back from the slab
                  const uint32_t cpiece = (uint32_t)((ci & 1) * 4 + (lane >> 3));
                  const uint32_t cbyte = (uint32_t)(lane & 7) * 2u;
#pragma unroll 8
                  for (int rr = 0; rr < 32; rr++) {
                    const uint32_t R = (uint32_t)(sub * 32 + rr);
                    const float x = __uint_as_float(lds_u16(stg + R * 128u + ((cpiece ^ (R & 7u)) * 16u) + cbyte) << 16);
                    a += x;
                    b += x * x;
                  }
                } else {
#pragma unroll 8
                  for (int rr = 0; rr < 32; rr++) {
                    const uint32_t w2 = reinterpret_cast<const uint32_t*>(tw)[rr * 17 + (lane >> 1)];
                    const float x = __uint_as_float((lane & 1) ? (w2 & 0xffff0000u) : (w2 << 16));
                    a += x;
                    b += x * x;
                  }
                }
                st_s[ci] += a;
                st_q[ci] += b;
                __syncwarp();
              }
              if (slab_last) {
                fence_proxy_async();
                group_bar(2u + eg);
                if (leader && !(dbg & 1)) {
                  const int cc = n0 + (ci >> 1) * 64;
                  if (p.epi_tma == 2) tma_reduce_add_4d(&tmO, stg, cc, pw * p.TW, ph * p.TH, img);
                  else tma_store_4d(&tmO, stg, cc, pw * p.TW, ph * p.TH, img);
                  bulk_commit();
                }
              }
            }
          }
        }
      }
    }
    if (do_stats && !(dbg & 16)) {     // dbg 16 (timing experiment, wrong results): per-tile statistics only, no cross-CTA tail
      // Deterministic reduction (the reference runs with cudnn.deterministic, train.py:24): the 8 epilogue warps
      // combine through smem in a fixed order; across CTAs the sums are fixed-point (see below); the last CTA to
      // arrive turns the totals into scale/shift + running statistics.
      const int n0c = (tile0 % p.n_tiles) * BN;
      float* comb = s_tr;                                  // reused as [8 warps][2][256]
      asm volatile("bar.sync 1, 256;" ::: "memory");       // every warp is done with its transpose tile
#pragma unroll
      for (int ci = 0; ci < 8; ci++) {
        if (ci * 32 < BN) {
          comb[((warp - 2) * 2 + 0) * 256 + ci * 32 + lane] = st_s[ci];
          comb[((warp - 2) * 2 + 1) * 256 + ci * 32 + lane] = st_q[ci];
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      // Every CTA adds its per-channel partials to ONE pair of 64-bit fixed-point accumulators per channel: integer
      // addition is associative, so the totals do not depend on the order the CTAs arrive in.
      unsigned long long* acc = reinterpret_cast<unsigned long long*>(p.bn.partial);   // [2][Cout], zero between launches
      for (int i = et; i < 2 * BN; i += 256) {
        const int q = i / BN, cl = i - q * BN, c = n0c + cl;
        if (c < p.Cout) {
          float a = 0.f;
#pragma unroll
          for (int w = 0; w < 8; w++) a += comb[(w * 2 + q) * 256 + cl];
          atomicAdd(acc + (size_t)q * p.Cout + c, (unsigned long long)__float2ll_rn(a * (q ? kSqScale : kSumScale)));
        }
      }
      // counter == nullptr: deferred finalize.  The consumer (ryolo_scale_shift_act_bn) turns the totals into scale / shift
      // itself, so this CTA is done: no fence, no arrival counter, no last-CTA pass (0.8 ms per yolov4 forward).
      if (p.bn.counter != nullptr) {
      __threadfence();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (et == 0) s_last = (atomicAdd(p.bn.counter, 1u) == gridDim.x - 1) ? 1 : 0;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (s_last) {
        __threadfence();
        if (et == 0 && p.bn.num_batches) *p.bn.num_batches += 1;
        for (int c = et; c < p.Cout; c += 256) {
          const long long is = (long long)__ldcg(acc + c), iq = (long long)__ldcg(acc + p.Cout + c);
          acc[c] = 0ull;                                   // leave the scratch zeroed for the next launch
          acc[p.Cout + c] = 0ull;
          const double ds = (double)is * (1.0 / kSumScale), dq = (double)iq * (1.0 / kSqScale);
          const float fs = (float)ds, fq = (float)dq;
          if (p.bn.sum) p.bn.sum[c] = fs;
          if (p.bn.sumsq) p.bn.sumsq[c] = fq;
          const double mean = ds / p.bn_count;
          double var = dq / p.bn_count - mean * mean;
          if (var < 0.0) var = 0.0;
          const float invstd = (float)(1.0 / sqrt(var + (double)p.bn.eps));
          const float sc = p.bn.gamma[c] * invstd;
          p.bn.scale[c] = sc;
          p.bn.shift[c] = p.bn.beta[c] - (float)mean * sc;
          if (p.bn.save_mean) p.bn.save_mean[c] = (float)mean;
          if (p.bn.save_invstd) p.bn.save_invstd[c] = invstd;
          if (p.bn.running_mean) {
            const double unbiased = p.bn_count > 1.0 ? var * p.bn_count / (p.bn_count - 1.0) : var;
            p.bn.running_mean[c] = (1.f - p.bn.momentum) * p.bn.running_mean[c] + p.bn.momentum * (float)mean;
            p.bn.running_var[c] = (1.f - p.bn.momentum) * p.bn.running_var[c] + p.bn.momentum * (float)unbiased;
          }
        }
      }
      }
    }
  }
  if (p.epi_tma && warp >= 2 && ((warp - 2) & 3) == 0 && lane == 0) bulk_wait0();   // stores complete before exit
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();                 // no CTA leaves while its peer may still arrive on its barriers
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// ------------------------------------------------------------------------------------ reference kernel
// Plain CUDA-core direct convolution on the same bf16 operands (fp32 accumulate).  Not a fallback: it is
// only reachable through ryolo_conv2d_reference and exists so that the tcgen05 path can be checked on the
// device, layer by layer, independently of any library.
__global__ void conv_ref_kernel(const __nv_bfloat16* __restrict__ x, long long x_cpitch, int N, int H, int W, int Cin,
                                const __nv_bfloat16* __restrict__ w, ConvKernelParams p) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long total = (long long)p.N * p.Ho * p.Wo * p.Cout;
  if (idx >= total) return;
  const int c = (int)(idx % p.Cout);
  long long pix = idx / p.Cout;
  const int wo = (int)(pix % p.Wo);
  const int ho = (int)((pix / p.Wo) % p.Ho);
  const int img = (int)(pix / ((long long)p.Wo * p.Ho));
  float acc = 0.f;
  for (int kh = 0; kh < p.ksize; kh++) {
    const int hi = ho * p.stride + kh - p.pad;
    if (hi < 0 || hi >= H) continue;
    for (int kw = 0; kw < p.ksize; kw++) {
      const int wi = wo * p.stride + kw - p.pad;
      if (wi < 0 || wi >= W) continue;
      const __nv_bfloat16* xp = x + (((long long)img * H + hi) * W + wi) * x_cpitch;
      const __nv_bfloat16* wp = w + ((long long)c * p.ksize * p.ksize + kh * p.ksize + kw) * Cin;
      for (int ci = 0; ci < Cin; ci++) acc += __bfloat162float(xp[ci]) * __bfloat162float(wp[ci]);
    }
  }
  float v = acc;
  if (p.scale) v = v * p.scale[c];
  if (p.shift) v = v + p.shift[c];
  v = act_apply(v, p.act);
  if (p.mode == RYOLO_OUT_NHWC_BF16) {
    if (p.residual) v += __bfloat162float(p.residual[pix * p.res_cpitch + c]);
    ((__nv_bfloat16*)p.out)[pix * p.out_cpitch + c] = __float2bfloat16_rn(v);
  } else {
    const int a = c / p.head_ch, k = c - a * p.head_ch;
    ((float*)p.out)[((((long long)img * p.head_na + a) * p.Ho + ho) * p.Wo + wo) * p.head_ch + k] = v;
  }
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// Patch shape (TH x TW <= 128) that wastes the fewest accumulator rows for an Ho x Wo output.
void pick_patch(int Ho, int Wo, int stride, int* TH, int* TW) {
  double best = -1.0;
  const int tw_max = 128;
  for (int tw = 1; tw <= tw_max && tw <= Wo; tw++) {
    int th = 128 / tw;
    if (th > Ho) th = Ho;
    if (th * stride > 256) th = 256 / stride;
    const double tiles = (double)((Ho + th - 1) / th) * ((Wo + tw - 1) / tw);
    const double eff = (double)Ho * Wo / (tiles * 128.0);
    // prefer wider rows on ties (longer contiguous runs for TMA)
    if (eff > best + 1e-9 || (eff > best - 1e-9 && tw > *TW)) { best = eff; *TH = th; *TW = tw; }
  }
}

// 3x3 / stride-1 tap sets can read nine shifted views of one halo box (TW = 8 so that every 8-row UMMA group is one
// patch row and the groups sit a constant (TW+2)*128 bytes apart).  Used when the 16x8 patch grid wastes little.
// RYOLO_HALO=0 disables it, 2 sets the descriptor's base_offset field (bring-up switch).
void maybe_enable_halo(ConvKernelParams* p) {
  const int mode = ryolo_knob(RYOLO_KNOB_HALO);
  p->halo = 0;
  if (!mode || p->ntaps != 9 || p->stride != 1 || p->ksize != 3 || p->kbk != kBK) return;
  const double tiles = (double)((p->Ho + 15) / 16) * ((p->Wo + 7) / 8);
  const double eff = (double)p->Ho * p->Wo / (tiles * 128.0);
  const double cur = (double)p->Ho * p->Wo / ((double)p->tiles_h * p->tiles_w * 128.0);
  if (eff < 0.85 * cur) return;
  p->halo = mode;
  p->TH = 16; p->TW = 8;
  p->tiles_h = (p->Ho + 15) / 16;
  p->tiles_w = (p->Wo + 7) / 8;
}

int fill_params(const ryolo_conv_desc* d, ConvKernelParams* p) {
  RY_CHECK_ARG(d->ksize == 1 || d->ksize == 3, "conv: ksize must be 1 or 3");
  RY_CHECK_ARG(d->stride == 1 || d->stride == 2, "conv: stride must be 1 or 2");
  RY_CHECK_ARG(d->Cin % 8 == 0 && d->x_cpitch % 8 == 0, "conv: Cin and the input channel pitch must be multiples of 8");
  RY_CHECK_ARG((((uintptr_t)d->x) & 15) == 0 && (((uintptr_t)d->w) & 15) == 0, "conv: operands must be 16-byte aligned");
  p->N = d->N; p->Cin = d->Cin; p->Cout = d->Cout;
  p->ksize = d->ksize; p->stride = d->stride; p->pad = (d->ksize - 1) / 2;
  p->Ho = (d->H + 2 * p->pad - d->ksize) / d->stride + 1;
  p->Wo = (d->W + 2 * p->pad - d->ksize) / d->stride + 1;
  {   // knob sw64: 1 = 32-wide K blocks for Cin == 32 only; 2 = also for every layer with Cout <= 128, 3 = for every layer
      // (experiment: half-size stages, twice the ring depth)
    const int sw = ryolo_knob(RYOLO_KNOB_SW64);
    p->kbk = ((d->Cin == 32 && sw) || (sw == 2 && d->Cout <= 128 && d->Cin % 32 == 0) || (sw == 3 && d->Cin % 32 == 0)) ? 32 : kBK;
  }
  p->kb_per_tap = (d->Cin + p->kbk - 1) / p->kbk;
  p->ntaps = d->ksize * d->ksize; p->Ktap = d->Cin;
  for (int t = 0; t < p->ntaps; t++) {
    p->tap_dh[t] = (signed char)(t / d->ksize - p->pad);
    p->tap_dw[t] = (signed char)(t % d->ksize - p->pad);
    p->tap_k[t] = (unsigned char)t;
  }
  p->out_s = 1; p->out_oh = 0; p->out_ow = 0; p->OutH = p->Ho; p->OutW = p->Wo;
  p->mode = d->out_mode; p->act = d->act;
  p->out = d->out; p->out_cpitch = d->out_cpitch;
  p->scale = d->scale; p->shift = d->shift;
  p->residual = (const __nv_bfloat16*)d->residual; p->res_cpitch = d->res_cpitch;
  p->head_na = d->head_na; p->head_ch = d->head_ch;
  if (d->bn) {
    RY_CHECK_ARG(d->out_mode == RYOLO_OUT_NHWC_BF16 && !d->scale && !d->shift && d->act == RYOLO_ACT_LINEAR &&
                     !d->residual, "conv: fused BatchNorm statistics need the raw (no scale/shift/act/residual) epilogue");
    RY_CHECK_ARG(d->bn->partial && (!d->bn->counter || (d->bn->gamma && d->bn->beta && d->bn->scale && d->bn->shift)),
                 "conv: incomplete ryolo_bn_fuse");
    p->bn = *d->bn;
  }
  if (d->out_mode == RYOLO_OUT_NHWC_BF16) {
    RY_CHECK_ARG(d->Cout % 8 == 0 && d->out_cpitch % 8 == 0 && (((uintptr_t)d->out) & 15) == 0,
                 "conv: bf16 output needs Cout and channel pitch multiples of 8 and a 16-byte aligned pointer");
    RY_CHECK_ARG(!d->residual || (d->res_cpitch % 8 == 0 && (((uintptr_t)d->residual) & 15) == 0),
                 "conv: residual must be 16-byte aligned with a channel pitch multiple of 8");
  } else {
    RY_CHECK_ARG(d->out_mode == RYOLO_OUT_HEAD_F32 && d->head_na * d->head_ch == d->Cout && !d->residual,
                 "conv: head output needs na*ch == Cout and no residual");
  }
  p->TH = 1; p->TW = 1;
  pick_patch(p->Ho, p->Wo, d->stride, &p->TH, &p->TW);
  p->tiles_h = (p->Ho + p->TH - 1) / p->TH;
  p->tiles_w = (p->Wo + p->TW - 1) / p->TW;
  p->bn_count = (double)p->N * p->Ho * p->Wo;
  maybe_enable_halo(p);
  return RYOLO_OK;
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

constexpr size_t kSmemBudget = 204 * 1024;   // dynamic; + ~20 KB static (transpose tile, scale/shift, barriers)

int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, ConvKernelParams& p, cudaStream_t st) {
  p.dbg = ryolo_knob(RYOLO_KNOB_DBG);
  p.n_tiles = (p.Cout + p.BN - 1) / p.BN;
  // TMA-store epilogue (knob epi_tma: 1 store, 2 also turns dgrad's read-add-write accumulation into a TMA reduce-add).
  // A slab is 64 channels wide, so a tile whose width is not a multiple of 64 may only be the last one of its row.
  CUtensorMap tmO = tmA;
  p.epi_tma = 0;
  const int knob = ryolo_knob(RYOLO_KNOB_EPI_TMA);
  // Measured per layer (profiles/r01_diag_knobs_bs32.txt): the slab costs 32 KB of the operand ring, which only the
  // deep-K BN=256 tiles miss; everything else gains, and dgrad's accumulation always does (no read-add-write).
  const bool reducible = knob == 2 && p.residual == (const __nv_bfloat16*)p.out && p.res_cpitch == p.out_cpitch &&
                         !p.scale && !p.shift && p.act == RYOLO_ACT_LINEAR;
  const bool fits = p.BN <= ryolo_knob(RYOLO_KNOB_EPI_MAXBN) || p.ntaps * p.kb_per_tap * p.kbk <= 1152 || reducible;
  if (knob && p.mode == RYOLO_OUT_NHWC_BF16 && fits && (p.BN % 64 == 0 || p.n_tiles == 1)) {
    p.epi_tma = 1;
    if (reducible) {
      p.epi_tma = 2;
      p.residual = nullptr;
    }
    const long long pitch = p.out_cpitch;
    __nv_bfloat16* base = (__nv_bfloat16*)p.out + ((long long)p.out_oh * p.OutW + p.out_ow) * pitch;
    cuuint64_t dims[4] = {(cuuint64_t)p.Cout, (cuuint64_t)p.Wo, (cuuint64_t)p.Ho, (cuuint64_t)p.N};
    cuuint64_t strides[3] = {(cuuint64_t)p.out_s * pitch * 2, (cuuint64_t)p.out_s * p.OutW * pitch * 2,
                             (cuuint64_t)p.OutH * p.OutW * pitch * 2};
    cuuint32_t box[4] = {(cuuint32_t)kBK, (cuuint32_t)p.TW, (cuuint32_t)p.TH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = get_encode()(&tmO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)base, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { ryolo_set_error("cuTensorMapEncodeTiled failed for the output operand"); return RYOLO_ERR_CUDA; }
  }
  size_t stage_bytes = (size_t)kBM * p.kbk * 2 + (size_t)p.BN * p.kbk * 2, fixed = 0;
  if (p.halo) {
    p.a_slots = 3;
    p.a_slot_bytes = (uint32_t)ry_align_up((size_t)(p.TH + 2) * (p.TW + 2) * kBK * 2, 1024);
    fixed = (size_t)p.a_slots * p.a_slot_bytes;
    stage_bytes = (size_t)p.BN * kBK * 2;
  }
  // Resident weights (knob wres = largest weight matrix in KB that stays in shared memory, default 96; 0 = off; + 1024 =
  // also layers with few tiles per CTA): a layer with ONE n-tile whose K blocks all fit loads them once per CTA; the ring
  // then carries activation boxes only, which halves the TMA issues of the narrow layers (their 2-8 KB weight boxes cost
  // a full issue + barrier transaction each).  Measured per layer (profiles/r02_wres_ab.txt): stem 0.83 -> 0.68 ms,
  // 3x3 s2 32->64 @400x400 0.61 -> 0.46, the 1x1 32/64-channel layers on 400x400 maps -15..-28 %, forward and dgrad alike.
  // Layers with ~17 tiles per CTA (100x100 maps) lose up to 10 %: the first MMA waits for the whole matrix, so the
  // mode needs >= 32 tiles per CTA to pay for its start.
  p.wres = 0;
  {
    const size_t wbytes = (size_t)p.ntaps * p.kb_per_tap * p.BN * p.kbk * 2;
    const int kw = ryolo_knob(RYOLO_KNOB_WRES);
    const long long mtiles = (long long)p.N * p.tiles_h * p.tiles_w;
    if ((kw & 1023) > 0 && !p.halo && !p.pair && !p.dbg && p.n_tiles == 1 && wbytes <= (size_t)(kw & 1023) * 1024 &&
        ((kw & 1024) || mtiles >= 32ll * sm_count())) {
      p.wres = 1;
      fixed = wbytes;
      stage_bytes = (size_t)kBM * p.kbk * 2;
    }
  }
  const size_t slab_bytes = p.epi_tma ? 2 * 16384 : 0;      // one slab per epilogue group, after the ring
  // K blocks per stage (knob kgrp: 0 = always one; 1 (default) = N <= 32 tiles only, stages of up to 40 KB with a ring of
  // >= 3; 2 = any tile, up to 64 KB with a ring of >= 2; 3 = up to 72 KB): the largest divisor (<= 3) of the tile's K-block
  // count that fits.  Measured per layer (gpurun_out/r4b_diag_knobs_bs32.txt -> profiles/r02_kgrp_ab.txt): the Cin = 32
  // 3x3 layers on 400x400 maps (10 KB per block, two N = 32 MMAs) spend their time on the per-stage barrier hand-shake:
  // three taps per stage take them from 0.47 to 0.35 ms (forward and dgrad alike), the stride-2 dgrad classes of the
  // 32 -> 64 layer gain 9 %.  Wider tiles lose: a 64 KB stage leaves N = 128 layers a ring of two (+12 %).
  p.kgrp = 1;
  if (!p.halo && !p.dbg && !p.pair) {
    const int knob_g = ryolo_knob(RYOLO_KNOB_KGRP);
    const size_t limit = knob_g == 1 ? 40 * 1024 : knob_g == 2 ? 64 * 1024 : knob_g >= 3 ? 72 * 1024 : 0;
    const size_t min_ring = knob_g == 1 ? 3 : 2;
    const int KB = p.ntaps * p.kb_per_tap;
    for (int g = 2; g <= 3; g++) {
      if (KB % g) continue;
      if (knob_g == 1 && p.BN > 32) continue;
      if (g * stage_bytes <= limit && (kSmemBudget - 1024 - fixed - slab_bytes) / (g * stage_bytes) >= min_ring) p.kgrp = g;
    }
    stage_bytes *= (size_t)p.kgrp;
  }
  int stages = (int)((kSmemBudget - 1024 - fixed - slab_bytes) / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  p.stages = stages;
  const size_t smem = fixed + (size_t)stages * stage_bytes + slab_bytes + 1024;
  // Accumulators in rotation: with two, a tile whose main loop is one or two K blocks long (1x1 convs on wide maps) waits
  // for the commit -> epilogue wake-up -> tcgen05.ld -> arrive -> MMA wake-up round trip (~1 us) every other tile
  // (ncu: the MMA warp's retries sit on acc_empty, not on the smem ring).  Narrow tiles fit up to 8 in the 512 columns.
  int nacc = 512 / p.BN;
  nacc = nacc > kMaxAcc ? kMaxAcc : nacc & ~1;
  if (nacc < 2 || !ryolo_knob(RYOLO_KNOB_NACC)) nacc = 2;
  p.nacc = nacc;
  uint32_t cols = 32;
  while (cols < (uint32_t)nacc * (uint32_t)p.BN) cols <<= 1;
  p.tmem_cols = cols;
  typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const ConvKernelParams);
  const bool affine = p.scale || p.shift || p.act != RYOLO_ACT_LINEAR || p.residual;
  int which;                      // 0 head | 1 raw | 2 leaky | 3 mish | 4 swish | 5 linear-affine
  if (p.mode == RYOLO_OUT_HEAD_F32) which = 0;
  else if (!affine) which = 1;
  else if (p.act == RYOLO_ACT_LEAKY) which = 2;
  else if (p.act == RYOLO_ACT_MISH) which = 3;
  else if (p.act == RYOLO_ACT_SWISH) which = 4;
  else which = 5;
  // [epilogue kind][K blocks per stage - 1 | 3 = the timing-experiment build (one block per stage) | 4 = cluster pairs |
  //                  5..7 = resident weights with 1..3 K blocks per stage]
#define RY_CONV_ROW(E, A)                                                                                                  \
  { conv_fwd_kernel<E, A, false, 1, false, false>, conv_fwd_kernel<E, A, false, 2, false, false>,                          \
    conv_fwd_kernel<E, A, false, 3, false, false>, conv_fwd_kernel<E, A, true, 1, false, false>,                           \
    conv_fwd_kernel<E, A, false, 1, true, false>, conv_fwd_kernel<E, A, false, 1, false, true>,                            \
    conv_fwd_kernel<E, A, false, 2, false, true>, conv_fwd_kernel<E, A, false, 3, false, true> }
  static const KernelFn all[6][8] = {RY_CONV_ROW(EPI_HEAD, 0), RY_CONV_ROW(EPI_RAW, 0), RY_CONV_ROW(EPI_AFFINE, RYOLO_ACT_LEAKY),
                                     RY_CONV_ROW(EPI_AFFINE, RYOLO_ACT_MISH), RY_CONV_ROW(EPI_AFFINE, RYOLO_ACT_SWISH),
                                     RY_CONV_ROW(EPI_AFFINE, RYOLO_ACT_LINEAR)};
#undef RY_CONV_ROW
  const KernelFn fn = all[which][p.pair ? 4 : p.dbg ? 3 : p.wres ? 4 + p.kgrp : p.kgrp - 1];
  static bool configured = false;
  if (!configured) {
    for (int i = 0; i < 48; i++) {
      const KernelFn k = all[i / 8][i % 8];
      cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget);
      if (e != cudaSuccess) { ryolo_set_error(cudaGetErrorString(e)); return RYOLO_ERR_CUDA; }
    }
    configured = true;
  }
  const long long tiles = (long long)p.N * p.tiles_h * p.tiles_w * p.n_tiles;
  RY_CHECK_ARG(tiles > 0 && tiles < (1ll << 31), "conv: tile count out of range");
  RY_CHECK_ARG(p.n_tiles <= sm_count(), "conv: too many output-channel tiles");
  if (p.pair) {
    // clusters of two: as many pairs as can be resident at once (a multiple of n_tiles, so that a pair keeps its n-tile)
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    static int max_pairs[6] = {0, 0, 0, 0, 0, 0};
    if (!max_pairs[which]) {
      cfg.gridDim = dim3(sm_count() & ~1);
      int n = 0;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&n, fn, &cfg);
      if (e != cudaSuccess || n <= 0) { cudaGetLastError(); n = sm_count() / 2 - 4; }
      max_pairs[which] = n < sm_count() / 2 ? n : sm_count() / 2;
    }
    const long long pair_items = (long long)p.n_tiles * (((long long)p.N * p.tiles_h * p.tiles_w + 1) / 2);
    int pairs = max_pairs[which];
    pairs -= pairs % p.n_tiles;
    if (pair_items < pairs) pairs = (int)pair_items;
    RY_CHECK_ARG(pairs > 0, "conv: no room for a cluster pair");
    cfg.gridDim = dim3(2 * pairs);
    cfg.numAttrs = ryolo_knob(RYOLO_KNOB_PDL) ? 2 : 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, fn, tmA, tmB, tmO, p);
    if (le != cudaSuccess) { ryolo_set_error(cudaGetErrorString(le)); return RYOLO_ERR_CUDA; }
    RY_CHECK_LAUNCH();
    return RYOLO_OK;
  }
  int grid = sm_count();
  grid -= grid % p.n_tiles;                          // a CTA always sees the same n-tile
  if (tiles < grid) grid = (int)tiles;               // tiles is a multiple of n_tiles, so this keeps the property
  cudaError_t le = ry_launch(fn, dim3(grid), dim3(kThreads), smem, st, tmA, tmB, tmO, p);
  if (le != cudaSuccess) { ryolo_set_error(cudaGetErrorString(le)); return RYOLO_ERR_CUDA; }
  RY_CHECK_LAUNCH();
  return RYOLO_OK;
}

// Output-channel tile: multiple of 32 (16 for tiny Cout) up to 256, minimising padded columns.
int pick_bn(int Cout) {
  if (Cout <= 16) return 16;
  if (Cout <= 256) return (Cout + 31) / 32 * 32;
  int best = 256, best_waste = 1 << 30;
  for (int bn = 256; bn >= 128; bn -= 32) {
    const int waste = (Cout + bn - 1) / bn * bn - Cout;
    if (waste < best_waste) { best_waste = waste; best = bn; }
  }
  return best;
}

// Encodes the activation (4-D, box = one patch, element stride = conv stride) and weight (2-D) tensor maps and launches.
int encode_and_launch(const void* x, int N, int H, int W, int C, long long cpitch, int estride, const void* w,
                      int Ktot, int wrows, ConvKernelParams& p, cudaStream_t st) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { ryolo_set_error("cuTensorMapEncodeTiled not available from the driver"); return RYOLO_ERR_CUDA; }
  CUtensorMap tmA, tmB;
  {
    // knob pair: 0 = off, otherwise the narrowest output-channel tile that runs as cluster pairs (weight tile multicast;
    // default 256).  Deep N = 256 layers are bound by L2 -> SM operand traffic (48 KB per K block; without operand loads
    // they run at the burst MMA rate, profiles/r02_knockout_uniform.txt): sharing the weight tile takes the K >= 2304
    // layers down 3-8 % (profiles/r02_pair_ab.txt).  Shorter K loops lose (the two CTAs release every ring slot
    // together, and with the store slab the ring is three deep), so they stay unpaired.
    const int kp = ryolo_knob(RYOLO_KNOB_PAIR);              // + 1024: also layers too small to gain (parity tests)
    const int min_bn = kp & 1023;
    const bool any_size = (kp & 1024) != 0;
    const long long mtiles = (long long)p.N * p.tiles_h * p.tiles_w;
    p.pair = (min_bn > 0 && !p.halo && !ryolo_knob(RYOLO_KNOB_DBG) && p.BN >= min_bn && p.BN % 16 == 0 && p.kbk == kBK &&
              mtiles >= 2 && (any_size || (mtiles * ((p.Cout + p.BN - 1) / p.BN) >= 2 * (long long)sm_count() && p.ntaps * p.kb_per_tap >= 36))) ? 1 : 0;
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)cpitch * 2, (cuuint64_t)cpitch * 2 * W, (cuuint64_t)cpitch * 2 * W * H};
    cuuint32_t box[4] = {(cuuint32_t)p.kbk, (cuuint32_t)(p.halo ? p.TW + 2 : p.TW * estride),
                         (cuuint32_t)(p.halo ? p.TH + 2 : p.TH * estride), 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)x, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, p.kbk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { ryolo_set_error("cuTensorMapEncodeTiled failed for the activation operand"); return RYOLO_ERR_CUDA; }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)wrows};
    cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
    cuuint32_t box[2] = {(cuuint32_t)p.kbk, (cuuint32_t)(p.pair ? p.BN / 2 : p.BN)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)w, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, p.kbk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { ryolo_set_error("cuTensorMapEncodeTiled failed for the weight operand"); return RYOLO_ERR_CUDA; }
  }
  return launch(tmA, tmB, p, st);
}

}  // namespace

extern "C" {

// Conv2d(k in {1,3}, stride in {1,2}, pad=(k-1)/2, no bias) on NHWC bf16 + fused epilogue; see conv_desc.h.
int ryolo_conv2d_forward(const ryolo_conv_desc* d, void* stream) {
  ConvKernelParams p{};
  int rc = fill_params(d, &p);
  if (rc) return rc;
  if (d->N == 0) return RYOLO_OK;
  p.BN = pick_bn(d->Cout);
  return encode_and_launch(d->x, d->N, d->H, d->W, d->Cin, d->x_cpitch, d->stride, d->w,
                           d->ksize * d->ksize * d->Cin, d->Cout, p, (cudaStream_t)stream);
}

// Gradient wrt the input of Conv2d(Cin -> Cout, ksize, stride, pad=(ksize-1)/2):  dx[N,H,W,Cin] (+)= conv^T(dy, w).
//   dy  bf16 NHWC view [N,Ho,Wo,Cout];  wt  bf16 [Cin][kh][kw][Cout] (ryolo_pack_weights with transpose=1)
// Runs on the same tcgen05 kernel: the taps are mirrored, and a stride-s conv splits into s*s output-parity
// classes, each an ordinary stride-1 implicit GEMM over dy writing an s-strided lattice of dx.
int ryolo_conv2d_dgrad(const void* dy, long long dy_cpitch, int N, int H, int W, int Cin, int Cout, int ksize,
                       int stride, const void* wt, void* dx, long long dx_cpitch, int accumulate, void* stream) {
  RY_CHECK_ARG(ksize == 1 || ksize == 3, "dgrad: ksize must be 1 or 3");
  RY_CHECK_ARG(stride == 1 || stride == 2, "dgrad: stride must be 1 or 2");
  RY_CHECK_ARG(Cin % 8 == 0 && Cout % 8 == 0 && dy_cpitch % 8 == 0 && dx_cpitch % 8 == 0,
               "dgrad: channel counts and pitches must be multiples of 8");
  RY_CHECK_ARG((((uintptr_t)dy) & 15) == 0 && (((uintptr_t)dx) & 15) == 0 && (((uintptr_t)wt) & 15) == 0,
               "dgrad: operands must be 16-byte aligned");
  if (N == 0) return RYOLO_OK;
  const int pad = (ksize - 1) / 2;
  const int Ho = (H + 2 * pad - ksize) / stride + 1, Wo = (W + 2 * pad - ksize) / stride + 1;
  for (int ph = 0; ph < stride; ph++) {
    for (int pw = 0; pw < stride; pw++) {
      const int Hl = (H - ph + stride - 1) / stride, Wl = (W - pw + stride - 1) / stride;
      if (Hl <= 0 || Wl <= 0) continue;
      ConvKernelParams p{};
      p.N = N; p.Ho = Hl; p.Wo = Wl; p.Cout = Cin; p.Cin = Cout; p.Ktap = Cout;
      p.ksize = ksize; p.stride = 1; p.pad = pad;
      {
        const int sw = ryolo_knob(RYOLO_KNOB_SW64);
        p.kbk = ((Cout == 32 && sw) || (sw == 2 && Cin <= 128 && Cout % 32 == 0) || (sw == 3 && Cout % 32 == 0)) ? 32 : kBK;
      }
      p.kb_per_tap = (Cout + p.kbk - 1) / p.kbk;
      p.ntaps = 0;
      for (int kh = 0; kh < ksize; kh++) {
        const int vh = ph + pad - kh;
        if (((vh % stride) + stride) % stride) continue;
        for (int kw = 0; kw < ksize; kw++) {
          const int vw = pw + pad - kw;
          if (((vw % stride) + stride) % stride) continue;
          p.tap_dh[p.ntaps] = (signed char)(vh >= 0 ? vh / stride : -((-vh) / stride));
          p.tap_dw[p.ntaps] = (signed char)(vw >= 0 ? vw / stride : -((-vw) / stride));
          p.tap_k[p.ntaps] = (unsigned char)(kh * ksize + kw);
          p.ntaps++;
        }
      }
      p.mode = RYOLO_OUT_NHWC_BF16; p.act = RYOLO_ACT_LINEAR;
      p.out = dx; p.out_cpitch = dx_cpitch;
      p.out_s = stride; p.out_oh = ph; p.out_ow = pw; p.OutH = H; p.OutW = W;
      if (accumulate) { p.residual = (const __nv_bfloat16*)dx; p.res_cpitch = dx_cpitch; }
      p.TH = 1; p.TW = 1;
      pick_patch(Hl, Wl, 1, &p.TH, &p.TW);
      p.tiles_h = (Hl + p.TH - 1) / p.TH;
      p.tiles_w = (Wl + p.TW - 1) / p.TW;
      if (stride == 1) maybe_enable_halo(&p);
      p.BN = pick_bn(Cin);
      if (p.ntaps == 0) {      // no contribution for this parity class (cannot happen for k=3/s=2, k=1/s=1)
        ryolo_set_error("dgrad: empty tap set");
        return RYOLO_ERR_INVALID;
      }
      int rc = encode_and_launch(dy, N, Ho, Wo, Cout, dy_cpitch, 1, wt, ksize * ksize * Cout, Cin, p,
                                 (cudaStream_t)stream);
      if (rc) return rc;
    }
  }
  return RYOLO_OK;
}

// Same contract as ryolo_conv2d_forward, computed by a plain CUDA-core kernel (device-side checker).
int ryolo_conv2d_reference(const ryolo_conv_desc* d, void* stream) {
  ConvKernelParams p{};
  int rc = fill_params(d, &p);
  if (rc) return rc;
  if (d->N == 0) return RYOLO_OK;
  const long long total = (long long)p.N * p.Ho * p.Wo * p.Cout;
  conv_ref_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)d->x, d->x_cpitch, d->N, d->H, d->W, d->Cin, (const __nv_bfloat16*)d->w, p);
  RY_CHECK_LAUNCH();
  return RYOLO_OK;
}

}  // extern "C"
