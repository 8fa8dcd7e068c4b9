// Weight gradient of Conv2d on the tcgen05 tensor cores (K3 in SURVEY.md §2.1): the conv_backward-weight half of
// autograd's backward at train.py:198.
//
//   dW[co, tap, ci] = sum over pixels p of  dY[p, co] * X[p shifted by tap, ci]
//
// GEMM view: M = 128 output channels, N = 64*nb input channels, K = pixels.  Both operands are NHWC bf16, i.e.
// "MN-major" for this GEMM (the channel dimension is contiguous, the reduction dimension strides by one pixel),
// which UMMA consumes directly through MN-major SWIZZLE_128B descriptors: a [pixels x 64 channels] TMA box is
// exactly one canonical MN-major atom column.  The K block is one TH x TW patch (<= 128 pixels); rows past
// TH*TW are zeroed once (TMA never touches them), out-of-image rows are zero-filled by TMA, and the shifted X box
// of every tap reuses the same tensor map as the forward pass (element strides = conv stride, OOB = padding).
//
//   work item   (128-channel Cout tile) x (64*nb-channel Cin tile) x (group of taps that fits TMEM: tg*64*nb <= 512)
//   split-K     the patches of an item are dealt round-robin to its CTAs (the SMs are shared out between the items by
//               their measured cost per patch); every CTA keeps its partial dW in TMEM for its whole life and adds it
//               to the K-major fp32 scratch with vector atomics once at the end
//   MMAs        narrow Cin tiles issue several taps per MMA (consecutive ring slots = consecutive N atoms)
//   pipeline    warp 0 TMA producer (dY ring of 2-6 slots, X ring of 2-12 stages: 14 boxes of smem shared out per CTA
//               by the length of its tap group) | warp 1 MMA issuer | warps 2-5 epilogue
#include "common.cuh"
#include "ryolo_b200.h"
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

namespace {

constexpr int kThreads = 192;
constexpr int kMaxBStages = 12;
constexpr int kMaxASlots = 6;
constexpr int kSmemBoxes = 14;              // 14 x 16 KB boxes + 1 KB alignment slack = 225 KB of the 227 KB a CTA may own

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
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
// one lane of a converged warp (see csrc/conv.cu: single-thread issue loops are written warp-uniform + elect)
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
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
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

// MN-major SWIZZLE_128B descriptor: 64-element (128-byte) rows along MN, one row per K index, 8-row (1024-byte)
// K groups (SBO); the next 64 MN elements live `lbo_bytes` further (LBO).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// MN-major SWIZZLE_64B: 32-element (64-byte) rows along MN, 8-row K groups 512 bytes apart.  Used for the X operand of
// Cin == 32 layers: a 64-channel box over a 32-channel tensor overhangs it, and TMA serves overhanging rows on a ~2x
// slower path (tools/oob_probe.py); a [pixels x 32 channels] box is one canonical SWIZZLE_64B MN-major atom column.
__device__ __forceinline__ uint64_t umma_desc_mn_sw64(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
// D fp32, A/B bf16, both MN-major, M = 128, N = n
__device__ __forceinline__ uint32_t umma_idesc_bf16_mn(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}

struct WgradParams {
  int N, Ho, Wo, Cin, Cout;
  int TH, TW, tiles_h, tiles_w;
  int ksize, stride, pad, ntaps;
  int nb, tg;                               // X boxes per item (N = xw*nb), taps per item
  int xw;                                   // channels per X box: 64 (SWIZZLE_128B rows) or 32 (SWIZZLE_64B rows, Cin == 32)
  int tap_grp;                              // taps issued per MMA (tap_grp * 64 * nb <= 256)
  int n_co_tiles, n_ci_tiles, n_tap_groups;
  int ks_total;                             // CTAs per (Cout tile, Cin tile) pair = sum of the tap groups' split counts
  int ks_first[10];                         // tap group g owns CTAs [ks_first[g], ks_first[g+1]) of a pair's ks_total
  int a_boxes;                              // 64-channel dY boxes per A slot: 2, or 1 when Cout <= 64 (upper atom = a shared zero box)
  int slot_rows;                            // rows per smem box slot: 128, or 64 for wide Cin tiles (deeper ring)
  int smem_boxes;                           // 16 KB boxes this CTA owns (knob wg_boxes: 14 = the whole SM, fewer leaves room for co-resident blocks)
  uint32_t tmem_cols;
  int dbg;                                  // RYOLO_WG_DBG timing experiments: 1 = no MMAs, 2 = no X loads (results are wrong)
  float* dw;                                // fp32 K-major [Cout][k*k][Cin] (16-byte aligned)
};

__global__ void __launch_bounds__(kThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX,
                  const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t kBoxBytes = (uint32_t)p.slot_rows * 128u;         // one [slot_rows pixels x 64 channels] bf16 box
  __shared__ __align__(8) uint64_t bars[2 * kMaxBStages + 2 * kMaxASlots + 1];
  __shared__ uint32_t tmem_slot;
  const uint32_t bar_bfull = smem_u32(&bars[0]), bar_bempty = smem_u32(&bars[kMaxBStages]),
                 bar_afull = smem_u32(&bars[2 * kMaxBStages]), bar_aempty = smem_u32(&bars[2 * kMaxBStages + kMaxASlots]),
                 bar_acc = smem_u32(&bars[2 * kMaxBStages + 2 * kMaxASlots]);
  // warp index through a shuffle: provably warp-uniform, so the role loops keep their state in uniform registers (csrc/conv.cu)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  // split-K shares are proportional to the taps a group carries (a 9-tap layer with tg = 8 has groups of 8 and 1 taps)
  const int pr = blockIdx.x % p.ks_total;
  const int pair = blockIdx.x / p.ks_total;
  int tgi = 0;
  while (tgi + 1 < p.n_tap_groups && pr >= p.ks_first[tgi + 1]) tgi++;
  const int split = pr - p.ks_first[tgi];
  const int ksplit = p.ks_first[tgi + 1] - p.ks_first[tgi];
  const int cit = pair % p.n_ci_tiles;
  const int cot = pair / p.n_ci_tiles;
  const int tap0 = tgi * p.tg, ntap = min(p.tg, p.ntaps - tap0);
  const int co0 = cot * 128, ci0 = cit * p.xw * p.nb;
  const int CIT = p.xw * p.nb;
  // smem partition of THIS CTA (in 16 KB boxes): [A ring: a_slots x a_boxes][zero box when a_boxes == 1][B ring].  A patch
  // costs a_boxes + ntap*nb boxes; a CTA whose tap group is short keeps more patches in flight (with two A slots a
  // one-tap item ran at two patches per memory round trip).
  const int zero_box = p.a_boxes == 1 ? 1 : 0;
  int a_slots = (p.smem_boxes - zero_box) / (p.a_boxes + ntap * p.nb);
  a_slots = max(2, min(kMaxASlots, a_slots));
  const int b_stages = min(kMaxBStages, (p.smem_boxes - zero_box - a_slots * p.a_boxes) / p.nb);
  const uint32_t sA = smem_base;
  const uint32_t sZ = sA + (uint32_t)(a_slots * p.a_boxes) * kBoxBytes;           // all-zero box (a_boxes == 1)
  const uint32_t sB = sZ + (uint32_t)zero_box * kBoxBytes;
  const int n_patches = p.N * p.tiles_h * p.tiles_w;
  const int rows = p.TH * p.TW;
  const uint32_t box_bytes = (uint32_t)rows * 128u;
  const uint32_t xrow = (uint32_t)p.xw * 2u;                       // bytes per pixel row of an X box (128 or 64)

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmG);
    prefetch_tmap(&tmX);
    for (int s = 0; s < b_stages; s++) {
      mbar_init(bar_bfull + 8 * s, 1);
      mbar_init(bar_bempty + 8 * s, 1);
    }
    for (int s = 0; s < a_slots; s++) {
      mbar_init(bar_afull + 8 * s, 1);
      mbar_init(bar_aempty + 8 * s, 1);
    }
    mbar_init(bar_acc, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_slot), p.tmem_cols);
    tmem_relinquish();
  }
  {
    // rows the TMA boxes never write but the last K step reads must be zeros; so must the whole shared zero box
    uint8_t* base = smem_raw + (smem_base - smem_u32(smem_raw));
    // (an X box with 64-byte rows occupies the first half of its slot: everything after its last row is cleared)
    const int first_x = a_slots * p.a_boxes + zero_box;            // boxes [first_x, smem_boxes) are X boxes
    const int tail16 = (int)((kBoxBytes - (uint32_t)rows * xrow) / 16u);   // 16-byte words after the last row of an X box (>= an A box's tail)
    if (rows < ((rows + 15) & ~15) || p.xw != 64) {
      for (int i = threadIdx.x; i < p.smem_boxes * tail16; i += kThreads) {
        const int b = i / tail16, w = i - b * tail16;
        const uint32_t rb = b >= first_x ? xrow : 128u;
        const uint32_t off = (uint32_t)rows * rb + (uint32_t)w * 16u;
        if (off < kBoxBytes)
          *reinterpret_cast<uint4*>(base + (size_t)b * kBoxBytes + off) = make_uint4(0, 0, 0, 0);
      }
    }
    if (zero_box) {
      for (int i = threadIdx.x; i < (int)(kBoxBytes / 16); i += kThreads)
        *reinterpret_cast<uint4*>(base + (size_t)(sZ - smem_base) + (size_t)i * 16) = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async();
  }
  const bool a_two = p.a_boxes == 2 && co0 + 64 < p.Cout;   // the tile's upper 64 channels exist
  if (p.a_boxes == 2 && !a_two) {                            // last Cout tile of a wide layer: its upper boxes stay zero
    uint8_t* base = smem_raw + (smem_base - smem_u32(smem_raw));
    for (int i = threadIdx.x; i < a_slots * (int)(kBoxBytes / 16); i += kThreads) {
      const int slot = i / (int)(kBoxBytes / 16), w = i - slot * (int)(kBoxBytes / 16);
      *reinterpret_cast<uint4*>(base + (size_t)(slot * 2 + 1) * kBoxBytes + (size_t)w * 16) = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async();
  }
  if (threadIdx.x == 0) ry_pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  ry_pdl_wait();                                // nothing above touches global memory
  const uint32_t tmem_base = tmem_slot;
  const bool has_work = split < n_patches;

  if (warp == 0) {
    // TMA producer: warp-uniform loop, one elected lane issues
    int s = 0, ab = 0;
    uint32_t bphase = 0, aphase = 0;
    const uint32_t b_tx = (p.dbg & 2) ? 0u : (uint32_t)p.nb * (uint32_t)rows * xrow;
    for (int patch = split; patch < n_patches; patch += ksplit) {
      const int pw = patch % p.tiles_w;
      const int ph = (patch / p.tiles_w) % p.tiles_h;
      const int img = patch / (p.tiles_w * p.tiles_h);
      const int h0 = ph * p.TH, w0 = pw * p.TW;
      const uint32_t a_dst = sA + (uint32_t)(ab * p.a_boxes) * kBoxBytes;
      mbar_wait(bar_aempty + 8 * ab, aphase ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(bar_afull + 8 * ab, a_two ? 2 * box_bytes : box_bytes);
        tma_load_4d(a_dst, &tmG, bar_afull + 8 * ab, co0, w0, h0, img);
        if (a_two) tma_load_4d(a_dst + kBoxBytes, &tmG, bar_afull + 8 * ab, co0 + 64, w0, h0, img);
      }
      __syncwarp();
      int kh = tap0 / p.ksize, kw = tap0 - kh * p.ksize;
      for (int ti = 0; ti < ntap; ti++) {
        mbar_wait(bar_bempty + 8 * s, bphase ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(bar_bfull + 8 * s, b_tx);
          if (!(p.dbg & 2)) {
            for (int j = 0; j < p.nb; j++)
              tma_load_4d(sB + (uint32_t)(s * p.nb + j) * kBoxBytes, &tmX, bar_bfull + 8 * s, ci0 + p.xw * j,
                          w0 * p.stride + kw - p.pad, h0 * p.stride + kh - p.pad, img);
          }
        }
        __syncwarp();
        if (++s == b_stages) { s = 0; bphase ^= 1u; }
        if (++kw == p.ksize) { kw = 0; kh++; }
      }
      if (++ab == a_slots) { ab = 0; aphase ^= 1u; }
    }
  } else if (warp == 1) {
    // MMA issuer: warp-uniform loop, one elected lane issues.  Descriptors of the K steps differ only in the 14-bit
    // start-address field (bytes >> 4): one 64-bit add per MMA instead of rebuilding them.
    const int ksteps = (p.dbg & 1) ? 0 : (rows + 15) / 16;
    int s = 0, ab = 0;
    uint32_t bphase = 0, aphase = 0, pit = 0;
    for (int patch = split; patch < n_patches; patch += ksplit, pit++) {
      mbar_wait(bar_afull + 8 * ab, aphase);
      tc_fence_after();
      const uint32_t a0 = sA + (uint32_t)(ab * p.a_boxes) * kBoxBytes;
      const uint32_t a_lbo = p.a_boxes == 2 ? kBoxBytes : sZ - a0;      // upper 64 channels: next box, or the zero box
      const uint64_t adesc = umma_desc_mn_sw128(a0, a_lbo);
      // Narrow Cin tiles are issued several taps at a time: consecutive ring slots are consecutive 64-channel N atoms
      // (LBO = one box) and the taps' accumulators are consecutive TMEM columns, so ONE N = g*CIT MMA covers g taps.
      for (int ti = 0; ti < ntap;) {
        int g = min(p.tap_grp, ntap - ti);
        g = min(g, b_stages - s);                              // a group never wraps around the ring
        for (int i = 0; i < g; i++) mbar_wait(bar_bfull + 8 * (s + i), bphase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t b0 = sB + (uint32_t)(s * p.nb) * kBoxBytes;
          const uint64_t bdesc = p.xw == 64 ? umma_desc_mn_sw128(b0, kBoxBytes) : umma_desc_mn_sw64(b0, kBoxBytes);
          const uint32_t bstep = xrow;                         // one K step = 16 pixel rows = 16 * xrow bytes = xrow units
          const uint32_t idesc_g = umma_idesc_bf16_mn(g * CIT);
          const uint32_t d = tmem_base + (uint32_t)(ti * CIT);
          for (int kk = 0; kk < ksteps; kk++)                  // dY: one K step = 16 pixel rows = 2048 bytes = 128 units
            umma_bf16(d, adesc + (uint64_t)(kk * 128), bdesc + (uint64_t)((uint32_t)kk * bstep), idesc_g,
                      (pit | (uint32_t)kk) ? 1u : 0u);
          for (int i = 0; i < g; i++) umma_commit(bar_bempty + 8 * (s + i));
          if (ti + g == ntap) {
            umma_commit(bar_aempty + 8 * ab);
            if (patch + ksplit >= n_patches) umma_commit(bar_acc);
          }
        }
        __syncwarp();
        ti += g;
        s += g;
        if (s == b_stages) { s = 0; bphase ^= 1u; }
      }
      if (++ab == a_slots) { ab = 0; aphase ^= 1u; }
    }
  } else if (has_work) {
    // dw is K-major like the packed forward weights: [Cout][tap][Cin] (stem: [Cout][Kpad]).  A lane owns one Cout row
    // and adds 4 consecutive Cin values per vector atomic (red.global.add.v4.f32).
    const int sub = warp & 3;
    const int co = co0 + sub * 32 + lane;
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    const int rowlen = p.ntaps * p.Cin;
    for (int ti = 0; ti < ntap; ti++) {
      const int tap = tap0 + ti;
#pragma unroll 1
      for (int c0 = 0; c0 < CIT; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(ti * CIT + c0), v);
        tmem_ld_wait();
        if (co < p.Cout) {
          float* row = p.dw + (long long)co * rowlen + (long long)tap * p.Cin + ci0 + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (ci0 + c0 + j < p.Cin)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(row + j), "f"(__uint_as_float(v[j])),
                           "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])),
                           "f"(__uint_as_float(v[j + 3]))
                           : "memory");
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// ------------------------------------------------------------------------------------ transposing variant (Cin <= 128)
// MN-major operands cost the tensor core ~2x the smem fetch time of K-major ones, independent of N, which leaves the
// narrow layers (N = 64 / 128) at a few percent of the MMA rate (measured: ~0.9 us per tap and patch whatever N is).
// Here the four warps that otherwise idle until the epilogue TRANSPOSE every TMA box in shared memory
// (ldmatrix.trans -> stmatrix, 128-byte swizzle on both sides) into K-major tiles [channels][pixels], so that the MMAs
// run on K-major descriptors with K = pixels:
//   RAW ring   4 boxes [128 px][64 ch]                  TMA destination, freed as soon as it has been transposed
//   TA  ring   2 slots [2 k-halves][128 co][64 px]      dY^T of a patch (rows 64..127 stay zero when the tile has <= 64 co)
//   TB  ring   3 stages [2 k-halves][128 rows][64 px]   X^T of two boxes = N <= 128 (two taps for Cin <= 64, one for Cin <= 128)
// Same work decomposition, TMEM layout and epilogue as conv_wgrad_kernel.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// D fp32, A/B bf16, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t umma_idesc_bf16_k(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void stmatrix_x4(uint32_t addr, const uint32_t (&r)[4]) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1, %2, %3, %4};"
               ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}

constexpr int kRawSlots = 4, kTaSlots = 2, kTbStages = 3;

__global__ void __launch_bounds__(kThreads, 1)
conv_wgrad_t_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX,
                    const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr uint32_t kBox = 16384u;
  const uint32_t sRaw = smem_base;                          // 4 boxes
  const uint32_t sTA = sRaw + kRawSlots * kBox;             // 2 slots x 32 KB
  const uint32_t sTB = sTA + kTaSlots * 2 * kBox;           // 3 stages x 32 KB
  __shared__ __align__(8) uint64_t bars[2 * kRawSlots + 2 * kTaSlots + 2 * kTbStages + 1];
  __shared__ uint32_t tmem_slot;
  const uint32_t bar_rfull = smem_u32(&bars[0]), bar_rempty = smem_u32(&bars[kRawSlots]),
                 bar_afull = smem_u32(&bars[2 * kRawSlots]), bar_aempty = smem_u32(&bars[2 * kRawSlots + kTaSlots]),
                 bar_bfull = smem_u32(&bars[2 * kRawSlots + 2 * kTaSlots]),
                 bar_bempty = smem_u32(&bars[2 * kRawSlots + 2 * kTaSlots + kTbStages]),
                 bar_acc = smem_u32(&bars[2 * kRawSlots + 2 * kTaSlots + 2 * kTbStages]);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform (see above)

  const int pr = blockIdx.x % p.ks_total;
  const int pair = blockIdx.x / p.ks_total;
  int tgi = 0;
  while (tgi + 1 < p.n_tap_groups && pr >= p.ks_first[tgi + 1]) tgi++;
  const int split = pr - p.ks_first[tgi];
  const int ksplit = p.ks_first[tgi + 1] - p.ks_first[tgi];
  const int cit = pair % p.n_ci_tiles;
  const int cot = pair / p.n_ci_tiles;
  const int tap0 = tgi * p.tg, ntap = min(p.tg, p.ntaps - tap0);
  const int co0 = cot * 128, ci0 = cit * 64 * p.nb;
  const int CIT = 64 * p.nb;
  const int n_patches = p.N * p.tiles_h * p.tiles_w;
  const int rows = p.TH * p.TW;
  const uint32_t box_bytes = (uint32_t)rows * 128u;
  const int na = (co0 + 64 < p.Cout) ? 2 : 1;              // dY boxes per patch
  const int nxb = ntap * p.nb;                             // X boxes per patch
  const int nst = (nxb + 1) >> 1;                          // TB stages per patch (two boxes each, the last may hold one)

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmG);
    prefetch_tmap(&tmX);
    for (int s = 0; s < kRawSlots; s++) { mbar_init(bar_rfull + 8 * s, 1); mbar_init(bar_rempty + 8 * s, 4); }
    for (int s = 0; s < kTaSlots; s++) { mbar_init(bar_afull + 8 * s, 4 * na); mbar_init(bar_aempty + 8 * s, 1); }
    for (int s = 0; s < kTbStages; s++) { mbar_init(bar_bfull + 8 * s, 8); mbar_init(bar_bempty + 8 * s, 1); }
    mbar_init(bar_acc, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_slot), p.tmem_cols);
    tmem_relinquish();
  }
  {
    uint8_t* base = smem_raw + (smem_base - smem_u32(smem_raw));
    // pixels a TMA box never writes (rows >= TH*TW) become K columns of the transposed tiles: keep them zero
    const int tail16 = (128 - rows) * 8;
    for (int i = threadIdx.x; i < kRawSlots * tail16; i += kThreads) {
      const int b = i / tail16, w = i - b * tail16;
      *reinterpret_cast<uint4*>(base + (size_t)b * kBox + (size_t)rows * 128 + (size_t)w * 16) = make_uint4(0, 0, 0, 0);
    }
    if (na == 1) {      // co rows 64..127 of both k-halves of both TA slots
      for (int i = threadIdx.x; i < kTaSlots * 2 * 512; i += kThreads) {
        const int h = i / 512, w = i - h * 512;              // h = slot*2 + k-half, 512 x 16 B = rows 64..127
        *reinterpret_cast<uint4*>(base + (size_t)(sTA - smem_base) + (size_t)h * kBox + 8192 + (size_t)w * 16) =
            make_uint4(0, 0, 0, 0);
      }
    }
    fence_proxy_async();
  }
  if (threadIdx.x == 0) ry_pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  ry_pdl_wait();                                // nothing above touches global memory
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int rs = 0;
      uint32_t rphase = 0;
      auto load = [&](const CUtensorMap* tm, int c0, int c1, int c2, int c3) {
        mbar_wait(bar_rempty + 8 * rs, rphase ^ 1u);
        mbar_expect_tx(bar_rfull + 8 * rs, box_bytes);
        tma_load_4d(sRaw + (uint32_t)rs * kBox, tm, bar_rfull + 8 * rs, c0, c1, c2, c3);
        if (++rs == kRawSlots) { rs = 0; rphase ^= 1u; }
      };
      for (int patch = split; patch < n_patches; patch += ksplit) {
        const int pw = patch % p.tiles_w;
        const int ph = (patch / p.tiles_w) % p.tiles_h;
        const int img = patch / (p.tiles_w * p.tiles_h);
        const int h0 = ph * p.TH, w0 = pw * p.TW;
        for (int j = 0; j < na; j++) load(&tmG, co0 + 64 * j, w0, h0, img);
        for (int ti = 0; ti < ntap; ti++) {
          const int tap = tap0 + ti;
          const int kh = tap / p.ksize, kw = tap - kh * p.ksize;
          for (int j = 0; j < p.nb; j++)
            load(&tmX, ci0 + 64 * j, w0 * p.stride + kw - p.pad, h0 * p.stride + kh - p.pad, img);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aphase = 0, bphase = 0, pit = 0;
      for (int patch = split; patch < n_patches; patch += ksplit, pit++) {
        mbar_wait(bar_afull + 8 * as, aphase);
        tc_fence_after();
        const uint32_t a0 = sTA + (uint32_t)as * 2 * kBox;
        for (int st = 0; st < nst; st++) {
          const int nbox = min(2, nxb - 2 * st);
          mbar_wait(bar_bfull + 8 * bs, bphase);
          tc_fence_after();
          const uint32_t b0 = sTB + (uint32_t)bs * 2 * kBox;
          const uint32_t idesc = umma_idesc_bf16_k(64 * nbox);
          const uint32_t d = tmem_base + (uint32_t)(st * 128);           // boxes are consecutive 64-column blocks
#pragma unroll
          for (int kk = 0; kk < 8; kk++) {                                // 2 k-halves x 4 K=16 steps
            const uint32_t off = (uint32_t)(kk >> 2) * kBox + (uint32_t)(kk & 3) * 32u;
            umma_bf16(d, umma_desc_k_sw128(a0 + off), umma_desc_k_sw128(b0 + off), idesc, (pit | (uint32_t)kk) ? 1u : 0u);
          }
          umma_commit(bar_bempty + 8 * bs);
          if (++bs == kTbStages) { bs = 0; bphase ^= 1u; }
        }
        umma_commit(bar_aempty + 8 * as);
        if (++as == kTaSlots) { as = 0; aphase ^= 1u; }
      }
      umma_commit(bar_acc);
    }
  } else {
    // ---------------- transposers (4 warps), then the epilogue
    const int tw = warp - 2;
    int rs = 0, as = 0, bs = 0;
    uint32_t rphase = 0, aphase = 0, bphase = 0;
    // one box: [128 px][64 ch] (SW128) -> two K-major tiles [64 ch][64 px] (SW128) at dst + khalf*16 KB + row0*128
    auto transpose_box = [&](uint32_t dst, int row0) {
      mbar_wait(bar_rfull + 8 * rs, rphase);
      const uint32_t src = sRaw + (uint32_t)rs * kBox;
      const int i = lane >> 3, r = lane & 7;
#pragma unroll
      for (int q8 = 0; q8 < 8; q8++) {
        const int q = tw + 4 * q8;                  // 32 (4-tile) jobs per box, 8 per warp
        const int cb = q & 7, pb0 = (q >> 3) * 4;   // channel block, first of four pixel blocks
        uint32_t v[4];
        ldmatrix_x4_trans(src + (uint32_t)(8 * (pb0 + i) + r) * 128u + (uint32_t)((cb ^ r) << 4), v);
        stmatrix_x4(dst + (uint32_t)(pb0 >> 3) * kBox + (uint32_t)(row0 + 8 * cb + r) * 128u +
                        (uint32_t)((((pb0 + i) & 7) ^ r) << 4), v);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_rempty + 8 * rs);
      if (++rs == kRawSlots) { rs = 0; rphase ^= 1u; }
    };
    for (int patch = split; patch < n_patches; patch += ksplit) {
      mbar_wait(bar_aempty + 8 * as, aphase ^ 1u);
      for (int j = 0; j < na; j++) {
        transpose_box(sTA + (uint32_t)as * 2 * kBox, 64 * j);
        if (lane == 0) mbar_arrive(bar_afull + 8 * as);
      }
      if (++as == kTaSlots) { as = 0; aphase ^= 1u; }
      for (int st = 0; st < nst; st++) {
        const int nbox = min(2, nxb - 2 * st);
        mbar_wait(bar_bempty + 8 * bs, bphase ^ 1u);
        for (int j = 0; j < 2; j++) {
          if (j < nbox) transpose_box(sTB + (uint32_t)bs * 2 * kBox, 64 * j);
          if (lane == 0) mbar_arrive(bar_bfull + 8 * bs);     // a missing second box still counts (fixed arrival count)
        }
        if (++bs == kTbStages) { bs = 0; bphase ^= 1u; }
      }
    }
    const int sub = warp & 3;
    const int co = co0 + sub * 32 + lane;
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    const int rowlen = p.ntaps * p.Cin;
    for (int ti = 0; ti < ntap; ti++) {
      const int tap = tap0 + ti;
#pragma unroll 1
      for (int c0 = 0; c0 < CIT; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(ti * CIT + c0), v);
        tmem_ld_wait();
        if (co < p.Cout) {
          float* row = p.dw + (long long)co * rowlen + (long long)tap * p.Cin + ci0 + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (ci0 + c0 + j < p.Cin)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(row + j), "f"(__uint_as_float(v[j])),
                           "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])),
                           "f"(__uint_as_float(v[j + 3]))
                           : "memory");
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

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

void pick_patch(int Ho, int Wo, int cap, int* TH, int* TW) {
  double best = -1.0;
  for (int tw = 1; tw <= cap && tw <= Wo; tw++) {
    int th = cap / tw;
    if (th > Ho) th = Ho;
    const double tiles = (double)((Ho + th - 1) / th) * ((Wo + tw - 1) / tw);
    const double eff = (double)Ho * Wo / (tiles * (double)cap);
    if (eff > best + 1e-9 || (eff > best - 1e-9 && tw > *TW)) { best = eff; *TH = th; *TW = tw; }
  }
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

int encode_nhwc(EncodeTiledFn enc, CUtensorMap* tm, const void* ptr, int N, int H, int W, int C, long long cpitch,
                int TH, int TW, int estride, int boxc = 64) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)cpitch * 2, (cuuint64_t)cpitch * 2 * W, (cuuint64_t)cpitch * 2 * W * H};
  cuuint32_t box[4] = {(cuuint32_t)boxc, (cuuint32_t)(TW * estride), (cuuint32_t)(TH * estride), 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)ptr, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, boxc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 1;
}

}  // namespace

extern "C" {

// dwk (fp32, K-major [Cout][kh*kw][Cin] = the layout of the packed forward weights; for the stem x is the Kpad-channel
// im2col tensor and dwk is [Cout][Kpad]) += conv_backward_weight(x, dy).   x: bf16 NHWC view [N,H,W,Cin]; dy: bf16 NHWC
// view [N,Ho,Wo,Cdy] with Cdy >= Cout channels readable (extra channels are ignored).  dwk must be zeroed (or hold a
// running sum) on entry and be 16-byte aligned; ryolo_unpack_wgrad_multi adds it into the OIHW gradients.
int ryolo_conv2d_wgrad(const void* x, long long x_cpitch, int N, int H, int W, int Cin, const void* dy,
                       long long dy_cpitch, int Cdy, int Cout, int ksize, int stride, float* dwk, void* stream) {
  RY_CHECK_ARG(ksize == 1 || ksize == 3, "wgrad: ksize must be 1 or 3");
  RY_CHECK_ARG(stride == 1 || stride == 2, "wgrad: stride must be 1 or 2");
  RY_CHECK_ARG(Cin % 8 == 0 && Cdy % 8 == 0 && x_cpitch % 8 == 0 && dy_cpitch % 8 == 0 && Cdy >= Cout,
               "wgrad: channel counts and pitches must be multiples of 8");
  RY_CHECK_ARG((((uintptr_t)x) & 15) == 0 && (((uintptr_t)dy) & 15) == 0 && (((uintptr_t)dwk) & 15) == 0,
               "wgrad: operands must be 16-byte aligned");
  if (N == 0) return RYOLO_OK;
  EncodeTiledFn enc = get_encode();
  if (!enc) { ryolo_set_error("cuTensorMapEncodeTiled not available from the driver"); return RYOLO_ERR_CUDA; }
  WgradParams p{};
  p.N = N; p.Cin = Cin; p.Cout = Cout; p.ksize = ksize; p.stride = stride; p.pad = (ksize - 1) / 2;
  p.Ho = (H + 2 * p.pad - ksize) / stride + 1;
  p.Wo = (W + 2 * p.pad - ksize) / stride + 1;
  p.ntaps = ksize * ksize;
  p.nb = (Cin + 63) / 64;
  if (p.nb > 4) p.nb = 4;
  p.slot_rows = 128;   // (64-row K blocks were measured slower: more barrier round trips per byte)
  p.TH = 1; p.TW = 1;
  pick_patch(p.Ho, p.Wo, p.slot_rows, &p.TH, &p.TW);
  if (p.TW * stride > 256) { p.TW = 256 / stride; p.TH = p.slot_rows / p.TW; if (p.TH > p.Ho) p.TH = p.Ho; }
  p.tiles_h = (p.Ho + p.TH - 1) / p.TH;
  p.tiles_w = (p.Wo + p.TW - 1) / p.TW;
  const bool trans = ryolo_knob(RYOLO_KNOB_WG_TRANS) && p.nb <= 2;
  // knob wg_x32: X boxes of Cin == 32 layers are 32 channels wide (SWIZZLE_64B rows): no overhanging TMA boxes, half the
  // operand bytes, and up to eight taps per N = 256 MMA
  p.xw = (Cin == 32 && !trans && ryolo_knob(RYOLO_KNOB_WG_X32)) ? 32 : 64;
  const int CIT = p.xw * p.nb;
  p.n_ci_tiles = (Cin + CIT - 1) / CIT;
  p.n_co_tiles = (Cout + 127) / 128;
  p.tg = 512 / CIT;
  if (p.tg > p.ntaps) p.tg = p.ntaps;
  p.n_tap_groups = (p.ntaps + p.tg - 1) / p.tg;
  p.tap_grp = ryolo_knob(RYOLO_KNOB_WG_TAPGRP) ? (256 / CIT > 0 ? 256 / CIT : 1) : 1;
  uint32_t cols = 32;
  while (cols < (uint32_t)(p.tg * CIT)) cols <<= 1;
  p.tmem_cols = cols;
  const int items = p.n_co_tiles * p.n_ci_tiles * p.n_tap_groups;
  const long long n_patches = (long long)N * p.tiles_h * p.tiles_w;
  RY_CHECK_ARG(n_patches < (1ll << 31), "wgrad: too many patches");
  // CTAs per (Cout tile, Cin tile) pair are dealt to the tap groups by cost (knob wg_split: 1 cost model, 2 tap counts,
  // 0 uniform).
  const int prop = ryolo_knob(RYOLO_KNOB_WG_SPLIT);
  const int pairs = p.n_co_tiles * p.n_ci_tiles;
  // Measured cost of one patch for an item of t taps (profiles/r01_wgrad_cost_model.txt): the MMA chain costs ~0.11 us
  // per MMA whatever its N (8 K steps per group of tap_grp taps), the TMA side ~0.29 us per 16 KB box; an item runs at
  // the slower of the two.  CTAs go to the item with the highest cost per CTA (greedy), so a one-tap leftover group gets
  // about a third of an eight-tap group's CTAs, not an eighth.
  int ks[9];
  double gw[9];
  const int a_boxes = Cout > 64 ? 2 : 1;
  for (int g = 0; g < p.n_tap_groups; g++) {
    ks[g] = 1;
    const int t = p.ntaps - g * p.tg < p.tg ? p.ntaps - g * p.tg : p.tg;
    const double mma = 8.0 * ((t + p.tap_grp - 1) / p.tap_grp) * 0.11;
    const double tma = (double)(t * p.nb + a_boxes) * 0.29;
    gw[g] = prop == 2 ? (double)t : (mma > tma ? mma : tma);
  }
  if (prop) {
    int budget_ctas = sm_count() / pairs;
    for (int used = p.n_tap_groups; used < budget_ctas; used++) {
      int best = 0;
      for (int g = 1; g < p.n_tap_groups; g++)
        if (gw[g] * ks[best] > gw[best] * ks[g]) best = g;
      ks[best]++;
    }
  } else {
    int u = sm_count() / items;
    if (u < 1) u = 1;
    for (int g = 0; g < p.n_tap_groups; g++) ks[g] = u;
  }
  p.ks_first[0] = 0;
  for (int g = 0; g < p.n_tap_groups; g++) {
    if (ks[g] > n_patches) ks[g] = (int)n_patches;
    p.ks_first[g + 1] = p.ks_first[g] + ks[g];
  }
  p.ks_total = p.ks_first[p.n_tap_groups];
  p.a_boxes = Cout > 64 ? 2 : 1;
  const size_t kBoxBytes = (size_t)p.slot_rows * 128;
  int boxes = ryolo_knob(RYOLO_KNOB_WG_BOXES);
  p.smem_boxes = boxes < 10 ? 10 : (boxes > kSmemBoxes ? kSmemBoxes : boxes);
  if (trans) p.smem_boxes = kSmemBoxes;
  const size_t smem = 1024 + (size_t)p.smem_boxes * kBoxBytes, smem_max = 1024 + (size_t)kSmemBoxes * kBoxBytes;
  p.dw = dwk;
  p.dbg = ryolo_knob(RYOLO_KNOB_WG_DBG);
  CUtensorMap tmG, tmX;
  if (encode_nhwc(enc, &tmG, dy, N, p.Ho, p.Wo, Cdy, dy_cpitch, p.TH, p.TW, 1) ||
      encode_nhwc(enc, &tmX, x, N, H, W, Cin, x_cpitch, p.TH, p.TW, stride, p.xw)) {
    ryolo_set_error("cuTensorMapEncodeTiled failed in wgrad");
    return RYOLO_ERR_CUDA;
  }
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_wgrad_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
    if (e != cudaSuccess) { ryolo_set_error(cudaGetErrorString(e)); return RYOLO_ERR_CUDA; }
    configured = true;
  }
  if (trans)
    ry_launch(conv_wgrad_t_kernel, dim3(pairs * p.ks_total), dim3(kThreads), smem, (cudaStream_t)stream, tmG, tmX, p);
  else
    ry_launch(conv_wgrad_kernel, dim3(pairs * p.ks_total), dim3(kThreads), smem, (cudaStream_t)stream, tmG, tmX, p);
  RY_CHECK_LAUNCH();
  return RYOLO_OK;
}

}  // extern "C"
