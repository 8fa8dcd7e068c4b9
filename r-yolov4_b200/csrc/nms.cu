// Rotated NMS / pairwise skew IoU / post_process front-end  (K10, K11, K12 in SURVEY.md §2.1).
//
// Replaces, for the reference's post_process (lib/general.py:136-183):
//   :155-161  cls*=obj, class max, confidence filter        -> pp_score_kernel      (HBM-bound)
//   :166-168  argsort(descending) + top max_nms             -> pp_select_sort_kernel (one CTA / image:
//                                                              radix select + smem bitonic sort)
//   :171-174  class offset, rad->deg                        -> fused into the gather of the sorted rows
//   :177      detectron2 nms_rotated (N^2 mask + host scan) -> nms_mask_kernel (tile compaction + warp
//                                                              ballot) + nms_scan_kernel (device scan)
//   :178-181  top max_det + gather                          -> tail of nms_scan_kernel
// No device->host traffic happens inside: counts stay on the device for the caller to read once.
//
// Compile with --fmad=false (see build.py): decisions must be bit-identical with the oracle.
#include "common.cuh"
#include "rotated_iou.cuh"

using namespace ryolo;

namespace {

constexpr int kTile = 64;          // NMS tile edge (one u64 of mask bits per row and tile)
constexpr int kSortCap = 8192;     // max keys sorted in shared memory by one CTA
constexpr float kPiF = 3.14159274101257324f;  // fl32(np.pi)

// ------------------------------------------------------------------------------------ prep
__global__ void prep_boxes_kernel(const float* __restrict__ boxes5, const int32_t* __restrict__ order,
                                  int64_t n, RPrep* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t src = order ? (int64_t)order[i] : i;
  const float* b = boxes5 + 5 * src;
  out[i] = rprep(b[0], b[1], b[2], b[3], b[4]);
}

// ------------------------------------------------------------------------------------ pairwise
__global__ void pairwise_iou_kernel(const RPrep* __restrict__ a, int64_t n, const RPrep* __restrict__ b,
                                    int64_t m, float* __restrict__ out) {
  __shared__ RPrep sa[16], sb[16];
  int tx = threadIdx.x, ty = threadIdx.y;
  int64_t j = blockIdx.x * 16 + tx, i = blockIdx.y * 16 + ty;
  if (ty == 0 && j < m) sb[tx] = b[j];
  if (ty == 1 && blockIdx.y * 16 + tx < n) sa[tx] = a[blockIdx.y * 16 + tx];
  __syncthreads();
  if (i < n && j < m) out[i * m + j] = rbox_iou(sa[ty], sb[tx]);
}

// ------------------------------------------------------------------------------------ eval matching (test.py:100-149)
// get_batch_statistics on the device, one CTA per image.  Reference semantics, kept verbatim:
//   * angles rad -> deg as (x / fl32(pi)) * 180 in fp32                                            (test.py:124-125)
//   * per target class c (ascending, torch.unique): detections of class c in index order; each takes the target of
//     its class with the highest skew IoU (first maximum, torch.max) and claims it if IoU > iouv[0] and nobody of
//     this class claimed it before; a claim marks tp[d, k] = IoU > iouv[k]                          (test.py:127-141)
//   * `if len(detected_boxes) == nl: break` leaves the CURRENT class only, and only at equality    (test.py:142-143)
// Phase 1 (all threads): best target per detection — n_det x n_tgt exact IoUs (rbox_iou = Appendix B).
// Phase 2 (one thread): the inherently sequential claim bookkeeping, <= n_classes_present * n_det steps.
constexpr int kMatchMaxTgt = 1024;     // targets of ONE image held in shared memory

__global__ void __launch_bounds__(256)
eval_match_kernel(const float* __restrict__ dets, const int32_t* __restrict__ n_det, int max_det,
                  const float* __restrict__ targets, int64_t T, int tcols, const float* __restrict__ iouv, int niou,
                  uint8_t* __restrict__ tp, float* __restrict__ best_iou, int32_t* __restrict__ best_tgt,
                  int32_t* __restrict__ status) {
  __shared__ RPrep tprep[kMatchMaxTgt];
  __shared__ float tcls[kMatchMaxTgt];
  __shared__ unsigned char claimed[kMatchMaxTgt];
  __shared__ int nt_s;
  const int img = blockIdx.x, tid = threadIdx.x;
  const int nd = min(n_det[img], max_det);
  dets += (int64_t)img * max_det * 7;
  tp += (int64_t)img * max_det * niou;
  best_iou += (int64_t)img * max_det;
  best_tgt += (int64_t)img * max_det;
  for (int i = tid; i < nd * niou; i += blockDim.x) tp[i] = 0;
  if (tid == 0) nt_s = 0;
  __syncthreads();
  // targets of this image, in row order (one thread: order matters for "first maximum" and T is small)
  if (tid == 0) {
    int n = 0;
    for (int64_t t = 0; t < T; t++) {
      const float* r = targets + t * tcols;
      if ((int)r[0] != img) continue;
      if (n < kMatchMaxTgt) {
        tprep[n] = rprep(r[2], r[3], r[4], r[5], __fmul_rn(__fdiv_rn(r[6], kPiF), 180.f));
        tcls[n] = r[1];
        claimed[n] = 0;
      }
      n++;
    }
    nt_s = n;
    if (n > kMatchMaxTgt) status[0] = 1;          // reported by the host wrapper (no silent truncation)
  }
  __syncthreads();
  const int nt = min(nt_s, kMatchMaxTgt);
  if (nd == 0 || nt == 0) return;
  for (int d = tid; d < nd; d += blockDim.x) {
    const float* r = dets + d * 7;
    const RPrep P = rprep(r[0], r[1], r[2], r[3], __fmul_rn(__fdiv_rn(r[4], kPiF), 180.f));
    float best = -1.f;
    int bi = -1;
    for (int t = 0; t < nt; t++) {
      if (tcls[t] != r[6]) continue;
      const float v = rbox_iou(P, tprep[t]);
      if (bi < 0 || v > best) { best = v; bi = t; }
    }
    best_iou[d] = best;
    best_tgt[d] = bi;
  }
  __syncthreads();
  if (tid != 0) return;
  const float thr0 = iouv[0];
  int n_detected = 0;
  float prev = -INFINITY;
  for (;;) {                                       // classes present among the targets, ascending
    float c = INFINITY;
    for (int t = 0; t < nt; t++) if (tcls[t] > prev && tcls[t] < c) c = tcls[t];
    if (c == INFINITY) break;
    prev = c;
    for (int d = 0; d < nd; d++) {
      if (dets[d * 7 + 6] != c) continue;
      const int t = best_tgt[d];
      const float v = best_iou[d];
      if (t < 0 || !(v > thr0) || claimed[t]) continue;
      claimed[t] = 1;
      n_detected++;
      for (int k = 0; k < niou; k++) tp[d * niou + k] = v > iouv[k] ? 1 : 0;
      if (n_detected == nt_s) break;
    }
  }
}

// ------------------------------------------------------------------------------------ NMS mask
// grid (col block, row block, image); 256 threads.  Phase 1 rejects far-apart pairs with a
// conservative bounding-circle test and compacts the survivors into a shared queue so that phase 2
// (the ~2 kFLOP polygon clip) runs on fully populated warps.
//
// Banded use (post_process, see ryolo_post_process): greedy NMS only ever needs the mask rows of boxes that end up KEPT,
// and only against columns that are still alive.  The candidates are walked in bands of `band` tiles; per band
//   nms_band_suppress_kernel   every box kept so far against the band's columns -> `rem` (suppressed bits), no mask stored
//   nms_mask_kernel            (this kernel, rb0 = cb0 = first tile of the band, grid band x band) the band's own
//                              triangle, rows / columns already suppressed skipped; mask words stored for the scan
//   nms_band_scan_kernel       the greedy scan of the band: kept boxes, running count, `done` once max_det is reached
// An image whose max_det keeps are complete drops out of every later launch, so no pair is evaluated for candidates
// the scan never reaches.  Same pair decisions, same greedy order: the survivors are those of the full mask.
struct BandState { int nk; int done; };
constexpr int kBandMax = 16;      // tiles per band: knob nms_band (default 8)

__global__ void __launch_bounds__(256)
nms_mask_kernel(const RPrep* __restrict__ prep, const int32_t* __restrict__ counts, int64_t prep_stride,
                int words, float thr, unsigned long long* __restrict__ mask, int64_t mask_stride,
                int rb0, int cb0, const unsigned long long* __restrict__ rem, const BandState* __restrict__ state) {
  const int cb = cb0 + blockIdx.x, rb = rb0 + blockIdx.y, img = blockIdx.z;
  if (rb > cb || cb >= words) return;
  const int K = counts[img];
  if (cb * kTile >= K) return;
  unsigned long long rowdead = 0ull, coldead = 0ull;      // bit set = this row / column cannot matter any more
  if (rem) {
    if (state[img].done) return;
    rem += (int64_t)img * words;
    coldead = rem[cb];
    rowdead = rem[rb];
    const int nvc = min(kTile, K - cb * kTile);
    const unsigned long long valid = nvc == 64 ? ~0ull : ((1ull << nvc) - 1ull);
    if (!(~coldead & valid) || !~rowdead) return;
  }
  prep += (int64_t)img * prep_stride;
  mask += (int64_t)img * mask_stride;

  __shared__ RPrep srow[kTile], scol[kTile];
  __shared__ float qx[2][kTile], qy[2][kTile], qr[2][kTile];
  __shared__ float qw[2][kTile], qh[2][kTile], qc[2][kTile], qs[2][kTile], qa[2][kTile];   // SoA for the early-outs
  __shared__ unsigned short queue[kTile * kTile], queue2[kTile * kTile];
  __shared__ unsigned long long tmask[kTile];
  __shared__ int qn, qn2;

  const int tid = threadIdx.x;
  if (tid < kTile) {
    int g = rb * kTile + tid;
    RPrep p = prep[min(g, K - 1)];
    srow[tid] = p; qx[0][tid] = p.cx; qy[0][tid] = p.cy; qr[0][tid] = p.reach;
    qw[0][tid] = p.w; qh[0][tid] = p.h; qc[0][tid] = p.c2; qs[0][tid] = p.s2; qa[0][tid] = p.area;
    tmask[tid] = 0ull;
  } else if (tid < 2 * kTile) {
    int t = tid - kTile, g = cb * kTile + t;
    RPrep p = prep[min(g, K - 1)];
    scol[t] = p; qx[1][t] = p.cx; qy[1][t] = p.cy; qr[1][t] = p.reach;
    qw[1][t] = p.w; qh[1][t] = p.h; qc[1][t] = p.c2; qs[1][t] = p.s2; qa[1][t] = p.area;
  }
  if (tid == 0) { qn = 0; qn2 = 0; }
  __syncthreads();

  const bool use_reject = thr >= 0.f;  // IoU==0 pairs only matter when thr < 0
  const bool use_bounds = thr >= 1e-6f; // area-ratio bound + separating-axis test (see rbox_cannot_exceed)
  const int lane = tid & 31;
#pragma unroll 4
  for (int p = tid; p < kTile * kTile; p += 256) {
    int i = p >> 6, j = p & 63;
    int gi = rb * kTile + i, gj = cb * kTile + j;
    bool live = (gi < gj) && (gj < K) && !((rowdead >> i) & 1ull) && !((coldead >> j) & 1ull);
    if (live && use_reject) {
      live = !rbox_far_soa(qx[0][i], qy[0][i], qw[0][i], qh[0][i], qr[0][i], qx[1][j], qy[1][j], qw[1][j], qh[1][j],
                           qr[1][j]);
      if (live && use_bounds)
        live = !rbox_cannot_exceed(qx[0][i], qy[0][i], qw[0][i], qh[0][i], qc[0][i], qs[0][i], qa[0][i], qx[1][j],
                                   qy[1][j], qw[1][j], qh[1][j], qc[1][j], qs[1][j], qa[1][j], thr);
    }
    unsigned b = __ballot_sync(0xffffffffu, live);
    if (b) {
      int base = 0;
      if (lane == 0) base = atomicAdd(&qn, __popc(b));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (live) queue[base + __popc(b & ((1u << lane) - 1))] = (unsigned short)p;
    }
  }
  __syncthreads();
  const int nq = qn;
  // phase 2: cheap fp32 estimate where its error is provably small (rbox_fast_ok) and far from the threshold; the
  // undecided pairs (~1 %) are compacted into a second queue ...
  for (int q0 = 0; q0 < nq; q0 += 256) {
    const int q = q0 + tid;
    int verdict = 0, p = 0;
    if (q < nq) {
      p = queue[q];
      verdict = rbox_iou_exceeds_quick(srow[p >> 6], scol[p & 63], thr);
      if (verdict > 0) atomicOr(&tmask[p >> 6], 1ull << (p & 63));
    }
    const unsigned b = __ballot_sync(0xffffffffu, verdict < 0);
    if (b) {
      int base = 0;
      if (lane == 0) base = atomicAdd(&qn2, __popc(b));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (verdict < 0) queue2[base + __popc(b & ((1u << lane) - 1))] = (unsigned short)p;
    }
  }
  __syncthreads();
  // ... phase 3: and take the bit-exact Appendix-B path on fully populated warps
  const int nq2 = qn2;
  for (int q = tid; q < nq2; q += 256) {
    const int p = queue2[q], i = p >> 6, j = p & 63;
    if (rbox_iou_full(srow[i], scol[j]) > thr) atomicOr(&tmask[i], 1ull << j);
  }
  __syncthreads();
  if (tid < kTile) {
    int gi = rb * kTile + tid;
    if (gi < K) mask[(int64_t)gi * words + cb] = tmask[tid];
  }
}

// One block per (column tile of the band, image): streams the boxes kept so far (keep_pos[0 .. nk), 64 at a time, best
// scores first) past the tile's 64 columns and marks the columns they suppress.  Columns die progressively — the
// top-scoring keeps suppress most of what can be suppressed — and the loop stops when none is left alive, so a pair is
// only evaluated between a KEPT box and a column that is still alive when its turn comes.
__global__ void __launch_bounds__(256)
nms_band_suppress_kernel(const RPrep* __restrict__ prep, const int32_t* __restrict__ counts, int64_t prep_stride,
                         int words, float thr, int cb0, unsigned long long* __restrict__ rem,
                         const BandState* __restrict__ state, const int32_t* __restrict__ keep_pos, int64_t keep_stride) {
  const int cb = cb0 + blockIdx.x, img = blockIdx.y;
  if (cb >= words || state[img].done) return;
  const int K = counts[img];
  if (cb * kTile >= K) return;
  const int nk = state[img].nk;                  // < max_det (otherwise the image is done): all of them are in keep_pos
  if (nk == 0) return;
  prep += (int64_t)img * prep_stride;
  rem += (int64_t)img * words;
  keep_pos += (int64_t)img * keep_stride;

  __shared__ RPrep srow[kTile], scol[kTile];
  __shared__ float qx[2][kTile], qy[2][kTile], qr[2][kTile];
  __shared__ float qw[2][kTile], qh[2][kTile], qc[2][kTile], qs[2][kTile], qa[2][kTile];
  __shared__ unsigned short queue[kTile * kTile], queue2[kTile * kTile];
  __shared__ unsigned long long s_dead;
  __shared__ int qn, qn2;

  const int tid = threadIdx.x, lane = tid & 31;
  const int nvc = min(kTile, K - cb * kTile);
  if (tid < kTile) {
    const RPrep p = prep[min(cb * kTile + tid, K - 1)];
    scol[tid] = p; qx[1][tid] = p.cx; qy[1][tid] = p.cy; qr[1][tid] = p.reach;
    qw[1][tid] = p.w; qh[1][tid] = p.h; qc[1][tid] = p.c2; qs[1][tid] = p.s2; qa[1][tid] = p.area;
  }
  if (tid == 0) s_dead = rem[cb] | (nvc == 64 ? 0ull : ~((1ull << nvc) - 1ull));
  const bool use_reject = thr >= 0.f;
  const bool use_bounds = thr >= 1e-6f;
  for (int r0 = 0; r0 < nk; r0 += kTile) {
    __syncthreads();                               // previous chunk done with srow / queues; s_dead up to date
    const unsigned long long dead = s_dead;
    if (!~dead) break;                             // uniform: every thread read the same word after the barrier
    const int nrow = min(kTile, nk - r0);
    if (tid < kTile) {
      const RPrep p = prep[keep_pos[r0 + min(tid, nrow - 1)]];
      srow[tid] = p; qx[0][tid] = p.cx; qy[0][tid] = p.cy; qr[0][tid] = p.reach;
      qw[0][tid] = p.w; qh[0][tid] = p.h; qc[0][tid] = p.c2; qs[0][tid] = p.s2; qa[0][tid] = p.area;
    }
    if (tid == 0) { qn = 0; qn2 = 0; }
    __syncthreads();
#pragma unroll 4
    for (int p = tid; p < kTile * kTile; p += 256) {
      const int i = p >> 6, j = p & 63;
      bool live = (i < nrow) && !((dead >> j) & 1ull);
      if (live && use_reject) {
        live = !rbox_far_soa(qx[0][i], qy[0][i], qw[0][i], qh[0][i], qr[0][i], qx[1][j], qy[1][j], qw[1][j], qh[1][j],
                             qr[1][j]);
        if (live && use_bounds)
          live = !rbox_cannot_exceed(qx[0][i], qy[0][i], qw[0][i], qh[0][i], qc[0][i], qs[0][i], qa[0][i], qx[1][j],
                                     qy[1][j], qw[1][j], qh[1][j], qc[1][j], qs[1][j], qa[1][j], thr);
      }
      const unsigned b = __ballot_sync(0xffffffffu, live);
      if (b) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&qn, __popc(b));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (live) queue[base + __popc(b & ((1u << lane) - 1))] = (unsigned short)p;
      }
    }
    __syncthreads();
    const int nq = qn;
    for (int q0 = 0; q0 < nq; q0 += 256) {
      const int q = q0 + tid;
      int verdict = 0, p = 0;
      if (q < nq) {
        p = queue[q];
        verdict = rbox_iou_exceeds_quick(srow[p >> 6], scol[p & 63], thr);
        if (verdict > 0) atomicOr(&s_dead, 1ull << (p & 63));
      }
      const unsigned b = __ballot_sync(0xffffffffu, verdict < 0);
      if (b) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&qn2, __popc(b));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (verdict < 0) queue2[base + __popc(b & ((1u << lane) - 1))] = (unsigned short)p;
      }
    }
    __syncthreads();
    const int nq2 = qn2;
    for (int q = tid; q < nq2; q += 256) {
      const int p = queue2[q];
      if (rbox_iou_full(srow[p >> 6], scol[p & 63]) > thr) atomicOr(&s_dead, 1ull << (p & 63));
    }
  }
  __syncthreads();
  if (tid == 0) rem[cb] = s_dead & (nvc == 64 ? ~0ull : ((1ull << nvc) - 1ull));
}

// One warp per image: the greedy scan of ONE band (chunks [c0, c0 + kBand)).  `rem` holds, for every chunk, the boxes
// suppressed by kept boxes of EARLIER bands (nms_band_suppress_kernel); suppression inside the band comes from the mask
// words nms_mask_kernel stored for the band's triangle.  Records the kept boxes (keep_pos, kept bits per chunk) and the running count.
__global__ void __launch_bounds__(32)
nms_band_scan_kernel(const unsigned long long* __restrict__ mask, int64_t mask_stride, int words,
                     const int32_t* __restrict__ counts, int max_det, int c0, int kBand,
                     const unsigned long long* __restrict__ rem, unsigned long long* __restrict__ kept,
                     BandState* __restrict__ state, int32_t* __restrict__ keep_pos, int64_t keep_stride) {
  __shared__ unsigned long long srem[kBandMax];
  const int img = blockIdx.x, lane = threadIdx.x;
  if (state[img].done) return;
  const int K = counts[img];
  mask += (int64_t)img * mask_stride;
  keep_pos += (int64_t)img * keep_stride;
  rem += (int64_t)img * words;
  kept += (int64_t)img * words;
  const int nchunk = (K + kTile - 1) / kTile;
  const int c1 = min(c0 + kBand, nchunk);
  if (lane < kBand) srem[lane] = (c0 + lane < words) ? rem[c0 + lane] : 0ull;
  __syncwarp();
  int nk = state[img].nk;
  for (int c = c0; c < c1 && nk < max_det; c++) {
    const int r0 = c * kTile;
    unsigned long long dlo = 0, dhi = 0;
    unsigned long long cur = srem[c - c0], kb = 0ull;
    // rows nms_mask_kernel skipped (already suppressed) have stale mask words: they are never read (never kept)
    if (r0 + lane < K && !((cur >> lane) & 1ull)) dlo = mask[(int64_t)(r0 + lane) * words + c];
    if (r0 + 32 + lane < K && !((cur >> (lane + 32)) & 1ull)) dhi = mask[(int64_t)(r0 + 32 + lane) * words + c];
    const int nb = min(kTile, K - r0);
    for (int b = 0; b < nb; b++) {
      unsigned long long d = __shfl_sync(0xffffffffu, (b < 32) ? dlo : dhi, b & 31);
      if (!((cur >> b) & 1ull)) { kb |= 1ull << b; cur |= d; }
    }
#pragma unroll
    for (int h = 0; h < 2; h++) {
      int b = lane + 32 * h;
      if ((kb >> b) & 1ull) {
        int pos = nk + __popcll(kb & ((1ull << b) - 1ull));
        if (pos < max_det) keep_pos[pos] = r0 + b;
      }
    }
    // boxes kept beyond max_det are never emitted, but they were kept: they keep suppressing (same as the full scan,
    // which stops right after this chunk anyway)
    nk += __popcll(kb);
    if (lane == 0) kept[c] = kb;
    // fold the kept rows into the later chunks of this band
    const int w = c + 1 + lane;
    if (w < c1) {
      unsigned long long v = 0ull, k2 = kb;
      while (k2) {
        const int b0 = __ffsll((long long)k2) - 1; k2 &= k2 - 1;
        v |= mask[(int64_t)(r0 + b0) * words + w];
      }
      srem[w - c0] |= v;
    }
    __syncwarp();
  }
  if (lane == 0) {
    state[img].nk = nk;
    if (nk >= max_det || c1 >= nchunk) state[img].done = 1;
  }
}

// tail of post_process (lib/general.py:178-181): first max_det kept boxes, gathered from the sorted rows
__global__ void __launch_bounds__(128)
nms_band_gather_kernel(const BandState* __restrict__ state, int max_det, const int32_t* __restrict__ keep_pos,
                       int64_t keep_stride, const float* __restrict__ dets_sorted,
                       const int32_t* __restrict__ rows_sorted, int64_t sorted_stride, float* __restrict__ dets_out,
                       int64_t* __restrict__ rows_out, int32_t* __restrict__ n_keep,
                       // plain nms API (one box set): order[] -> int64 indices into the caller's boxes
                       const int32_t* __restrict__ order, int64_t* __restrict__ keep_idx64) {
  const int img = blockIdx.x;
  const int nk = min(state[img].nk, max_det);
  keep_pos += (int64_t)img * keep_stride;
  if (keep_idx64) {
    if (threadIdx.x == 0) n_keep[img] = nk;
    for (int j = threadIdx.x; j < nk; j += blockDim.x) keep_idx64[j] = (int64_t)order[keep_pos[j]];
    return;
  }
  dets_sorted += (int64_t)img * sorted_stride * 7;
  rows_sorted += (int64_t)img * sorted_stride;
  dets_out += (int64_t)img * max_det * 7;
  rows_out += (int64_t)img * max_det;
  if (threadIdx.x == 0) n_keep[img] = nk;
  for (int e = threadIdx.x; e < nk * 7; e += blockDim.x) {
    const int j = e / 7, f = e - 7 * j;
    dets_out[e] = dets_sorted[(int64_t)keep_pos[j] * 7 + f];
  }
  for (int j = threadIdx.x; j < nk; j += blockDim.x) rows_out[j] = (int64_t)rows_sorted[keep_pos[j]];
}

// ------------------------------------------------------------------------------------ NMS scan
// One warp per image.  Walks the 64-box chunks in score order; inside a chunk the diagonal tile is
// resolved sequentially with warp shuffles, then the rows of the boxes that were kept are OR-ed into
// the running "removed" bit vector (shared memory) in parallel.
// Emits keep positions (indices into the sorted order), and optionally gathers output rows.
__global__ void __launch_bounds__(32)
nms_scan_kernel(const unsigned long long* __restrict__ mask, int64_t mask_stride, int words,
                const int32_t* __restrict__ counts, int max_det,
                int32_t* __restrict__ keep_pos, int64_t keep_stride, int32_t* __restrict__ n_keep,
                // optional gather (post_process): sorted det rows [.,7] and their source rows
                const float* __restrict__ dets_sorted, const int32_t* __restrict__ rows_sorted,
                int64_t sorted_stride, float* __restrict__ dets_out, int64_t* __restrict__ rows_out,
                // optional gather (plain nms API): order[] -> int64 indices into the caller's boxes
                const int32_t* __restrict__ order, int64_t* __restrict__ keep_idx64) {
  extern __shared__ unsigned long long rem[];
  const int img = blockIdx.x, lane = threadIdx.x;
  const int K = counts[img];
  mask += (int64_t)img * mask_stride;
  keep_pos += (int64_t)img * keep_stride;
  const int nchunk = (K + kTile - 1) / kTile;
  for (int w = lane; w < words; w += 32) rem[w] = 0ull;
  __syncwarp();
  int nk = 0;
  for (int c = 0; c < nchunk && nk < max_det; c++) {
    const int r0 = c * kTile;
    unsigned long long dlo = 0, dhi = 0;
    if (r0 + lane < K) dlo = mask[(int64_t)(r0 + lane) * words + c];
    if (r0 + 32 + lane < K) dhi = mask[(int64_t)(r0 + 32 + lane) * words + c];
    unsigned long long cur = rem[c], kept = 0ull;
    const int nb = min(kTile, K - r0);
    for (int b = 0; b < nb; b++) {
      unsigned long long d = __shfl_sync(0xffffffffu, (b < 32) ? dlo : dhi, b & 31);
      if (!((cur >> b) & 1ull)) { kept |= 1ull << b; cur |= d; }
    }
    // record keeps (lanes own bits lane and lane+32)
#pragma unroll
    for (int h = 0; h < 2; h++) {
      int b = lane + 32 * h;
      if ((kept >> b) & 1ull) {
        int pos = nk + __popcll(kept & ((1ull << b) - 1ull));
        if (pos < max_det) keep_pos[pos] = r0 + b;
      }
    }
    nk += __popcll(kept);
    // fold the kept rows into the removed vector for the chunks still to come
    unsigned long long k2 = kept;
    while (k2) {
      int b0 = __ffsll((long long)k2) - 1; k2 &= k2 - 1;
      int b1 = -1, b2 = -1, b3 = -1;
      if (k2) { b1 = __ffsll((long long)k2) - 1; k2 &= k2 - 1; }
      if (k2) { b2 = __ffsll((long long)k2) - 1; k2 &= k2 - 1; }
      if (k2) { b3 = __ffsll((long long)k2) - 1; k2 &= k2 - 1; }
      for (int w = c + 1 + lane; w < nchunk; w += 32) {
        unsigned long long v = mask[(int64_t)(r0 + b0) * words + w];
        if (b1 >= 0) v |= mask[(int64_t)(r0 + b1) * words + w];
        if (b2 >= 0) v |= mask[(int64_t)(r0 + b2) * words + w];
        if (b3 >= 0) v |= mask[(int64_t)(r0 + b3) * words + w];
        rem[w] |= v;
      }
    }
    __syncwarp();
  }
  nk = min(nk, max_det);
  if (lane == 0) n_keep[img] = nk;
  __syncwarp();
  if (dets_out) {
    dets_sorted += (int64_t)img * sorted_stride * 7;
    rows_sorted += (int64_t)img * sorted_stride;
    dets_out += (int64_t)img * max_det * 7;
    rows_out += (int64_t)img * max_det;
    for (int e = lane; e < nk * 7; e += 32) {
      int j = e / 7, f = e - 7 * j;
      dets_out[e] = dets_sorted[(int64_t)keep_pos[j] * 7 + f];
    }
    for (int j = lane; j < nk; j += 32) rows_out[j] = (int64_t)rows_sorted[keep_pos[j]];
  }
  if (keep_idx64) {
    for (int j = lane; j < nk; j += 32) keep_idx64[j] = (int64_t)order[keep_pos[j]];
  }
}

// ------------------------------------------------------------------------------------ sorting
// ascending bitonic sort of n (power of two) u64 keys living in shared memory
__device__ void bitonic_sort_smem(unsigned long long* s, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
        int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        int hi = lo | j;
        bool up = ((lo & k) == 0);
        unsigned long long a = s[lo], b = s[hi];
        if ((a > b) == up) { s[lo] = b; s[hi] = a; }
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// composite key: ascending order == (score descending, index ascending)
__device__ __forceinline__ unsigned long long score_key(float s, uint32_t idx) {
  return ((unsigned long long)(~ry_float_order(s)) << 32) | idx;
}

// Stable descending argsort of up to kSortCap scores by ONE CTA (plain nms API).
__global__ void __launch_bounds__(1024)
argsort_desc_smem_kernel(const float* __restrict__ scores, int n, int32_t* __restrict__ order,
                         int32_t* __restrict__ count_out) {
  extern __shared__ unsigned long long keys[];
  int np2 = next_pow2(n);
  for (int i = threadIdx.x; i < np2; i += blockDim.x)
    keys[i] = (i < n) ? score_key(scores[i], (uint32_t)i) : ~0ull;
  __syncthreads();
  bitonic_sort_smem(keys, np2);
  for (int i = threadIdx.x; i < n; i += blockDim.x) order[i] = (int32_t)(keys[i] & 0xffffffffu);
  if (threadIdx.x == 0) count_out[0] = n;
}

// Large-n fallback for the plain nms API: global-memory bitonic network (one launch per stage).
__global__ void keys_init_kernel(const float* __restrict__ scores, int64_t n, int64_t np2,
                                 unsigned long long* __restrict__ keys) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < np2) keys[i] = (i < n) ? score_key(scores[i], (uint32_t)i) : ~0ull;
}
__global__ void bitonic_global_step_kernel(unsigned long long* __restrict__ keys, int64_t half, int64_t j, int64_t k) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= half) return;
  int64_t lo = ((t & ~(j - 1)) << 1) | (t & (j - 1)), hi = lo | j;
  bool up = ((lo & k) == 0);
  unsigned long long a = keys[lo], b = keys[hi];
  if ((a > b) == up) { keys[lo] = b; keys[hi] = a; }
}
__global__ void keys_to_order_kernel(const unsigned long long* __restrict__ keys, int64_t n,
                                     int32_t* __restrict__ order, int32_t* __restrict__ count_out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) order[i] = (int32_t)(keys[i] & 0xffffffffu);
  if (i == 0) count_out[0] = (int32_t)n;
}

// ------------------------------------------------------------------------------------ post_process
// lib/general.py:155-161.  One thread per prediction row, rows streamed with 128-bit loads when the
// row pitch allows.  Writes the per-row best score and class; optionally writes cls*=obj back.
template <int NC_STATIC>
__global__ void __launch_bounds__(256)
pp_score_kernel(float* __restrict__ pred, int64_t rows, int nc, int mutate, float* __restrict__ score,
                uint8_t* __restrict__ cls) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int n = NC_STATIC > 0 ? NC_STATIC : nc;
  float* p = pred + r * (6 + n);
  float best = 0.f;
  int bi = 0;
  if (NC_STATIC == 2) {  // 8 floats / row = two aligned float4
    float4 lo = ry_ld_stream(reinterpret_cast<const float4*>(p) + 1);  // (theta, obj, c0, c1)
    float c0 = __fmul_rn(lo.z, lo.y), c1 = __fmul_rn(lo.w, lo.y);
    best = c0; bi = 0;
    if (c1 > best) { best = c1; bi = 1; }
    if (mutate) *reinterpret_cast<float2*>(p + 6) = make_float2(c0, c1);
  } else {
    const float obj = p[5];
    for (int c = 0; c < n; c++) {
      float v = __fmul_rn(p[6 + c], obj);
      if (mutate) p[6 + c] = v;
      if (c == 0 || v > best) { best = v; bi = c; }
    }
  }
  score[r] = best;
  cls[r] = (uint8_t)bi;
}

// lib/general.py:161-174 for one image per CTA (1024 threads, 64 KiB dynamic smem for the sort).
__global__ void __launch_bounds__(1024)
pp_select_sort_kernel(const float* __restrict__ pred, const float* __restrict__ score,
                      const uint8_t* __restrict__ cls, int64_t R, int nc, float conf_thres, int max_nms,
                      float max_wh, float* __restrict__ dets_sorted, int32_t* __restrict__ rows_sorted,
                      RPrep* __restrict__ prep, int32_t* __restrict__ counts) {
  extern __shared__ unsigned long long keys[];   // kSortCap entries
  __shared__ int hist[2048];
  __shared__ int s_cnt, s_out, s_eq_seen;
  __shared__ uint32_t s_prefix;
  __shared__ int s_remaining, s_bin_total;
  const int img = blockIdx.x, tid = threadIdx.x;
  score += (int64_t)img * R;
  cls += (int64_t)img * R;
  pred += (int64_t)img * R * (6 + nc);

  // -- count candidates
  if (tid == 0) { s_cnt = 0; s_out = 0; s_eq_seen = 0; }
  __syncthreads();
  int local = 0;
  for (int64_t i = tid; i < R; i += 1024) local += (score[i] > conf_thres) ? 1 : 0;
  local = __reduce_add_sync(0xffffffffu, local);
  if ((tid & 31) == 0) atomicAdd(&s_cnt, local);
  __syncthreads();
  const int cnt = s_cnt;
  const int K = min(cnt, max_nms);
  if (K == 0) { if (tid == 0) counts[img] = 0; return; }

  // -- radix select: threshold key T such that exactly K composite keys are <= (T, row)
  uint32_t T = 0xffffffffu;
  int take_eq = 0x7fffffff;
  bool eq_all = true;
  if (cnt > max_nms) {
    uint32_t prefix = 0, pmask = 0;
    int remaining = K;
    const int shifts[3] = {21, 10, 0};
    const int bits[3] = {11, 11, 10};
    for (int pass = 0; pass < 3; pass++) {
      for (int i = tid; i < 2048; i += 1024) hist[i] = 0;
      __syncthreads();
      const int sh = shifts[pass];
      const uint32_t dm = (1u << bits[pass]) - 1u;
      for (int64_t i = tid; i < R; i += 1024) {
        float s = score[i];
        if (s > conf_thres) {
          uint32_t k = ~ry_float_order(s);
          if ((k & pmask) == prefix) atomicAdd(&hist[(k >> sh) & dm], 1);
        }
      }
      __syncthreads();
      if (tid == 0) {
        int cum = 0, d = 0;
        const int nb = 1 << bits[pass];
        for (d = 0; d < nb; d++) {
          if (cum + hist[d] >= remaining) break;
          cum += hist[d];
        }
        d = min(d, nb - 1);
        s_prefix = prefix | ((uint32_t)d << sh);
        s_remaining = remaining - cum;
        s_bin_total = hist[d];
      }
      __syncthreads();
      prefix = s_prefix;
      remaining = s_remaining;
      pmask |= dm << sh;
      __syncthreads();
    }
    T = prefix;
    take_eq = remaining;
    eq_all = (s_bin_total == remaining);
  }

  // -- compaction of the K selected (key,row) pairs into shared memory
  if (eq_all) {
    for (int64_t i = tid; i < R; i += 1024) {
      float s = score[i];
      if (s > conf_thres) {
        uint32_t k = ~ry_float_order(s);
        if (k <= T) keys[atomicAdd(&s_out, 1)] = ((unsigned long long)k << 32) | (uint32_t)i;
      }
    }
  } else {
    // ties straddle the cut: equal-score rows are admitted in row order (stable sort semantics)
    __shared__ int warp_tot[32];
    for (int64_t base = 0; base < R; base += 1024) {
      int64_t i = base + tid;
      bool cand = false, eq = false;
      uint32_t k = 0;
      if (i < R) {
        float s = score[i];
        if (s > conf_thres) { k = ~ry_float_order(s); cand = (k < T); eq = (k == T); }
      }
      unsigned be = __ballot_sync(0xffffffffu, eq);
      if ((tid & 31) == 0) warp_tot[tid >> 5] = __popc(be);
      __syncthreads();
      int before = s_eq_seen;
      for (int w = 0; w < (tid >> 5); w++) before += warp_tot[w];
      int rank = before + __popc(be & ((1u << (tid & 31)) - 1u));
      if (cand || (eq && rank < take_eq)) keys[atomicAdd(&s_out, 1)] = ((unsigned long long)k << 32) | (uint32_t)i;
      __syncthreads();
      if (tid == 0) { int t = 0; for (int w = 0; w < 32; w++) t += warp_tot[w]; s_eq_seen += t; }
      __syncthreads();
    }
  }
  __syncthreads();
  const int np2 = next_pow2(K);
  for (int i = K + tid; i < np2; i += 1024) keys[i] = ~0ull;
  __syncthreads();
  bitonic_sort_smem(keys, np2);

  // -- gather sorted rows, build NMS boxes (class offset, rad->deg) : lib/general.py:159,171-174
  dets_sorted += (int64_t)img * max_nms * 7;
  rows_sorted += (int64_t)img * max_nms;
  prep += (int64_t)img * max_nms;
  for (int k = tid; k < K; k += 1024) {
    const uint32_t row = (uint32_t)(keys[k] & 0xffffffffu);
    const float* p = pred + (int64_t)row * (6 + nc);
    const float x = p[0], y = p[1], w = p[2], h = p[3], th = p[4];
    const float sc = score[row], cf = (float)cls[row];
    float* d = dets_sorted + (int64_t)k * 7;
    d[0] = x; d[1] = y; d[2] = w; d[3] = h; d[4] = th; d[5] = sc; d[6] = cf;
    rows_sorted[k] = (int32_t)row;
    const float off = __fmul_rn(cf, max_wh);
    prep[k] = rprep(__fadd_rn(x, off), __fadd_rn(y, off), w, h, __fmul_rn(__fdiv_rn(th, kPiF), 180.f));
  }
  if (tid == 0) counts[img] = K;
}

}  // namespace

// ======================================================================================= C ABI
extern "C" {

size_t ryolo_pairwise_iou_rotated_workspace(int64_t n, int64_t m) { return (size_t)(n + m) * sizeof(RPrep) + 256; }

int ryolo_pairwise_iou_rotated(const float* a, int64_t n, const float* b, int64_t m, float* out,
                               void* workspace, size_t ws_bytes, void* stream) {
  RY_CHECK_ARG(n >= 0 && m >= 0, "pairwise_iou_rotated: negative size");
  if (n == 0 || m == 0) return RYOLO_OK;
  RY_CHECK_ARG(ws_bytes >= ryolo_pairwise_iou_rotated_workspace(n, m), "pairwise_iou_rotated: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  RPrep* pa = (RPrep*)workspace;
  RPrep* pb = pa + n;
  prep_boxes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, nullptr, n, pa);
  prep_boxes_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(b, nullptr, m, pb);
  dim3 grid((unsigned)((m + 15) / 16), (unsigned)((n + 15) / 16));
  pairwise_iou_kernel<<<grid, dim3(16, 16), 0, st>>>(pa, n, pb, m, out);
  RY_CHECK_LAUNCH();
  return RYOLO_OK;
}

// test.py:100-149 get_batch_statistics for a whole batch in one launch.
//   dets [B, max_det, 7] (x, y, w, h, theta RAD, score, class) with n_det[B] valid rows each (post_process_device's
//   outputs); targets [T, tcols>=7] (image, class, x, y, w, h, theta RAD) in any image order; iouv [niou] ascending.
//   tp uint8 [B, max_det, niou]; workspace: ryolo_eval_match_workspace(B, max_det) bytes; status int32[1] (zeroed by
//   the caller) becomes 1 if an image has more than 1024 targets (result then invalid).
size_t ryolo_eval_match_workspace(int64_t B, int max_det) { return (size_t)B * max_det * 8 + 256; }

int ryolo_eval_match(const float* dets, const int32_t* n_det, int64_t B, int max_det, const float* targets, int64_t T,
                     int tcols, const float* iouv, int niou, uint8_t* tp, int32_t* status, void* workspace,
                     size_t ws_bytes, void* stream) {
  RY_CHECK_ARG(B >= 0 && max_det > 0 && T >= 0 && tcols >= 7 && niou > 0 && niou <= 32, "eval_match: bad arguments");
  RY_CHECK_ARG(ws_bytes >= ryolo_eval_match_workspace(B, max_det), "eval_match: workspace too small");
  if (B == 0) return RYOLO_OK;
  float* best_iou = (float*)workspace;
  int32_t* best_tgt = (int32_t*)(best_iou + (size_t)B * max_det);
  eval_match_kernel<<<(unsigned)B, 256, 0, (cudaStream_t)stream>>>(dets, n_det, max_det, targets, T, tcols, iouv, niou,
                                                                  tp, best_iou, best_tgt, status);
  RY_CHECK_LAUNCH();
  return RYOLO_OK;
}

static int64_t pow2_ge(int64_t v) { int64_t p = 1; while (p < v) p <<= 1; return p; }

size_t ryolo_nms_rotated_workspace(int64_t n) {
  int64_t words = (n + 63) / 64;
  size_t s = 0;
  s += ry_align_up((size_t)n * sizeof(RPrep), 256);                  // prep
  s += ry_align_up((size_t)n * words * 8, 256);                      // mask
  s += ry_align_up((size_t)n * 4, 256) * 2;                          // order, keep_pos
  s += ry_align_up((size_t)pow2_ge(n > 0 ? n : 1) * 8, 256);         // keys (large-n sort)
  s += 256;                                                          // count
  s += ry_align_up((size_t)words * 16 + sizeof(BandState), 256);     // banded scan: rem | kept | state
  return s;
}

// detectron2.layers.nms.nms_rotated drop-in (reference call site lib/general.py:177).
// boxes5: [n,5] (cx,cy,w,h,deg), scores [n]; keep: int64[n] device; n_keep: int32[1] device.
int ryolo_nms_rotated(const float* boxes5, const float* scores, int64_t n, float iou_thr, int64_t* keep,
                      int32_t* n_keep, void* workspace, size_t ws_bytes, void* stream) {
  RY_CHECK_ARG(n >= 0 && n < (1ll << 31), "nms_rotated: bad n");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) { cudaMemsetAsync(n_keep, 0, 4, st); return RYOLO_OK; }
  RY_CHECK_ARG(ws_bytes >= ryolo_nms_rotated_workspace(n), "nms_rotated: workspace too small");
  const int words = (int)((n + 63) / 64);
  char* w = (char*)workspace;
  RPrep* prep = (RPrep*)w; w += ry_align_up((size_t)n * sizeof(RPrep), 256);
  unsigned long long* mask = (unsigned long long*)w; w += ry_align_up((size_t)n * words * 8, 256);
  int32_t* order = (int32_t*)w; w += ry_align_up((size_t)n * 4, 256);
  int32_t* keep_pos = (int32_t*)w; w += ry_align_up((size_t)n * 4, 256);
  unsigned long long* keys = (unsigned long long*)w; w += ry_align_up((size_t)pow2_ge(n) * 8, 256);
  int32_t* count = (int32_t*)w; w += 256;
  unsigned long long* rem = (unsigned long long*)w;

  if (n <= kSortCap) {
    size_t sm = (size_t)pow2_ge(n) * 8;
    cudaFuncSetAttribute(argsort_desc_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortCap * 8);
    argsort_desc_smem_kernel<<<1, 1024, sm, st>>>(scores, (int)n, order, count);
  } else {
    int64_t np2 = pow2_ge(n), half = np2 / 2;
    keys_init_kernel<<<(unsigned)((np2 + 255) / 256), 256, 0, st>>>(scores, n, np2, keys);
    for (int64_t k = 2; k <= np2; k <<= 1)
      for (int64_t j = k >> 1; j > 0; j >>= 1)
        bitonic_global_step_kernel<<<(unsigned)((half + 255) / 256), 256, 0, st>>>(keys, half, j, k);
    keys_to_order_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(keys, n, order, count);
  }
  prep_boxes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(boxes5, order, n, prep);
  int kBand = ryolo_knob(RYOLO_KNOB_NMS_BAND);
  if (kBand > kBandMax) kBand = kBandMax;
  if (kBand > 0) {                                  // banded greedy NMS, as in ryolo_post_process (no max_det cap here)
    unsigned long long* keptb = rem + words;
    BandState* state = (BandState*)(keptb + words);
    cudaMemsetAsync(rem, 0, (size_t)words * 16 + sizeof(BandState), st);
    for (int c0 = 0; c0 < words; c0 += kBand) {
      if (c0 > 0)
        nms_band_suppress_kernel<<<dim3(kBand, 1), 256, 0, st>>>(prep, count, 0, words, iou_thr, c0, rem, state, keep_pos, 0);
      nms_mask_kernel<<<dim3(kBand, kBand, 1), 256, 0, st>>>(prep, count, 0, words, iou_thr, mask, 0, c0, c0, rem, state);
      nms_band_scan_kernel<<<1, 32, 0, st>>>(mask, 0, words, count, (int)n, c0, kBand, rem, keptb, state, keep_pos, 0);
    }
    nms_band_gather_kernel<<<1, 128, 0, st>>>(state, (int)n, keep_pos, 0, nullptr, nullptr, 0, nullptr, nullptr, n_keep,
                                             order, keep);
    RY_CHECK_LAUNCH();
    return RYOLO_OK;
  }
  dim3 grid(words, words, 1);
  nms_mask_kernel<<<grid, 256, 0, st>>>(prep, count, 0, words, iou_thr, mask, 0, 0, 0, nullptr, nullptr);
  size_t scan_smem = (size_t)words * 8;
  if (scan_smem > 48 * 1024)
    cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_smem);
  nms_scan_kernel<<<1, 32, scan_smem, st>>>(mask, 0, words, count, (int)n, keep_pos, 0, n_keep, nullptr, nullptr, 0,
                                            nullptr, nullptr, order, keep);
  RY_CHECK_LAUNCH();
  return RYOLO_OK;
}

static void pp_layout(int64_t B, int64_t R, int max_nms, size_t* off, size_t* total) {   // off[9]
  const int words = (max_nms + 63) / 64;
  size_t s = 0;
  off[0] = s; s += ry_align_up((size_t)B * R * 4, 256);                          // score
  off[1] = s; s += ry_align_up((size_t)B * R, 256);                              // cls
  off[2] = s; s += ry_align_up((size_t)B * max_nms * 7 * 4, 256);                // dets_sorted
  off[3] = s; s += ry_align_up((size_t)B * max_nms * 4, 256);                    // rows_sorted
  off[4] = s; s += ry_align_up((size_t)B * max_nms * sizeof(RPrep), 256);        // prep
  off[5] = s; s += ry_align_up((size_t)B * 4, 256);                              // counts (K per image)
  off[6] = s; s += ry_align_up((size_t)B * max_nms * words * 8, 256);            // mask
  off[7] = s; s += ry_align_up((size_t)B * max_nms * 4, 256);                    // keep_pos
  off[8] = s; s += ry_align_up((size_t)B * words * 8 * 2 + (size_t)B * sizeof(BandState), 256);   // rem | kept | state
  *total = s;
}

size_t ryolo_post_process_workspace(int64_t B, int64_t R, int nc, int max_nms) {
  (void)nc;
  size_t off[9], total;
  pp_layout(B, R, max_nms, off, &total);
  return total;
}

// post_process drop-in (lib/general.py:136-183), whole batch, no host sync.
//   pred      [B,R,6+nc] fp32, device; class columns are multiplied by objectness IN PLACE when
//             mutate != 0 (the reference does, :155)
//   dets_out  [B,max_det,7] fp32 (x,y,w,h,theta,score,cls) score-descending
//   rows_out  [B,max_det]   int64 source row of every detection (for index-exact parity checks)
//   n_out     [B]           int32 detections per image
int ryolo_post_process(float* pred, int64_t B, int64_t R, int nc, float conf_thres, float iou_thres,
                       int max_nms, int max_det, float max_wh, int mutate, float* dets_out,
                       int64_t* rows_out, int32_t* n_out, void* workspace, size_t ws_bytes, void* stream) {
  RY_CHECK_ARG(B >= 0 && R >= 0 && nc >= 1 && nc <= 255, "post_process: bad shape");
  RY_CHECK_ARG(max_nms >= 1 && max_nms <= kSortCap, "post_process: max_nms must be in [1, 8192]");
  RY_CHECK_ARG(R < (1ll << 31), "post_process: too many rows per image");
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) return RYOLO_OK;
  if (R == 0) { cudaMemsetAsync(n_out, 0, (size_t)B * 4, st); return RYOLO_OK; }
  size_t off[9], total;
  pp_layout(B, R, max_nms, off, &total);
  RY_CHECK_ARG(ws_bytes >= total, "post_process: workspace too small");
  char* w = (char*)workspace;
  float* score = (float*)(w + off[0]);
  uint8_t* cls = (uint8_t*)(w + off[1]);
  float* dets_sorted = (float*)(w + off[2]);
  int32_t* rows_sorted = (int32_t*)(w + off[3]);
  RPrep* prep = (RPrep*)(w + off[4]);
  int32_t* counts = (int32_t*)(w + off[5]);
  unsigned long long* mask = (unsigned long long*)(w + off[6]);
  int32_t* keep_pos = (int32_t*)(w + off[7]);
  const int words = (max_nms + 63) / 64;

  const int64_t rows = B * R;
  const unsigned g = (unsigned)((rows + 255) / 256);
  if (nc == 2 && (((uintptr_t)pred) & 15) == 0)
    pp_score_kernel<2><<<g, 256, 0, st>>>(pred, rows, nc, mutate, score, cls);
  else
    pp_score_kernel<0><<<g, 256, 0, st>>>(pred, rows, nc, mutate, score, cls);
  cudaFuncSetAttribute(pp_select_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortCap * 8);
  pp_select_sort_kernel<<<(unsigned)B, 1024, kSortCap * 8, st>>>(pred, score, cls, R, nc, conf_thres, max_nms, max_wh,
                                                               dets_sorted, rows_sorted, prep, counts);
  int kBand = ryolo_knob(RYOLO_KNOB_NMS_BAND);
  if (kBand > kBandMax) kBand = kBandMax;
  if (kBand <= 0) {           // the full N^2 / 2 mask, then one scan (round-1 path, A/B switch)
    dim3 grid(words, words, (unsigned)B);
    nms_mask_kernel<<<grid, 256, 0, st>>>(prep, counts, max_nms, words, iou_thres, mask, (int64_t)max_nms * words, 0, 0,
                                          nullptr, nullptr);
    nms_scan_kernel<<<(unsigned)B, 32, (size_t)words * 8, st>>>(mask, (int64_t)max_nms * words, words, counts, max_det,
                                                               keep_pos, max_nms, n_out, dets_sorted, rows_sorted,
                                                               max_nms, dets_out, rows_out, nullptr, nullptr);
    RY_CHECK_LAUNCH();
    return RYOLO_OK;
  }
  // Banded greedy NMS (see nms_mask_kernel): per band of kBand tiles  kept-so-far rows against the band's columns ->
  // triangle mask of the band -> scan.  3 launches per band, no host synchronisation (the band count comes from max_nms).
  unsigned long long* rem = (unsigned long long*)(w + off[8]);
  unsigned long long* keptb = rem + (size_t)B * words;
  BandState* state = (BandState*)(keptb + (size_t)B * words);
  cudaMemsetAsync(rem, 0, (size_t)B * words * 8 * 2 + (size_t)B * sizeof(BandState), st);
  const int64_t mstride = (int64_t)max_nms * words;
  for (int c0 = 0; c0 < words; c0 += kBand) {
    if (c0 > 0)          // every box kept so far against this band's columns: fills rem[c0 .. c0 + kBand)
      nms_band_suppress_kernel<<<dim3(kBand, (unsigned)B), 256, 0, st>>>(prep, counts, max_nms, words, iou_thres, c0, rem,
                                                                        state, keep_pos, max_nms);
    nms_mask_kernel<<<dim3(kBand, kBand, (unsigned)B), 256, 0, st>>>(prep, counts, max_nms, words, iou_thres, mask, mstride,
                                                                    c0, c0, rem, state);
    nms_band_scan_kernel<<<(unsigned)B, 32, 0, st>>>(mask, mstride, words, counts, max_det, c0, kBand, rem, keptb, state,
                                                    keep_pos, max_nms);
  }
  nms_band_gather_kernel<<<(unsigned)B, 128, 0, st>>>(state, max_det, keep_pos, max_nms, dets_sorted, rows_sorted, max_nms,
                                                     dets_out, rows_out, n_out, nullptr, nullptr);
  RY_CHECK_LAUNCH();
  return RYOLO_OK;
}

}  // extern "C"
