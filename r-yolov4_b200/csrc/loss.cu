// Target assignment + CSL / KFIoU loss, forward and backward in one pass  (K7, K8, K9 in SURVEY.md §2.1).
//
// Replaces ComputeCSLLoss.__call__ / build_targets (lib/loss.py:191-331) and
// ComputeKFIoULoss.__call__ / build_targets (lib/loss.py:368-492):
//   assign_{count,scan,emit}  ordered compaction of the [5, na, T] candidate grid (reference order)
//   csl_pos_kernel            one warp per positive: gather row, CIoU, class BCE, 180-bin CSL BCE,
//                             gradients scattered with atomics (duplicates accumulate like index_put's
//                             backward), objectness target written with a (order,value) atomicMax so
//                             that the LAST positive in reference order wins
//   kfiou_pos_kernel          one thread per positive: O(N) KFIoU (the reference's N x N broadcast,
//                             lib/loss.py:148, is reproduced in value only)
//   obj_dense_kernel          objectness BCE over every cell + its gradient
//   loss_finalize_kernel      gains, per-level means, loss_items; no host synchronisation anywhere
// Compile with --fmad=false.
#include "common.cuh"
#include "assign.cuh"

using namespace ryolo;

namespace {

constexpr int kAssignBlock = 1024;

struct LevelArgs {
  const float* level[3];
  float* grad[3];
  unsigned long long* tconf[3];
  Pos* pos[3];
  int gh[3], gw[3];
  long long cells[3];
};

// ------------------------------------------------------------------------------------------ assign
__global__ void __launch_bounds__(kAssignBlock)
assign_count_kernel(const float* __restrict__ targets, int T, int tcols, const float* __restrict__ anchors,
                    int na, int rotated, int nimg, LevelArgs L, int nblk, int* __restrict__ blockcnt) {
  const int lvl = blockIdx.y;
  const long long E = 5ll * na * T;
  const long long e = blockIdx.x * (long long)kAssignBlock + threadIdx.x;
  bool in = false;
  if (e < E) {
    const int t = (int)(e % T), a = (int)((e / T) % na), o = (int)(e / ((long long)T * na));
    in = assign_entry(targets + (long long)t * tcols, o, a, anchors + (lvl * na + a) * 3, rotated, na, L.gh[lvl],
                      L.gw[lvl], t, nimg, nullptr);
  }
  const int c = __syncthreads_count(in);
  if (threadIdx.x == 0) blockcnt[lvl * nblk + blockIdx.x] = c;
}

__global__ void __launch_bounds__(1024)
assign_scan_kernel(const int* __restrict__ blockcnt, int nblk, int* __restrict__ blockoff, int* __restrict__ counts) {
  __shared__ int wsum[32];
  __shared__ int carry;
  const int lvl = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblk; base += 1024) {
    const int i = base + tid;
    const int v = (i < nblk) ? blockcnt[lvl * nblk + i] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int n = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += n;
    }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      int w = wsum[lane], winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += n;
      }
      wsum[lane] = winc - w;   // exclusive warp offsets
    }
    __syncthreads();
    const int excl = carry + wsum[wid] + inc - v;
    if (i < nblk) blockoff[lvl * nblk + i] = excl;
    __syncthreads();
    if (tid == 1023) carry = excl + v;
    __syncthreads();
  }
  if (tid == 0) counts[lvl] = carry;
}

__global__ void __launch_bounds__(kAssignBlock)
assign_emit_kernel(const float* __restrict__ targets, int T, int tcols, const float* __restrict__ anchors,
                   int na, int rotated, int nimg, LevelArgs L, int nblk, const int* __restrict__ blockoff) {
  __shared__ int wsum[32];
  const int lvl = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const long long E = 5ll * na * T;
  const long long e = blockIdx.x * (long long)kAssignBlock + tid;
  bool in = false;
  Pos rec;
  if (e < E) {
    const int t = (int)(e % T), a = (int)((e / T) % na), o = (int)(e / ((long long)T * na));
    in = assign_entry(targets + (long long)t * tcols, o, a, anchors + (lvl * na + a) * 3, rotated, na, L.gh[lvl],
                      L.gw[lvl], t, nimg, &rec);
  }
  const unsigned bal = __ballot_sync(0xffffffffu, in);
  if (lane == 0) wsum[wid] = __popc(bal);
  __syncthreads();
  if (wid == 0) {
    int w = wsum[lane], winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int n = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += n;
    }
    wsum[lane] = winc - w;
  }
  __syncthreads();
  if (in) {
    const int dst = blockoff[lvl * nblk + blockIdx.x] + wsum[wid] + __popc(bal & ((1u << lane) - 1u));
    L.pos[lvl][dst] = rec;
  }
}

// ------------------------------------------------------------------------------------------ losses
struct Hyp {
  float box, obj, cls, obj_pw, cls_pw, gamma, theta;
};

__device__ __forceinline__ unsigned long long tconf_key(int order, float v) {
  return ((unsigned long long)(unsigned)order << 32) | (unsigned long long)__float_as_uint(fmaxf(v, 0.f));
}

// sums layout per level: [0] reg, [1] cls, [2] theta, [3] obj
template <bool GRAD>
__global__ void __launch_bounds__(256)
csl_pos_kernel(LevelArgs L, int lvl, const int* __restrict__ counts, const float* __restrict__ targets, int tcols,
               const float* __restrict__ anchors, int na, int nc, Hyp hyp, double* __restrict__ sums) {
  const int n = counts[lvl];
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const int ch = nc + 185;
  const float* __restrict__ level = L.level[lvl];
  float* __restrict__ grad = L.grad[lvl];
  const Pos* __restrict__ pos = L.pos[lvl];
  unsigned long long* __restrict__ tconf = L.tconf[lvl];
  const float inv_n = n > 0 ? 1.f / (float)n : 0.f;
  const float g_box = hyp.box * inv_n, g_cls = hyp.cls * inv_n / (float)nc, g_th = hyp.theta * inv_n / 180.f;
  double a_reg = 0.0, a_cls = 0.0, a_th = 0.0;
  for (int i = warp; i < n; i += nwarps) {
    const Pos r = pos[i];
    const float* p = level + (long long)r.cell * ch;
    float* g = GRAD ? grad + (long long)r.cell * ch : nullptr;
    if (lane == 0) {
      const float* an = anchors + (lvl * na + r.a) * 3;
      const float sx = ry_sigmoid(p[0]), sy = ry_sigmoid(p[1]), sw = ry_sigmoid(p[2]), sh = ry_sigmoid(p[3]);
      Box4 pb = {sx * 2.f - 0.5f, sy * 2.f - 0.5f, (sw * 2.f) * (sw * 2.f) * an[0], (sh * 2.f) * (sh * 2.f) * an[1]};
      Box4 tb = {r.bx, r.by, r.bw, r.bh};
      Box4 dc;
      const float c = ciou_fwd_bwd(pb, tb, GRAD ? &dc : nullptr);
      a_reg += (double)(1.f - c);
      atomicMax(&tconf[r.cell], tconf_key(i, c));
      if (GRAD) {
        atomicAdd(g + 0, -g_box * dc.x * 2.f * sx * (1.f - sx));
        atomicAdd(g + 1, -g_box * dc.y * 2.f * sy * (1.f - sy));
        atomicAdd(g + 2, -g_box * dc.w * 8.f * sw * an[0] * sw * (1.f - sw));
        atomicAdd(g + 3, -g_box * dc.h * 8.f * sh * an[1] * sh * (1.f - sh));
      }
    }
    float lsum = 0.f;
    if (nc > 1) {
      for (int c = lane; c < nc; c += 32) {
        float l, dx;
        bce_logits(p[5 + c], c == r.cls ? 1.f : 0.f, hyp.cls_pw, hyp.gamma, &l, &dx);
        lsum += l;
        if (GRAD) atomicAdd(g + 5 + c, g_cls * dx);
      }
      a_cls += (double)lsum;
    }
    lsum = 0.f;
    const float* tg = targets + (long long)r.row * tcols + 7;
    for (int j = lane; j < 180; j += 32) {
      float l, dx;
      bce_logits(p[5 + nc + j], tg[j], 1.0f, hyp.gamma, &l, &dx);
      lsum += l;
      if (GRAD) atomicAdd(g + 5 + nc + j, g_th * dx);
    }
    a_th += (double)lsum;
  }
  // one set of atomics per BLOCK: ~9500 warps adding three doubles each to the same three addresses serialise in L2 and
  // were most of this kernel's time (lane 0 holds a_reg, every lane a share of a_cls / a_th)
  __shared__ double spart[3][8];
  a_reg = ry_warp_sum_d(a_reg);
  a_cls = ry_warp_sum_d(a_cls);
  a_th = ry_warp_sum_d(a_th);
  const int wib = threadIdx.x >> 5;
  if (lane == 0) { spart[0][wib] = a_reg; spart[1][wib] = a_cls; spart[2][wib] = a_th; }
  __syncthreads();
  if (threadIdx.x < 3 && n > 0 && (int)(blockIdx.x * (blockDim.x >> 5)) < n) {     // blocks without a positive add nothing
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += spart[threadIdx.x][w];
    atomicAdd(&sums[lvl * 4 + threadIdx.x], t);
  }
}

template <bool GRAD>
__global__ void __launch_bounds__(128)
kfiou_pos_kernel(LevelArgs L, int lvl, const int* __restrict__ counts, const float* __restrict__ anchors, int na,
                 int nc, Hyp hyp, double* __restrict__ sums) {
  const int n = counts[lvl];
  const int lane = threadIdx.x & 31;
  const int ch = nc + 6;
  const float* __restrict__ level = L.level[lvl];
  float* __restrict__ grad = L.grad[lvl];
  const Pos* __restrict__ pos = L.pos[lvl];
  unsigned long long* __restrict__ tconf = L.tconf[lvl];
  const float inv_n = n > 0 ? 1.f / (float)n : 0.f;
  const float g_box = hyp.box * inv_n, g_cls = hyp.cls * inv_n / (float)nc;
  double a_reg = 0.0, a_cls = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const Pos r = pos[i];
    const float* p = level + (long long)r.cell * ch;
    float* g = GRAD ? grad + (long long)r.cell * ch : nullptr;
    const float* an = anchors + (lvl * na + r.a) * 3;
    const float sx = ry_sigmoid(p[0]), sy = ry_sigmoid(p[1]), sw = ry_sigmoid(p[2]), sh = ry_sigmoid(p[3]);
    const float sa = ry_sigmoid(p[4]);
    Box5 pb = {sx * 2.f - 0.5f, sy * 2.f - 0.5f, (sw * 2.f) * (sw * 2.f) * an[0], (sh * 2.f) * (sh * 2.f) * an[1],
               norm_angle1((sa - 0.5f) * 1.1f + an[2])};                                  // lib/loss.py:388-390
    Box5 tb = {r.bx, r.by, r.bw, r.bh, r.ang};
    Box5 dg;
    float xl, kl, kfiou;
    kf_fwd_bwd(pb, tb, &xl, &kl, &kfiou, GRAD ? &dg : nullptr);
    a_reg += (double)xl + (double)kl;
    atomicMax(&tconf[r.cell], tconf_key(i, kfiou));
    if (GRAD) {
      atomicAdd(g + 0, g_box * dg.x * 2.f * sx * (1.f - sx));
      atomicAdd(g + 1, g_box * dg.y * 2.f * sy * (1.f - sy));
      atomicAdd(g + 2, g_box * dg.w * 8.f * sw * an[0] * sw * (1.f - sw));
      atomicAdd(g + 3, g_box * dg.h * 8.f * sh * an[1] * sh * (1.f - sh));
      atomicAdd(g + 4, g_box * dg.r * 1.1f * sa * (1.f - sa));
    }
    if (nc > 1) {
      float lsum = 0.f;
      for (int c = 0; c < nc; c++) {
        float l, dx;
        bce_logits(p[6 + c], c == r.cls ? 1.f : 0.f, hyp.cls_pw, hyp.gamma, &l, &dx);
        lsum += l;
        if (GRAD) atomicAdd(g + 6 + c, g_cls * dx);
      }
      a_cls += (double)lsum;
    }
  }
  __shared__ double spart[2][4];                 // one pair of atomics per block (see csl_pos_kernel)
  a_reg = ry_warp_sum_d(a_reg);
  a_cls = ry_warp_sum_d(a_cls);
  const int wib = threadIdx.x >> 5;
  if (lane == 0) { spart[0][wib] = a_reg; spart[1][wib] = a_cls; }
  __syncthreads();
  if (threadIdx.x < 2 && n > 0 && (int)(blockIdx.x * blockDim.x) < n) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += spart[threadIdx.x][w];
    atomicAdd(&sums[lvl * 4 + threadIdx.x], t);
  }
}

// objectness BCE over every cell of one level (lib/loss.py:248 / :407) + its gradient.
template <bool GRAD>
__global__ void __launch_bounds__(256)
obj_dense_kernel(LevelArgs L, int lvl, int ch, int oc, Hyp hyp, double* __restrict__ sums) {
  __shared__ double wpart[8];
  const long long C = L.cells[lvl];
  const float* __restrict__ level = L.level[lvl];
  float* __restrict__ grad = L.grad[lvl];
  const unsigned long long* __restrict__ tconf = L.tconf[lvl];
  const float gscale = hyp.obj / (float)C;
  double acc = 0.0;
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < C; c += (long long)gridDim.x * blockDim.x) {
    const float x = __ldg(level + c * ch + oc);
    const float t = __uint_as_float((unsigned)(tconf[c] & 0xffffffffull));
    float l, dx;
    bce_logits(x, t, hyp.obj_pw, hyp.gamma, &l, &dx);
    acc += (double)l;
    if (GRAD) grad[c * ch + oc] = gscale * dx;
  }
  acc = ry_warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) wpart[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += wpart[w];
    atomicAdd(&sums[lvl * 4 + 3], s);
  }
}

__global__ void loss_finalize_kernel(const double* __restrict__ sums, const int* __restrict__ counts, LevelArgs L,
                                     int csl, int nc, Hyp hyp, float* __restrict__ items) {
  float reg = 0.f, cls = 0.f, th = 0.f, conf = 0.f;
  for (int l = 0; l < 3; l++) {
    const int n = counts[l];
    if (n > 0) {
      reg += (float)(sums[l * 4 + 0] / (double)n);
      if (nc > 1) cls += (float)(sums[l * 4 + 1] / ((double)n * nc));
      if (csl) th += (float)(sums[l * 4 + 2] / ((double)n * 180.0));
    }
    if (L.cells[l] > 0) conf += (float)(sums[l * 4 + 3] / (double)L.cells[l]);
  }
  reg *= hyp.box; th *= hyp.theta; conf *= hyp.obj; cls *= hyp.cls;
  items[0] = reg;
  items[1] = th;
  items[2] = conf;
  items[3] = cls;
  items[4] = csl ? (reg + conf + cls + th) : (reg + conf + cls);   // lib/loss.py:255 / :413
  items[5] = (float)counts[0];
  items[6] = (float)counts[1];
  items[7] = (float)counts[2];
}

struct Layout {
  size_t blockcnt, blockoff, counts, sums, pos[3], tconf[3], total;
  int nblk;
  long long cap;
};

Layout loss_layout(int64_t B, int na, const int32_t* grid_hw, int64_t T) {
  Layout y;
  y.cap = 5ll * na * T;
  y.nblk = (int)((y.cap + kAssignBlock - 1) / kAssignBlock);
  if (y.nblk < 1) y.nblk = 1;
  size_t s = 0;
  y.blockcnt = s; s += ry_align_up((size_t)3 * y.nblk * 4, 256);
  y.blockoff = s; s += ry_align_up((size_t)3 * y.nblk * 4, 256);
  y.counts = s; s += 256;
  y.sums = s; s += 256;
  for (int l = 0; l < 3; l++) { y.pos[l] = s; s += ry_align_up((size_t)y.cap * sizeof(Pos), 256); }
  for (int l = 0; l < 3; l++) {
    y.tconf[l] = s;
    s += ry_align_up((size_t)B * na * grid_hw[2 * l] * grid_hw[2 * l + 1] * 8, 256);
  }
  y.total = s;
  return y;
}

int run_assign(const float* targets, int64_t T, int tcols, const float* anchors, int na, int rotated, int64_t B,
               LevelArgs& L, const Layout& y, char* w, cudaStream_t st) {
  int* blockcnt = (int*)(w + y.blockcnt);
  int* blockoff = (int*)(w + y.blockoff);
  int* counts = (int*)(w + y.counts);
  if (T == 0) {
    cudaMemsetAsync(counts, 0, 16, st);
    return RYOLO_OK;
  }
  dim3 grid(y.nblk, 3);
  assign_count_kernel<<<grid, kAssignBlock, 0, st>>>(targets, (int)T, tcols, anchors, na, rotated, (int)B, L, y.nblk,
                                                     blockcnt);
  assign_scan_kernel<<<3, 1024, 0, st>>>(blockcnt, y.nblk, blockoff, counts);
  assign_emit_kernel<<<grid, kAssignBlock, 0, st>>>(targets, (int)T, tcols, anchors, na, rotated, (int)B, L, y.nblk,
                                                    blockoff);
  return RYOLO_OK;
}

}  // namespace

extern "C" {

size_t ryolo_loss_workspace(int64_t B, int na, const int32_t* grid_hw, int64_t T) {
  return loss_layout(B, na, grid_hw, T).total;
}

size_t ryolo_pos_record_bytes(void) { return sizeof(Pos); }

// build_targets drop-in (lib/loss.py:270-331 csl / :427-492 kfiou).
//   targets  device [T, tcols] fp32 (img, cls, x, y, w, h, theta, ...);  anchors device [3, na, 3]
//   pos      device [3, cap] records (cap = 5*na*T), counts device int32[3]
int ryolo_build_targets(const float* targets, int64_t T, int tcols, int rotated, const float* anchors, int na,
                        int64_t B, const int32_t* grid_hw, void* pos, int32_t* counts, void* workspace,
                        size_t ws_bytes, void* stream) {
  RY_CHECK_ARG(T >= 0 && tcols >= 7 && na >= 1 && B >= 0, "build_targets: bad shape");
  RY_CHECK_ARG(5ll * na * T < (1ll << 31), "build_targets: too many candidate entries");
  Layout y = loss_layout(B, na, grid_hw, T);
  RY_CHECK_ARG(ws_bytes >= y.total, "build_targets: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  LevelArgs L{};
  for (int l = 0; l < 3; l++) {
    L.gh[l] = grid_hw[2 * l]; L.gw[l] = grid_hw[2 * l + 1];
    L.pos[l] = (Pos*)pos + (size_t)l * y.cap;
  }
  char* w = (char*)workspace;
  int rc = run_assign(targets, T, tcols, anchors, na, rotated, B, L, y, w, st);
  if (rc) return rc;
  cudaMemcpyAsync(counts, w + y.counts, 12, cudaMemcpyDeviceToDevice, st);
  RY_CHECK_LAUNCH();
  return RYOLO_OK;
}

// ComputeCSLLoss.__call__ (mode 0) / ComputeKFIoULoss.__call__ (mode 1), value + gradient.
//   levels   3 device pointers, [B, na, gh, gw, ch] fp32, ch = nc+185 (csl) or nc+6 (kfiou)
//   grads    3 device pointers of the same shapes, fully overwritten with d loss / d level; or NULL
//   hyp      host float[7] = box, obj, cls, obj_pw, cls_pw, fl_gamma, theta gain (lambda_theta, 0.5)
//   items    device float[8] = reg, theta, conf, cls, total, n_pos level 0..2
int ryolo_loss(int mode, const float* const* levels, float* const* grads, int64_t B, int na, int nc,
               const int32_t* grid_hw, const float* targets, int64_t T, int tcols, const float* anchors,
               const float* hyp, float* items, void* workspace, size_t ws_bytes, void* stream) {
  RY_CHECK_ARG(mode == 0 || mode == 1, "loss: mode must be 0 (csl) or 1 (kfiou)");
  RY_CHECK_ARG(B >= 0 && na >= 1 && nc >= 1 && T >= 0, "loss: bad shape");
  RY_CHECK_ARG(tcols >= (mode == 0 ? 187 : 7), "loss: targets have too few columns");
  RY_CHECK_ARG(5ll * na * T < (1ll << 31), "loss: too many candidate entries");
  Layout y = loss_layout(B, na, grid_hw, T);
  RY_CHECK_ARG(ws_bytes >= y.total, "loss: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  char* w = (char*)workspace;
  const int ch = mode == 0 ? nc + 185 : nc + 6;
  const int oc = mode == 0 ? 4 : 5;
  LevelArgs L{};
  for (int l = 0; l < 3; l++) {
    L.level[l] = levels[l];
    L.grad[l] = grads ? grads[l] : nullptr;
    L.gh[l] = grid_hw[2 * l]; L.gw[l] = grid_hw[2 * l + 1];
    L.cells[l] = (long long)B * na * L.gh[l] * L.gw[l];
    RY_CHECK_ARG(L.cells[l] < (1ll << 31), "loss: level too large for int32 cell index");
    L.pos[l] = (Pos*)(w + y.pos[l]);
    L.tconf[l] = (unsigned long long*)(w + y.tconf[l]);
  }
  Hyp h{hyp[0], hyp[1], hyp[2], hyp[3], hyp[4], hyp[5], hyp[6]};
  int* counts = (int*)(w + y.counts);
  double* sums = (double*)(w + y.sums);
  cudaMemsetAsync(sums, 0, 12 * sizeof(double), st);
  cudaMemsetAsync(w + y.tconf[0], 0, y.total - y.tconf[0], st);
  if (grads)
    for (int l = 0; l < 3; l++) cudaMemsetAsync(grads[l], 0, (size_t)L.cells[l] * ch * 4, st);
  int rc = run_assign(targets, T, tcols, anchors, na, mode == 1, B, L, y, w, st);
  if (rc) return rc;
  for (int l = 0; l < 3; l++) {
    if (T > 0) {
      if (mode == 0) {
        const int blocks = (int)std::min<long long>((y.cap * 32 + 255) / 256, 148 * 8);
        if (grads) csl_pos_kernel<true><<<blocks, 256, 0, st>>>(L, l, counts, targets, tcols, anchors, na, nc, h, sums);
        else csl_pos_kernel<false><<<blocks, 256, 0, st>>>(L, l, counts, targets, tcols, anchors, na, nc, h, sums);
      } else {
        const int blocks = (int)std::min<long long>((y.cap + 127) / 128, 148 * 16);
        if (grads) kfiou_pos_kernel<true><<<blocks, 128, 0, st>>>(L, l, counts, anchors, na, nc, h, sums);
        else kfiou_pos_kernel<false><<<blocks, 128, 0, st>>>(L, l, counts, anchors, na, nc, h, sums);
      }
    }
    if (L.cells[l] > 0) {
      const int blocks = (int)std::min<long long>((L.cells[l] + 255) / 256, 148 * 16);
      if (grads) obj_dense_kernel<true><<<blocks, 256, 0, st>>>(L, l, ch, oc, h, sums);
      else obj_dense_kernel<false><<<blocks, 256, 0, st>>>(L, l, ch, oc, h, sums);
    }
  }
  loss_finalize_kernel<<<1, 1, 0, st>>>(sums, counts, L, mode == 0, nc, h, items);
  RY_CHECK_LAUNCH();
  return RYOLO_OK;
}

}  // extern "C"
