// Test harness: compiles the PRODUCT's __host__ __device__ loss math (r-yolov4_b200/csrc/loss_math.cuh,
// assign.cuh) for the host so that it can be checked against the oracle without a GPU.
// Built by tests/test_loss_math_host.py with g++ -ffp-contract=off.
#include <cstdint>
#include "assign.cuh"

using namespace ryolo;

extern "C" {

void hm_ciou(int64_t n, const float* p, const float* t, float* out, float* grad) {
  for (int64_t i = 0; i < n; i++) {
    Box4 g;
    out[i] = ciou_fwd_bwd(Box4{p[4 * i], p[4 * i + 1], p[4 * i + 2], p[4 * i + 3]},
                          Box4{t[4 * i], t[4 * i + 1], t[4 * i + 2], t[4 * i + 3]}, &g);
    grad[4 * i] = g.x; grad[4 * i + 1] = g.y; grad[4 * i + 2] = g.w; grad[4 * i + 3] = g.h;
  }
}

void hm_kf(int64_t n, const float* p, const float* t, float* xy, float* kf, float* kfiou, float* grad) {
  for (int64_t i = 0; i < n; i++) {
    Box5 g;
    kf_fwd_bwd(Box5{p[5 * i], p[5 * i + 1], p[5 * i + 2], p[5 * i + 3], p[5 * i + 4]},
               Box5{t[5 * i], t[5 * i + 1], t[5 * i + 2], t[5 * i + 3], t[5 * i + 4]}, xy + i, kf + i, kfiou + i, &g);
    grad[5 * i] = g.x; grad[5 * i + 1] = g.y; grad[5 * i + 2] = g.w; grad[5 * i + 3] = g.h; grad[5 * i + 4] = g.r;
  }
}

void hm_bce(int64_t n, const float* x, const float* t, float pw, float gamma, float* loss, float* dx) {
  for (int64_t i = 0; i < n; i++) bce_logits(x[i], t[i], pw, gamma, loss + i, dx + i);
}

int64_t hm_pos_bytes() { return sizeof(Pos); }

// one level, reference emission order; out has room for 5*na*T records
int64_t hm_assign(const float* targets, int64_t T, int tcols, const float* anchors, int na, int rotated, int gh,
                  int gw, int nimg, Pos* out) {
  int64_t n = 0;
  for (int o = 0; o < 5; o++)
    for (int a = 0; a < na; a++)
      for (int64_t t = 0; t < T; t++)
        if (assign_entry(targets + t * tcols, o, a, anchors + a * 3, rotated, na, gh, gw, (int)t, nimg, out + n)) n++;
  return n;
}
}
